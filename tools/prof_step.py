"""Driver for ncu: one training step (bench shape by default) bracketed by cudaProfilerStart/Stop, plus two sampler
forwards, so that `ncu --profile-from-start off` sees exactly one step's launches.  Usage:
  ncu --profile-from-start off --section SpeedOfLight --metrics dram__bytes_read.sum,dram__bytes_write.sum \
      -k regex:'^(?!.*attn_)' --csv --log-file out.csv python tools/prof_step.py [B] [L]
then tools/agg_ncu_csv.py out.csv"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import denoiser_oracle as O
from osu_dreamer_b200.denoiser import default_args
from osu_dreamer_b200.trainer import DiffusionTrainer, LRScheduleArgs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
L = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
tr = DiffusionTrainer(val_batches=8, opt_args=dict(lr=3e-4, weight_decay=0.01),
                      schedule_args=LRScheduleArgs(warmup_steps=1000, warmup_init=0.3, decay_start=30000),
                      osl_weight=1.0, del_weight=30.0, emb_dim=6, a_dim=128, style_dim=32, diffusion_args=default_args())
sd = O.make_state_dict(1234)
tr.diffusion.load_state_dict(sd)
tr.diffusion_ema.module.load_state_dict(sd)
tr = tr.cuda()
g = torch.Generator().manual_seed(5)
h = torch.randn(B, 128, L, generator=g).cuda()
x1 = torch.randn(B, 6, L, generator=g)
x1 = (x1 * x1.pow(2).mean(1, keepdim=True).add(1e-6).rsqrt()).cuda()
s = torch.randn(B, 32, generator=g).cuda()
batch = (h, x1, s, torch.zeros(B, 5).cuda())
tr.training_step(batch)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.training_step(batch)
torch.cuda.synchronize()
if os.environ.get('PROF_SAMPLER', '0') == '1':
    m = tr.diffusion_ema.module.eval()
    with torch.no_grad():
        m.sample(h, s, 1)
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('done')
