"""BASELINE config 5: seq_len sweep {2048, 4096, 8192, 16384}, batch 8, one GPU (+ the reference's own shapes): train-step (fwd+bwd+optimizer)
and forward-only time, algorithmic TFLOP/s and algorithmic HBM GB/s against the measured peaks."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import f_fwd, load_peaks
from oracle import denoiser_oracle as O
from osu_dreamer_b200.denoiser import default_args
from osu_dreamer_b200.trainer import DiffusionTrainer, LRScheduleArgs

ACT_BYTES_PER_TOKEN_FWD = 225952  # ideal-fusion activation traffic per token per forward (bf16)
peaks = load_peaks()
out = []
tr = DiffusionTrainer(val_batches=8, opt_args=dict(lr=3e-4, weight_decay=0.01),
                      schedule_args=LRScheduleArgs(warmup_steps=1000, warmup_init=0.3, decay_start=30000),
                      osl_weight=1.0, del_weight=30.0, emb_dim=6, a_dim=128, style_dim=32, diffusion_args=default_args())
sd = O.make_state_dict(1234)
tr.diffusion.load_state_dict(sd)
tr.diffusion_ema.module.load_state_dict(sd)
tr = tr.cuda()
# BASELINE configs[4] (B = 8) plus the shapes the reference itself runs: fit-denoiser at batch 128 x 152 frames
# (models/diffusion/model.yml:44,47) and predict on a ~4-minute song (l ~ 1500 latent frames, a few difficulties)
for B, L in ((8, 2048), (8, 4096), (8, 8192), (8, 16384), (128, 152), (4, 1482)):
    g = torch.Generator().manual_seed(L)
    h = torch.randn(B, 128, L, generator=g).cuda()
    x1 = torch.randn(B, 6, L, generator=g)
    x1 = (x1 * x1.pow(2).mean(1, keepdim=True).add(1e-6).rsqrt()).cuda()
    s = torch.randn(B, 32, generator=g).cuda()
    batch = (h, x1, s, torch.zeros(B, 5).cuda())
    tr.diffusion._rt.ws.clear()
    torch.cuda.empty_cache()
    for _ in range(3):
        tr.training_step(batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 4
    e0.record()
    for _ in range(n):
        tr.training_step(batch)
    e1.record()
    torch.cuda.synchronize()
    ms_train = e0.elapsed_time(e1) / n
    m = tr.diffusion_ema.module.eval()
    with torch.no_grad():
        for _ in range(2):
            m(h, s, x1)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            m(h, s, x1)
        e1.record()
        torch.cuda.synchronize()
    ms_fwd = e0.elapsed_time(e1) / n
    T = B * L
    row = {'L': L, 'B': B, 'ms_train_step': ms_train, 'ms_forward': ms_fwd,
           'train_samples_per_s': B / ms_train * 1e3,
           'train_tflops_algorithmic': 3 * f_fwd(L) * B / ms_train / 1e9,
           'fwd_tflops_algorithmic': f_fwd(L) * B / ms_fwd / 1e9,
           'fwd_hbm_gbs_algorithmic': (ACT_BYTES_PER_TOKEN_FWD * T + 93.8e6) / ms_fwd / 1e6,
           'fwd_frac_tensor_peak_sustained': f_fwd(L) * B / ms_fwd / 1e9 / peaks['tf_sustained'],
           'fwd_frac_hbm_peak': (ACT_BYTES_PER_TOKEN_FWD * T + 93.8e6) / ms_fwd / 1e6 / peaks['hbm_gbs'],
           'attention_share_of_flops': 4096 * L * 8 / (8 * (8523776 + 4096 * L) + 54436)}
    print(json.dumps(row), flush=True)
    out.append(row)
os.makedirs('gpurun_out', exist_ok=True)
json.dump({'peaks': peaks, 'rows': out}, open('gpurun_out/seq_sweep.json', 'w'), indent=1)
