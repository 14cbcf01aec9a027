"""fp32-grade path (precision='fp32'): which split-product terms of the attention kernel are worth their energy?
For each (qk_mask, pv_mask) -- bit 1 = lo*hi, bit 2 = hi*lo, hi*hi always on -- in its own process (OSD_X3_TERMS is read once):
  * error of one forward at B=2, L=8192 and of the 64-step sampler at B=1, L=8192 against the oracle (strict fp32 on the GPU);
  * time of one forward at B=32, L=8192 (x65 = one 64-step sample of BASELINE configs[2]).
-> gpurun_out/x3_terms_sweep.json"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, time
sys.path.insert(0, %r)
import torch
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
from oracle import denoiser_oracle as O
from osu_dreamer_b200.denoiser import DiffusionModel, default_args
sd = O.make_state_dict(1234)
sdc = {k: v.cuda() for k, v in sd.items()}
m = DiffusionModel(6, 128, 32, default_args()); m.load_state_dict(sd); m = m.cuda().eval(); m.precision = 'fp32'
def mx(a, b): return float((a.double() - b.double()).abs().max() / b.double().abs().max())
out = {}
with torch.no_grad():
    inp = {k: v.cuda() for k, v in O.make_inputs(2, 8192, seed=91).items()}
    xt = torch.lerp(inp['x0'], inp['x1'], inp['t'][:, None, None])
    ur, vr = O.forward(sdc, inp['h'], inp['s'], xt)
    u, v = m(inp['h'], inp['s'], xt)
    out['fwd_v_err'] = mx(v, vr); out['fwd_u_err'] = mx(u, ur)
    inp = {k: v.cuda() for k, v in O.make_inputs(1, 8192, seed=95).items()}
    x_init = torch.randn(1, 6, 8192, generator=torch.Generator().manual_seed(96)).cuda()
    xr, u0, eta = O.sample(sdc, inp['h'], inp['s'], x_init, 64)
    m.graph_sampler = False
    x = m.sample_from(inp['h'], inp['s'], x_init, 64)
    out['sample64_err'] = mx(x, xr)
    del xr, x, ur, vr
    torch.cuda.empty_cache()
    B = 32
    inp = {k: v.cuda() for k, v in O.make_inputs(B, 8192, seed=3).items()}
    m._rt.reset()
    for _ in range(2): m(inp['h'], inp['s'], inp['x0'])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4): m(inp['h'], inp['s'], inp['x0'])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 4
    out['forward_ms_B32'] = ms
    out['latents_per_s_est'] = B / (65 * ms * 1e-3)
print('RESULT ' + json.dumps(out))
''' % ROOT

res = []
combos = ((7, 7), (3, 7), (5, 7), (7, 3), (7, 5), (3, 3), (3, 5), (3, 1), (1, 1))
if len(sys.argv) > 1 and sys.argv[1] == 'short':
    combos = ((7, 7), (3, 5))
for qk, pv in combos + ((0, 0),):  # (0, 0) = the shipped kernel: the db pipeline with the (3, 5) terms (attn_fwd_db.cu, X3)
    env = dict(os.environ, OSD_X3_TERMS=f'{qk},{pv}', OSD_X3_KERNEL='old')
    if (qk, pv) == (0, 0):
        env.pop('OSD_X3_KERNEL')
    r = subprocess.run([sys.executable, '-c', CHILD], env=env, capture_output=True, text=True, timeout=900)
    line = [ln for ln in r.stdout.splitlines() if ln.startswith('RESULT ')]
    d = json.loads(line[0][7:]) if line else {'error': (r.stderr or r.stdout)[-500:]}
    d.update(qk_mask=qk or 3, pv_mask=pv or 5, mma_units=bin(qk or 3).count('1') + bin(pv or 5).count('1'),
             kernel='attn_fwd_db_kernel<X3>' if (qk, pv) == (0, 0) else 'attn_fwd_x3_kernel')
    res.append(d)
    print(json.dumps(d), flush=True)
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'x3_terms_sweep.json'), 'w'), indent=1)
