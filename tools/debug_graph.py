import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import denoiser_oracle as O
from osu_dreamer_b200.denoiser import DiffusionModel, default_args
def mn(a, b): return float((a - b).abs().max() / b.abs().max())
m = DiffusionModel(6, 128, 32, default_args()); m.load_state_dict(O.make_state_dict(1234)); m = m.cuda().eval()
outs = {}
for mode in (False, 'again', True):
    m.graph_sampler = bool(mode is True)
    res = []
    for seed in (41, 42, 43, 41):
        inp = O.make_inputs(2, 192, seed=seed)
        x0 = torch.randn(2, 6, 192, generator=torch.Generator().manual_seed(seed)).cuda()
        res.append(m.sample_from(inp['h'].cuda(), inp['s'].cuda(), x0, 4).cpu())
        print(mode, seed, 'finite', bool(torch.isfinite(res[-1]).all()), 'absmax', float(res[-1].abs().max()), 'eta_u0', m.last_eta_u0.tolist())
    outs[mode] = res
for i in range(4):
    print(i, 'eager-vs-eager', mn(outs['again'][i], outs[False][i]), 'graph-vs-eager', mn(outs[True][i], outs[False][i]))
print('replay same inputs', mn(outs[True][3], outs[True][0]), 'different inputs', mn(outs[True][1], outs[True][0]))
