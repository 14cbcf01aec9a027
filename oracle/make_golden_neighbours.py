"""Generate tests/golden/nb_*.npz + nb_spec.json by running the UNMODIFIED reference LatentModel / StyleModel
(imported from /root/reference with the stub modules of oracle/refimport.py).  Run in the build container only:
    python -m oracle.make_golden_neighbours
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

from . import neighbours_oracle as N
from . import refimport

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def main():
    torch.set_num_threads(8)
    refimport._install_stubs()
    if refimport.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, refimport.REFERENCE_ROOT)
    from osu_dreamer.models.latent.model import LatentModel, LatentModelArgs
    from osu_dreamer.models.latent.unet import LayerArgs
    from osu_dreamer.models.style.model import StyleModel, StyleModelArgs
    hp = N.LATENT_HP
    lm = LatentModel(hp['emb_dim'], hp['style_dim'], hp['n_downs'], hp['stride'],
                     LatentModelArgs(h_dim=hp['h_dim'], ae_args=LayerArgs(n_layers=hp['n_layers'], expand=hp['expand'], radius=hp['radius']),
                                     style_head_dim=64, style_heads=16)).eval()
    sm = StyleModel(N.STYLE_HP['style_dim'], StyleModelArgs(label_features=N.STYLE_HP['label_features'], h_dim=N.STYLE_HP['h_dim'],
                                                             depth=N.STYLE_HP['depth'], expand=N.STYLE_HP['expand'])).eval()
    spec = {'latent': [(k, list(v.shape)) for k, v in lm.state_dict().items()],
            'style': [(k, list(v.shape)) for k, v in sm.state_dict().items()]}
    json.dump(spec, open(os.path.join(OUT, 'nb_spec.json'), 'w'))
    lsd = N.seeded_state_dict(spec['latent'], 4321)
    ssd = N.seeded_state_dict(spec['style'], 8765)
    lm.load_state_dict(lsd, strict=True)
    sm.load_state_dict(ssd, strict=True)

    g = torch.Generator().manual_seed(11)
    for tag, (Ba, B, l) in {'a': (1, 2, 8), 'b': (1, 3, 5)}.items():
        audio = torch.randn(Ba, N.A_DIM, 27 * l, generator=g)
        z = torch.randn(B, 6, l, generator=g)
        z = z * z.pow(2).mean(1, keepdim=True).add(1e-6).rsqrt()
        s = torch.randn(B, 32, generator=g)
        s = s * s.pow(2).mean(1, keepdim=True).add(1e-6).rsqrt()
        with torch.no_grad():
            skips, h = lm.audio_encoder(audio)
            chart, labels = lm.decode(z, s, skips=list(skips))
            # fp64 run of the same modules: the bar a GPU implementation is measured against
            lm64 = lm.double()
            skips64, h64 = lm64.audio_encoder(audio.double())
            chart64, labels64 = lm64.decode(z.double(), s.double(), skips=list(skips64))
            lm.float()
        np.savez(os.path.join(OUT, f'nb_latent_{tag}.npz'), audio=audio.numpy(), z=z.numpy(), s=s.numpy(), h=h.numpy(),
                 skip0=skips[0].numpy(), skip1=skips[1].numpy(), skip2=skips[2].numpy(), chart=chart.numpy(), labels=labels.numpy(),
                 h64=h64.numpy(), chart64=chart64.numpy(), labels64=labels64.numpy())
        print('latent', tag, 'h absmax', float(h.abs().max()), 'chart absmax', float(chart.abs().max()),
              'fp32-vs-fp64 chart', float((chart.double() - chart64).abs().max()))

    labels = 10 * torch.rand(4, 5, generator=g)
    labels[1, 2] = -1.0  # dropped label -> learned null embedding (models/style/model.py:78)
    labels[3, :] = -1.0
    st = torch.randn(4, 32, generator=g)
    with torch.no_grad():
        u, v = sm(st, labels)
        torch.manual_seed(99)
        s_fin = sm.sample(labels, 16)
        torch.manual_seed(99)
        s_init = torch.randn(4, 32)
        sm64 = sm.double()
        u64, v64 = sm64(st.double(), labels.double())
        sm.float()
    np.savez(os.path.join(OUT, 'nb_style.npz'), labels=labels.numpy(), st=st.numpy(), u=u.numpy(), v=v.numpy(), s_init=s_init.numpy(),
             s_final=s_fin.numpy(), u64=u64.numpy(), v64=v64.numpy(), c0=sm.c0, u_scale=sm.u_scale,
             n_params_latent=sum(p.numel() for p in lm.parameters()), n_params_style=sum(p.numel() for p in sm.parameters()))
    print('style u', u.tolist(), 'v absmax', float(v.abs().max()), 's_final absmax', float(s_fin.abs().max()))

    # ---- StyleTrainer.forward (models/style/train.py:48-91): loss + gradient norms with seed-matched random draws
    from osu_dreamer.common.lr_schedule import LRScheduleArgs
    from osu_dreamer.models.style.train import StyleTrainer
    tr = StyleTrainer(opt_args=dict(lr=3e-4, weight_decay=0.01), schedule_args=LRScheduleArgs(), label_drop_prob=0.2, osl_weight=1.0,
                      del_weight=30.0, style_dim=32,
                      style_args=StyleModelArgs(label_features=N.STYLE_HP['label_features'], h_dim=N.STYLE_HP['h_dim'],
                                                depth=N.STYLE_HP['depth'], expand=N.STYLE_HP['expand']))
    torch.set_float32_matmul_precision('highest')  # the constructor sets 'medium' process-wide (train.py:33)
    tr.style.load_state_dict(ssd, strict=True)
    g2 = torch.Generator().manual_seed(17)
    B = 6
    s1 = torch.randn(B, 32, generator=g2)
    s1 = s1 * s1.pow(2).mean(1, keepdim=True).add(1e-6).rsqrt()
    lab = 10 * torch.rand(B, 5, generator=g2)
    torch.manual_seed(123)  # the draws of train.py:58-64, in order: randperm, rand, randn_like(s1), rand_like(labels)
    uu = (torch.randperm(B) + torch.rand(B)) / B
    t = torch.special.ndtri(uu.clamp(1e-6, 1 - 1e-6)).sigmoid()
    s0 = torch.randn_like(s1)
    drop = torch.rand_like(lab) < 0.2
    torch.manual_seed(123)
    loss, log = tr(tr.style, torch.empty(B, 0, 0), torch.empty(B, 0, 0), s1, lab)
    loss.backward()
    gn = {k: float(p.grad.norm()) for k, p in tr.style.named_parameters()}
    np.savez(os.path.join(OUT, 'nb_style_loss.npz'), s1=s1.numpy(), labels=lab.numpy(), t=t.numpy(), s0=s0.numpy(), drop=drop.numpy(),
             loss=float(loss.detach()), osl=float(log['osl']), del_=float(log['del']), u_mape=float(log['u_mape']),
             grad_names=np.array(list(gn.keys())), grad_norms=np.array(list(gn.values())))
    print('style loss', float(loss.detach()), {k: float(v) for k, v in log.items()})


if __name__ == '__main__':
    main()
