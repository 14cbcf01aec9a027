"""GPU parity of the backward path: attention gradients vs torch autograd, parameter gradients of the
reference trainer loss vs the golden summaries produced by the unmodified reference (CPU fp32)."""
import os

import numpy as np
import pytest
import torch

from oracle import denoiser_oracle as O

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('B,L', [(1, 128), (2, 320), (1, 1000), (1, 60)])
def test_attention_backward_matches_autograd(B, L, fused):
    """both attention-backward paths: two kernels (dK/dV, dQ) and the single-pass kernel with TMA reduce-add dQ"""
    from osu_dreamer_b200 import lib
    g = torch.Generator().manual_seed(L + 1)
    qkv = (torch.randn(B * L, 3072, generator=g)).cuda().to(torch.bfloat16)
    dy = torch.randn(B * L, 1024, generator=g).cuda().to(torch.bfloat16)
    y, lse = lib.attn_fwd(qkv, B, L)
    dqkv = (lib.attn_bwd_fused if fused else lib.attn_bwd)(qkv, y, dy, lse, B, L)
    torch.cuda.synchronize()
    ref_in = qkv.float().clone().requires_grad_(True)
    q, k, v = ref_in.view(B, L, 3, 16, 64).permute(2, 0, 3, 1, 4)
    out = torch.softmax((q @ k.transpose(-1, -2)) / 8, -1) @ v
    out = out.permute(0, 2, 1, 3).reshape(B * L, 1024)
    out.backward(dy.float())
    gr = ref_in.grad
    for name, sl in [('dq', slice(0, 1024)), ('dk', slice(1024, 2048)), ('dv', slice(2048, 3072))]:
        e = _rel(dqkv[:, sl], gr[:, sl])
        print(name, e)
        assert e < 2e-2, (name, e)


@pytest.mark.parametrize('B,L', [(1, 512), (3, 1920)])
def test_attention_backward_fused_falls_back_on_spread_statistics(B, L):
    """a query row whose lse is > 96 octaves above its tile neighbours: the single-pass call must detect it on the
    device and produce the two-kernel result (which makes no assumption on the statistics).  The fallback kernels are
    launched on a compact grid (two CTAs per SM) and walk their work items; at (3, 1920) there are 720 items, so every
    CTA re-initialises its barriers and runs two or three of them -- bit-identical to the one-CTA-per-item launch."""
    from osu_dreamer_b200 import lib
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(B * L, 3072, generator=g)
    qkv[3, :1024] *= 40.0
    qkv = qkv.cuda().to(torch.bfloat16)
    dy = torch.randn(B * L, 1024, generator=g).cuda().to(torch.bfloat16)
    y, lse = lib.attn_fwd(qkv, B, L)
    assert float(lse.max() - lse.min()) * 1.4427 > 96
    ref = lib.attn_bwd(qkv, y, dy, lse, B, L)
    got = lib.attn_bwd_fused(qkv, y, dy, lse, B, L)
    torch.cuda.synchronize()
    assert torch.isfinite(got.float()).all()
    assert torch.equal(got, ref)


@pytest.mark.parametrize('B,L', [(1, 8192), (2, 4000)])
def test_attention_backward_paths_agree_at_full_length(B, L):
    """BASELINE sequence length (and a ragged one): the single-pass kernel (TMA reduce-add dQ, w-scaled dO) and the
    two-kernel path are independent implementations; both were checked against autograd at small sizes above."""
    from osu_dreamer_b200 import lib
    g = torch.Generator().manual_seed(L)
    qkv = torch.randn(B * L, 3072, generator=g).cuda().to(torch.bfloat16)
    dy = torch.randn(B * L, 1024, generator=g).cuda().to(torch.bfloat16)
    y, lse = lib.attn_fwd(qkv, B, L)
    a = lib.attn_bwd(qkv, y, dy, lse, B, L)
    b = lib.attn_bwd_fused(qkv, y, dy, lse, B, L)
    torch.cuda.synchronize()
    assert torch.isfinite(b.float()).all()
    for name, sl in [('dq', slice(0, 1024)), ('dk', slice(1024, 2048)), ('dv', slice(2048, 3072))]:
        e = _rel(b[:, sl], a[:, sl])
        rl2 = float((b[:, sl].float() - a[:, sl].float()).norm() / a[:, sl].float().norm())
        print(name, e, rl2)
        assert e < 2e-2 and rl2 < 1e-2, (name, e, rl2)
    # softmax-Jacobian property: every row of dS sums to zero, hence sum_j dq[i, :] . 1 ... checked through dk:
    # sum over kv of dK equals (dS^T Q) summed = Q^T (dS 1) = 0 only row-wise in dS; use the cheap global identity
    # sum_i q_i . dq_i == sum_j k_j . dk_j (both equal sum_ij dS_ij S_ij * 8)
    q, k = qkv[:, :1024].float(), qkv[:, 1024:2048].float()
    lhs = float((q * b[:, :1024].float()).sum())
    rhs = float((k * b[:, 1024:2048].float()).sum())
    assert abs(lhs - rhs) <= 2e-2 * max(abs(lhs), abs(rhs), 1.0), (lhs, rhs)


def _trainer_grads(sd, inp, x0, t):
    """parameter gradients of the reference loss through the CUDA path (torch ops only for the tiny loss)."""
    from osu_dreamer_b200.denoiser import DiffusionModel, default_args
    m = DiffusionModel(6, 128, 32, default_args())
    m.load_state_dict(sd)
    m = m.cuda().train()
    h, x1, s = inp['h'].cuda(), inp['x1'].cuda(), inp['s'].cuda()
    x0, t = x0.cuda(), t.cuda()
    xt = torch.lerp(x0, x1, t[:, None, None])
    u_pred, v_pred = m(h, s, xt)
    d_sq = O.frame_dist_sq(xt, x1)
    u_target = (d_sq + m.c0).sqrt()
    osl = (O.frame_dist_sq(xt - u_pred[:, None, None] * v_pred, x1) / (d_sq + m.c0)).mean()
    del_ = O.frame_dist_sq(v_pred, (xt - x1) / u_target[:, None, None]).mean()
    loss = osl + 30.0 * del_
    loss.backward()
    torch.cuda.synchronize()
    return m, float(loss)


def test_trainer_gradients_match_reference_golden(golden_dir, oracle_sd):
    g = np.load(os.path.join(golden_dir, 'loss_B2_L128.npz'))
    inp = O.make_inputs(2, 128, seed=21)
    m, loss = _trainer_grads(oracle_sd, inp, torch.from_numpy(g['x0']), torch.from_numpy(g['t']))
    print('loss', loss, 'ref', float(g['loss']))
    assert abs(loss - float(g['loss'])) < 2e-2 * abs(float(g['loss']))
    worst = []
    for (name, p), gn, gh in zip(m.named_parameters(), g['grad_norms'], g['grad_heads']):
        assert name == str(g['grad_names'][list(dict(m.named_parameters()).keys()).index(name)])
        n_mine = float(p.grad.double().norm())
        rel_norm = abs(n_mine - gn) / max(gn, 1e-12)
        k = min(8, p.numel())
        head_err = float(np.abs(p.grad.reshape(-1)[:k].cpu().numpy() - gh[:k]).max() / (gn / max(1.0, p.numel() ** 0.5) + 1e-12))
        worst.append((rel_norm, head_err, name))
    worst.sort(reverse=True)
    print('worst norm errors:', [(round(a, 4), n) for a, _, n in worst[:8]])
    print('worst head errors:', sorted([(round(h, 3), n) for _, h, n in worst], reverse=True)[:8])
    assert worst[0][0] < 2e-2, worst[:5]  # north_star bf16 tolerance (measured <= 1.1e-2)


def test_trainer_gradients_match_live_oracle(oracle_sd):
    """full per-element comparison against the CPU oracle's autograd at B=2, L=256."""
    inp = O.make_inputs(2, 256, seed=77)
    sd = {k: v.clone().requires_grad_(True) for k, v in oracle_sd.items()}
    loss_ref, _ = O.trainer_loss(sd, inp['h'], inp['x1'], inp['s'], inp['x0'], inp['t'])
    loss_ref.backward()
    m, loss = _trainer_grads(oracle_sd, inp, inp['x0'], inp['t'])
    assert abs(loss - float(loss_ref)) < 2e-2 * abs(float(loss_ref))
    errs = []
    for name, p in m.named_parameters():
        gr = sd[name].grad
        # cosine-type error: ||g - g_ref|| / ||g_ref||
        e = float((p.grad.cpu().double() - gr.double()).norm() / gr.double().norm().clamp_min(1e-20))
        errs.append((e, name))
    errs.sort(reverse=True)
    print('worst relative L2 gradient errors:', [(round(e, 4), n) for e, n in errs[:10]])
    assert errs[0][0] < 2e-2, errs[:5]  # north_star bf16 tolerance (measured <= 1.2e-2)
