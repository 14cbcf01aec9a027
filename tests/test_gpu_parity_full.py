"""GPU parity at the shapes the metric is quoted on (BASELINE.json configs[1..3]: seq_len 8192, batch 16 / 32, 64 steps).

The checker is `oracle/denoiser_oracle.py` -- the restatement pinned to the unmodified reference by tests/test_oracle.py --
evaluated in strict fp32 (TF32 off for cuBLAS and cuDNN) ON THE GPU, because the host cores cannot finish L = 8192 in
test time.  `test_gpu_oracle_is_the_pinned_cpu_oracle` shows that moving the oracle to the device does not change what it
computes (same functions, same torch ops; only the summation order of the fp32 library kernels differs).  Torch is test
infrastructure here; the product path under test is libosd_b200.so through the C ABI.

Tolerances are north_star's: 2e-2 (bf16 operands) / 1e-3 (precision='fp32'), max-normalised against the fp32 oracle for
outputs, relative L2 per tensor for gradients.
"""
import math
import os

import numpy as np
import pytest
import torch

from oracle import denoiser_oracle as O

pytestmark = pytest.mark.gpu

BF16_TOL = 2e-2
FP32_TOL = 1e-3
SEQ = 8192


@pytest.fixture(autouse=True)
def _strict_fp32():
    """fp32 means fp32 for the oracle: no TF32 in cuBLAS matmuls nor in cuDNN convolutions"""
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision('highest')
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    torch.cuda.empty_cache()


def _model(sd, precision='bf16', train=False):
    from osu_dreamer_b200.denoiser import DiffusionModel, default_args
    m = DiffusionModel(6, 128, 32, default_args())
    m.load_state_dict(sd)
    m = m.cuda()
    m.precision = precision
    return m.train() if train else m.eval()


def _maxnorm(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double().to(a.device)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


def _cuda(d):
    return {k: v.cuda() for k, v in d.items()}


def _oracle_loss_grads(sd_cuda, inp, per_sample):
    """parameter gradients of the reference trainer loss (train.py:78-101) by the oracle's autograd on the device.  The
    loss is a mean over samples of per-sample terms, so with `per_sample` the batch is walked one sample at a time
    (the [B,16,L,L] fp32 scores of a whole batch do not fit) and the per-sample losses / gradients are averaged."""
    B = inp['x1'].shape[0]
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in sd_cuda.items()}
    total = 0.0
    groups = [slice(b, b + 1) for b in range(B)] if per_sample else [slice(0, B)]
    for g in groups:
        n = g.stop - g.start
        loss, _ = O.trainer_loss(sd, inp['h'][g], inp['x1'][g], inp['s'][g], inp['x0'][g], inp['t'][g])
        (loss * (n / B)).backward()
        total += float(loss) * n / B
    return total, {k: v.grad for k, v in sd.items()}


def _our_loss_grads(sd, inp):
    m = _model(sd, train=True)
    xt = torch.lerp(inp['x0'], inp['x1'], inp['t'][:, None, None])
    u_pred, v_pred = m(inp['h'], inp['s'], xt)
    d_sq = O.frame_dist_sq(xt, inp['x1'])
    u_target = (d_sq + m.c0).sqrt()
    osl = (O.frame_dist_sq(xt - u_pred[:, None, None] * v_pred, inp['x1']) / (d_sq + m.c0)).mean()
    del_ = O.frame_dist_sq(v_pred, (xt - inp['x1']) / u_target[:, None, None]).mean()
    loss = osl + 30.0 * del_
    loss.backward()
    torch.cuda.synchronize()
    grads = {n: p.grad for n, p in m.named_parameters()}
    m._rt.reset()
    return float(loss), grads


def _grad_errors(ours, ref):
    errs = []
    for name, gr in ref.items():
        e = float((ours[name].double() - gr.double()).norm() / gr.double().norm().clamp_min(1e-30))
        errs.append((e, name))
    errs.sort(reverse=True)
    return errs


# ------------------------------------------------------------------------------------------------ the checker itself
def test_gpu_oracle_is_the_pinned_cpu_oracle(oracle_sd):
    """the oracle evaluated on the device (strict fp32) vs the same oracle on the host: outputs, loss and gradients"""
    inp = O.make_inputs(2, 256, seed=77)
    xt = torch.lerp(inp['x0'], inp['x1'], inp['t'][:, None, None])
    with torch.no_grad():
        uc, vc = O.forward(oracle_sd, inp['h'], inp['s'], xt)
        ug, vg = O.forward(_cuda(oracle_sd), inp['h'].cuda(), inp['s'].cuda(), xt.cuda())
    assert _maxnorm(ug.cpu(), uc) < 1e-5 and _maxnorm(vg.cpu(), vc) < 1e-5
    sd = {k: v.clone().requires_grad_(True) for k, v in oracle_sd.items()}
    loss_c, _ = O.trainer_loss(sd, inp['h'], inp['x1'], inp['s'], inp['x0'], inp['t'])
    loss_c.backward()
    loss_g, grads_g = _oracle_loss_grads(_cuda(oracle_sd), _cuda(inp), per_sample=True)
    assert abs(loss_g - float(loss_c)) < 1e-5 * abs(float(loss_c))
    errs = _grad_errors({k: v.cpu() for k, v in grads_g.items()}, {k: v.grad for k, v in sd.items()})
    print('device oracle vs host oracle, worst gradient rel-L2:', errs[:3])
    assert errs[0][0] < 1e-4


# ------------------------------------------------------------------------------------------------ forward, L = 8192
@pytest.mark.parametrize('precision,tol', [('bf16', BF16_TOL), ('fp32', FP32_TOL)])
def test_forward_seq8192_matches_oracle(oracle_sd, precision, tol):
    inp = _cuda(O.make_inputs(2, SEQ, seed=91))
    xt = torch.lerp(inp['x0'], inp['x1'], inp['t'][:, None, None])
    with torch.no_grad():
        ur, vr = O.forward(_cuda(oracle_sd), inp['h'], inp['s'], xt)
        m = _model(oracle_sd, precision)
        u, v = m(inp['h'], inp['s'], xt)
    torch.cuda.synchronize()
    eu, ev = _maxnorm(u, ur), _maxnorm(v, vr)
    print(f'L={SEQ} B=2 {precision}: u err {eu:.3e}, v err {ev:.3e} (tolerance {tol:g})')
    assert eu < tol and ev < tol


def test_forward_config2_batch16_seq8192_matches_oracle(oracle_sd):
    """BASELINE configs[1] exactly (B = 16, L = 8192, bf16): one forward of the whole batch through the CUDA path; samples 0,
    7 and 15 are checked against the oracle run on that sample alone (samples do not interact in the forward)."""
    B = 16
    inp = _cuda(O.make_inputs(B, SEQ, seed=92))
    xt = torch.lerp(inp['x0'], inp['x1'], inp['t'][:, None, None])
    m = _model(oracle_sd)
    with torch.no_grad():
        u, v = m(inp['h'], inp['s'], xt)
        sdc = _cuda(oracle_sd)
        for b in (0, 7, 15):
            ur, vr = O.forward(sdc, inp['h'][b:b + 1], inp['s'][b:b + 1], xt[b:b + 1])
            eu, ev = _maxnorm(u[b:b + 1], ur), _maxnorm(v[b:b + 1], vr)
            print(f'config 2, sample {b}: u err {eu:.3e}, v err {ev:.3e}')
            assert eu < BF16_TOL and ev < BF16_TOL


# ------------------------------------------------------------------------------------------------ gradients, L = 8192
def test_gradients_seq8192_match_oracle(oracle_sd):
    """all 164 parameter gradients of the trainer loss at L = 8192 (B = 1) against the oracle's autograd"""
    inp = _cuda(O.make_inputs(1, SEQ, seed=93))
    loss_r, gr = _oracle_loss_grads(_cuda(oracle_sd), inp, per_sample=True)
    torch.cuda.empty_cache()
    loss, g = _our_loss_grads(oracle_sd, inp)
    errs = _grad_errors(g, gr)
    print(f'L={SEQ} B=1: loss {loss:.6f} vs oracle {loss_r:.6f}; worst gradient rel-L2 errors: '
          f'{[(round(e, 4), n) for e, n in errs[:6]]}')
    assert abs(loss - loss_r) < BF16_TOL * abs(loss_r)
    assert errs[0][0] < BF16_TOL, errs[:5]


def test_gradients_config2_batch16_seq8192_match_oracle(oracle_sd):
    """BASELINE configs[1] exactly: the fit-denoiser loss gradient of a B = 16, L = 8192 batch through the CUDA path vs the
    oracle, which walks the batch one sample at a time (the loss is a mean of per-sample terms)"""
    B = 16
    inp = _cuda(O.make_inputs(B, SEQ, seed=94))
    loss_r, gr = _oracle_loss_grads(_cuda(oracle_sd), inp, per_sample=True)
    torch.cuda.empty_cache()
    loss, g = _our_loss_grads(oracle_sd, inp)
    errs = _grad_errors(g, gr)
    print(f'config 2 (B=16, L={SEQ}): loss {loss:.6f} vs oracle {loss_r:.6f}; worst gradient rel-L2 errors: '
          f'{[(round(e, 4), n) for e, n in errs[:6]]}')
    assert abs(loss - loss_r) < BF16_TOL * abs(loss_r)
    assert errs[0][0] < BF16_TOL, errs[:5]


# ------------------------------------------------------------------------------------------------ 64-step sampler
@pytest.mark.parametrize('B,L', [(2, 512), (1, SEQ)])
@pytest.mark.parametrize('precision,tol', [('bf16', BF16_TOL), ('fp32', FP32_TOL)])
def test_sampler_64_steps_matches_oracle(oracle_sd, B, L, precision, tol):
    """the metric's sampler: 64 steps = 65 chained forwards (model.py:117-138), same initial noise, against the oracle"""
    inp = _cuda(O.make_inputs(B, L, seed=95))
    x_init = torch.randn(B, 6, L, generator=torch.Generator().manual_seed(96)).cuda()
    xr, u0, eta = O.sample(_cuda(oracle_sd), inp['h'], inp['s'], x_init, 64)
    m = _model(oracle_sd, precision)
    m.graph_sampler = False
    x = m.sample_from(inp['h'], inp['s'], x_init, 64)
    torch.cuda.synchronize()
    err = _maxnorm(x, xr)
    eta_u0 = m.last_eta_u0.cpu()
    print(f'64-step sampler B={B} L={L} {precision}: x err {err:.3e} (tolerance {tol:g}); eta {float(eta_u0[0]):.7f} vs '
          f'{eta:.7f}, u0 {float(eta_u0[1]):.6f} vs {u0:.6f}')
    assert abs(float(eta_u0[1]) - u0) < tol * u0 and abs(float(eta_u0[0]) - eta) < tol * eta
    assert torch.isfinite(x).all() and err < tol


@pytest.mark.parametrize('precision,tol', [('fp32', FP32_TOL), ('bf16', BF16_TOL)])
def test_sampler_config3_batch32_seq8192_matches_oracle(oracle_sd, precision, tol):
    """BASELINE configs[2] exactly (64 steps, B = 32, L = 8192, the fp32 tolerance check -- and the bf16 path): the whole
    batch is sampled by the CUDA path.  Samples only interact through u0 = mean(u) of the probe forward, so the oracle
    (a) computes the probe's u for all 32 samples, one at a time, which pins u0 and eta, and (b) samples two of the 32 with
    that u0 injected; those two latents are compared."""
    B = 32
    inp = _cuda(O.make_inputs(B, SEQ, seed=97))
    x_init = torch.randn(B, 6, SEQ, generator=torch.Generator().manual_seed(98)).cuda()
    sdc = _cuda(oracle_sd)
    with torch.no_grad():
        us = [O.forward(sdc, inp['h'][b:b + 1], inp['s'][b:b + 1], x_init[b:b + 1])[0] for b in range(B)]
    u0 = float(torch.cat(us).mean())
    m = _model(oracle_sd, precision)
    x = m.sample_from(inp['h'], inp['s'], x_init, 64)
    torch.cuda.synchronize()
    eta_u0 = m.last_eta_u0.cpu()
    assert abs(float(eta_u0[1]) - u0) < tol * u0
    m._rt.reset()
    torch.cuda.empty_cache()
    for b in (3, 31):
        xr, _, eta = O.sample(sdc, inp['h'][b:b + 1], inp['s'][b:b + 1], x_init[b:b + 1], 64, u0=u0)
        err = _maxnorm(x[b:b + 1], xr)
        print(f'config 3 ({precision}), sample {b}: x err {err:.3e} (tolerance {tol:g}); eta {float(eta_u0[0]):.7f} vs {eta:.7f}')
        assert err < tol


# ------------------------------------------------------------------------------------------------ validation_step values
def test_validation_step_values_match_reference_golden(golden_dir, oracle_sd, monkeypatch):
    """DiffusionTrainer.validation_step (train.py:128-139) on a ragged map (l = 773 -> 8 x 96): the four logged values
    against what the unmodified reference logged, with the reference's draws injected (its CPU generator stream cannot be
    reproduced by a CUDA generator)."""
    from osu_dreamer_b200.denoiser import default_args
    from osu_dreamer_b200.trainer import DiffusionTrainer, LRScheduleArgs
    g = np.load(os.path.join(golden_dir, 'val_l773.npz'))
    l = int(g['l'])
    inp = O.make_inputs(1, l, seed=int(g['seed']))
    tr = DiffusionTrainer(val_batches=8, opt_args=dict(lr=3e-4, weight_decay=0.01),
                          schedule_args=LRScheduleArgs(warmup_steps=1000, warmup_init=0.3, decay_start=30000),
                          osl_weight=1.0, del_weight=30.0, emb_dim=6, a_dim=128, style_dim=32, diffusion_args=default_args())
    tr.diffusion_ema.module.load_state_dict(oracle_sd)
    tr = tr.cuda()
    ppr = torch.from_numpy(g['perm_plus_rand']).float()
    perm = ppr.floor()
    monkeypatch.setattr(torch, 'randperm', lambda n, **k: perm.to(k.get('device', 'cpu')))
    monkeypatch.setattr(torch, 'rand', lambda n, **k: (ppr - perm).to(k.get('device', 'cpu')))
    monkeypatch.setattr(torch, 'randn_like', lambda x, **k: torch.from_numpy(g['x0']).to(x.device))
    out = tr.validation_step((inp['h'].cuda(), inp['x1'].cuda(), inp['s'].cuda(), torch.zeros(1, 5).cuda()))
    torch.cuda.synchronize()
    ref = dict(zip([str(k) for k in g['keys']], g['vals']))
    assert set(out) == set(ref)
    for k, v in ref.items():
        e = abs(float(out[k]) - v) / abs(v)
        print(f'{k}: {float(out[k]):.6f} vs reference {v:.6f} (rel {e:.2e})')
        assert e < BF16_TOL


# ------------------------------------------------------------------------------------------------ _pred honours (a, cg)
def test_pred_uses_the_conditioning_it_is_given(oracle_sd):
    """model.py:86-103: `_pred(a, cg, xt)` is a function of the tensors it receives -- a caller may edit cg (guidance,
    interpolation) or move / slice `a` between _precompute_conditioning and _pred"""
    inp = _cuda(O.make_inputs(2, 320, seed=61))
    sdc = _cuda(oracle_sd)
    m = _model(oracle_sd)
    with torch.no_grad():
        a, cg = m._precompute_conditioning(inp['h'], inp['s'])
        ar, cgr = O.precompute_conditioning(sdc, inp['h'], inp['s'])
        assert _maxnorm(a, ar) < BF16_TOL and _maxnorm(cg, cgr) < 1e-5
        # (i) untouched
        u, v = m._pred(a, cg, inp['x0'])
        ur, vr = O.pred(sdc, ar, cgr, inp['x0'])
        assert _maxnorm(u, ur) < BF16_TOL and _maxnorm(v, vr) < BF16_TOL
        # (ii) edited global conditioning (blend of the two samples' vectors) and a re-made `a` (clone drops attributes)
        cg2 = (0.3 * cg + 0.7 * cg.flip(0)).contiguous()
        a2 = (a.clone() * 0.5)
        u2, v2 = m._pred(a2, cg2, inp['x0'])
        ur2, vr2 = O.pred(sdc, ar * 0.5, 0.3 * cgr + 0.7 * cgr.flip(0), inp['x0'])
        assert _maxnorm(u2, ur2) < BF16_TOL and _maxnorm(v2, vr2) < BF16_TOL
        assert _maxnorm(v2, vr) > 5 * BF16_TOL  # and it really is a different function value


# ------------------------------------------------------------------------------------------------ run-time depth
@pytest.mark.parametrize('depth', [1, 3, 11])
def test_other_backbone_depths_match_oracle(depth):
    """BackboneArgs.depth (backbone.py:19) is a run-time argument of the C ABI (OSD_MODE): forward in both precisions,
    all 20 + 18 * depth parameter gradients and an 8-step sampler against the oracle built with the same depth"""
    from osu_dreamer_b200.denoiser import BackboneArgs, DiffusionModel, DiffusionModelArgs
    hp = dict(O.HP, depth=depth)
    sd = O.make_state_dict(500 + depth, hp=hp)
    assert len(sd) == 20 + 18 * depth
    sdc = _cuda(sd)

    def model(precision='bf16', train=False):
        m = DiffusionModel(6, 128, 32, DiffusionModelArgs(512, 512, BackboneArgs(depth=depth, expand=4, head_dim=64,
                                                                               n_heads=16, radius=2)))
        m.load_state_dict(sd)
        m = m.cuda()
        m.precision = precision
        return m.train() if train else m.eval()

    inp = _cuda(O.make_inputs(2, 384, seed=depth))
    xt = torch.lerp(inp['x0'], inp['x1'], inp['t'][:, None, None])
    with torch.no_grad():
        ur, vr = O.forward(sdc, inp['h'], inp['s'], xt, hp=hp)
        for precision, tol in (('bf16', BF16_TOL), ('fp32', FP32_TOL)):
            u, v = model(precision)(inp['h'], inp['s'], xt)
            assert _maxnorm(u, ur) < tol and _maxnorm(v, vr) < tol, (precision, _maxnorm(u, ur), _maxnorm(v, vr))
        x_init = torch.randn(2, 6, 384, generator=torch.Generator().manual_seed(depth)).cuda()
        xr, _, _ = O.sample(sdc, inp['h'], inp['s'], x_init, 8, hp=hp)
        assert _maxnorm(model('fp32').sample_from(inp['h'], inp['s'], x_init, 8), xr) < FP32_TOL
    sdg = {k: v.detach().clone().requires_grad_(True) for k, v in sdc.items()}
    loss_r, _ = O.trainer_loss(sdg, inp['h'], inp['x1'], inp['s'], inp['x0'], inp['t'], hp=hp)
    loss_r.backward()
    m = model(train=True)
    u_pred, v_pred = m(inp['h'], inp['s'], xt)
    d_sq = O.frame_dist_sq(xt, inp['x1'])
    osl = (O.frame_dist_sq(xt - u_pred[:, None, None] * v_pred, inp['x1']) / (d_sq + m.c0)).mean()
    del_ = O.frame_dist_sq(v_pred, (xt - inp['x1']) / (d_sq + m.c0).sqrt()[:, None, None]).mean()
    (osl + 30.0 * del_).backward()
    errs = _grad_errors({n: p.grad for n, p in m.named_parameters()}, {k: v.grad for k, v in sdg.items()})
    print(f'depth {depth}: worst gradient rel-L2 errors {[(round(e, 4), n) for e, n in errs[:4]]}')
    assert errs[0][0] < BF16_TOL


def test_precision_switch_on_a_live_model(oracle_sd):
    """`precision` may be flipped on a model that has already run (bench.py, LDM): the packed operand copies and the
    workspaces are sized per mode (a compute-sanitizer memcheck finding of round 2: the fp32-grade pack used to overflow the
    bf16-sized buffer)"""
    inp = _cuda(O.make_inputs(2, 256, seed=5))
    with torch.no_grad():
        ur, vr = O.forward(_cuda(oracle_sd), inp['h'], inp['s'], inp['x0'])
        m = _model(oracle_sd, 'bf16')
        outs = []
        for prec, tol in (('bf16', BF16_TOL), ('fp32', FP32_TOL), ('bf16', BF16_TOL), ('fp32', FP32_TOL)):
            m.precision = prec
            u, v = m(inp['h'], inp['s'], inp['x0'])
            torch.cuda.synchronize()
            assert _maxnorm(u, ur) < tol and _maxnorm(v, vr) < tol, (prec, _maxnorm(v, vr))
            outs.append(v.clone())
    assert torch.equal(outs[0], outs[2]) and torch.equal(outs[1], outs[3])
