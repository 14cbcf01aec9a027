"""CPU: host-side mirror of the reference interface (no kernels run)."""
import copy
import os

import numpy as np
import pytest
import torch
import yaml

from oracle import denoiser_oracle as O
from oracle import refimport


def _new():
    from osu_dreamer_b200.denoiser import DiffusionModel, default_args
    return DiffusionModel(6, 128, 32, default_args())


def test_state_dict_contract(golden_dir):
    g = np.load(os.path.join(golden_dir, 'constants.npz'))
    m = _new()
    sd = m.state_dict()
    assert list(sd.keys()) == list(g['names'])
    assert [str(tuple(v.shape)) for v in sd.values()] == list(g['shapes'])
    assert sum(p.numel() for p in m.parameters()) == int(g['n_params'])
    assert abs(m.c0 - float(g['c0'])) < 1e-12 and abs(m.u_scale - float(g['u_scale'])) < 1e-12
    assert len(list(m.buffers())) == 0


def test_reference_init_facts():
    """zero-initialised tensors (backbone.py:12-16, model.py:51-53,66-68) and u_out.bias (model.py:71)."""
    m = _new()
    sd = m.state_dict()
    zero = [k for k, v in sd.items() if float(v.abs().max()) == 0.0]
    assert len(zero) == 37
    assert all(any(t in k for t in ('ssg1', 'ssg2', 'proj_out', 'u_mod', 'u_out.weight')) for k in zero)
    assert abs(float(sd['u_out.bias']) + 0.4328) < 1e-6
    assert float(sd['net.layers.0.attn.q_norm.weight'].min()) == 1.0
    w = sd['net.layers.2.attn.qkv_proj.weight']
    assert float(w.abs().max()) <= 512 ** -0.5 + 1e-6 and float(w.std()) > 0.02


def test_deepcopy_and_pickle_free_runtime():
    m = _new()
    m2 = copy.deepcopy(m)
    assert m2._rt is not m._rt and m2._rt.parr is None
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))


def test_args_validation():
    from osu_dreamer_b200.denoiser import DiffusionModel, DiffusionModelArgs, BackboneArgs
    with pytest.raises(ValueError):  # the widths are compile-time constants of the kernels ...
        DiffusionModel(6, 128, 32, DiffusionModelArgs(512, 256, BackboneArgs(depth=8, expand=4, head_dim=64, n_heads=16, radius=2)))
    with pytest.raises(ValueError):
        DiffusionModel(6, 128, 32, DiffusionModelArgs(512, 512, BackboneArgs(depth=8, expand=4, head_dim=64, n_heads=8, radius=2)))
    # ... the depth is a run-time argument of the C ABI: 20 + 18 * depth tensors in the reference's registration order
    m4 = DiffusionModel(6, 128, 32, DiffusionModelArgs(512, 512, BackboneArgs(depth=4, expand=4, head_dim=64, n_heads=16, radius=2)))
    assert len(m4.state_dict()) == 20 + 18 * 4 and 'net.layers.3.ffn.proj_o.bias' in m4.state_dict()
    assert 'net.layers.4.ssg1.weight' not in m4.state_dict() and m4._mode() == (4 << 8)
    # dict-shaped args as produced by models/inference/artifact.py:52-71 dataclass_from_dict on older ckpts
    DiffusionModel(6, 128, 32, DiffusionModelArgs(512, 512, dict(depth=8, expand=4, head_dim=64, n_heads=16, radius=2)))


def test_lr_schedule_matches_golden(golden_dir):
    from osu_dreamer_b200.trainer import LRScheduleArgs, make_lr_schedule
    g = np.load(os.path.join(golden_dir, 'constants.npz'))
    f = make_lr_schedule(LRScheduleArgs(warmup_steps=1000, warmup_init=0.3, decay_start=30000))
    for s, v in zip(g['lr_steps'], g['lr_vals']):
        assert abs(f(int(s)) - float(v)) < 1e-12


def test_trainer_from_reference_yaml_schema():
    from osu_dreamer_b200.cli import build_trainer
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = yaml.safe_load(open(os.path.join(root, 'osu-dreamer_b200', 'denoiser.yml')))
    tr = build_trainer(cfg)
    keys = list(tr.state_dict().keys())
    assert len(keys) == 329 and 'diffusion_ema.n_averaged' in keys
    assert sum(k.startswith('diffusion_ema.module.') for k in keys) == 164 and keys[0] == 'diffusion.proj_audio.0.weight'
    assert tr.gradient_clip_val == 1.0 and abs(tr.current_lr() - 3e-4 * 0.3) < 1e-12


def test_cli_commands_fail_loudly_without_cuda():
    """fit-denoiser / predict are the reference's two commands for this path (osu_dreamer/__main__.py:19-29); without
    a CUDA device both must refuse (no CPU fallback) with a clean click error, not a traceback"""
    from click.testing import CliRunner
    from osu_dreamer_b200.cli import main
    assert {'fit-denoiser', 'predict', 'export-inference'} <= set(main.commands)
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = CliRunner().invoke(main, ['fit-denoiser', '-c', os.path.join(root, 'osu-dreamer_b200', 'denoiser.yml'), '--synthetic'])
    assert r.exit_code != 0 and 'no CPU fallback' in r.output
    r = CliRunner().invoke(main, ['predict', '--model-path', 'x.pt', '--audio-file', 'a.mp3'])
    assert r.exit_code != 0 and 'no CPU fallback' in r.output


@pytest.mark.skipif(not refimport.available(), reason='reference checkout not present on this host')
def test_reference_checkpoint_interchange():
    """a reference state dict loads into the mirror and back (strict), and the reference YAML parses."""
    ns = refimport.import_reference()
    ref = refimport.build_reference_model(ns, O.make_state_dict(1234))
    m = _new()
    m.load_state_dict(ref.state_dict(), strict=True)
    ref.load_state_dict(m.state_dict(), strict=True)
    from osu_dreamer_b200.cli import build_trainer
    cfg = yaml.safe_load(open(os.path.join(refimport.REFERENCE_ROOT, 'osu_dreamer/models/diffusion/model.yml')))
    assert build_trainer(cfg).val_batches == 8


@pytest.mark.skipif(not refimport.available(), reason='reference checkout not present on this host')
def test_checkpoint_feeds_reference_export_inference(tmp_path):
    """a checkpoint written by this package goes through the REFERENCE's save_inference, and through this package's
    export-inference, to the same artifact; the reference rebuilds its DiffusionModel from the artifact's hparams and
    loads the EMA weights strictly (models/inference/artifact.py:9-49)"""
    import sys
    from osu_dreamer_b200.cli import build_trainer, load_checkpoint, save_checkpoint, save_inference
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = yaml.safe_load(open(os.path.join(root, 'osu-dreamer_b200', 'denoiser.yml')))
    tr = build_trainer(cfg)
    tr.diffusion.load_state_dict(O.make_state_dict(1234))
    tr.diffusion_ema.module.load_state_dict(O.make_state_dict(77))
    tr.global_step = 123
    save_checkpoint(str(tmp_path / 'denoiser.ckpt'), tr, epoch=4)
    ck = torch.load(tmp_path / 'denoiser.ckpt', weights_only=True)  # plain tensors / containers only: no pickled classes
    assert ck['global_step'] == 123 and isinstance(ck['hyper_parameters']['diffusion_args'], dict)
    latent_hp = dict(emb_dim=6, style_dim=32, n_downs=3, stride=3, latent_args=dict(h_dim=128))
    torch.save({'hyper_parameters': latent_hp, 'state_dict': {'latent.proj_emb.weight': torch.zeros(128, 6, 1), 'other': torch.zeros(1)}},
               tmp_path / 'latent.ckpt')
    torch.save({'hyper_parameters': dict(style_args=dict(h_dim=256)), 'state_dict': {'style_ema.module.u_out.bias': torch.zeros(1),
                                                                                     'style.u_out.bias': torch.ones(1)}},
               tmp_path / 'style.ckpt')
    ns = refimport.import_reference()
    refimport._install_stubs()
    from osu_dreamer.models.inference.artifact import save_inference as ref_save, dataclass_from_dict
    paths = [str(tmp_path / f) for f in ('latent.ckpt', 'denoiser.ckpt', 'style.ckpt')]
    ref_save(*paths, str(tmp_path / 'ref.pt'))
    save_inference(*paths, str(tmp_path / 'ours.pt'))
    a, b = torch.load(tmp_path / 'ref.pt', weights_only=True), torch.load(tmp_path / 'ours.pt', weights_only=True)
    assert a['hparams'] == b['hparams'] and a['state_dict'].keys() == b['state_dict'].keys()
    assert all(torch.equal(a['state_dict'][k], b['state_dict'][k]) for k in a['state_dict'])
    assert sum(k.startswith('diffusion.') for k in a['state_dict']) == 164 and 'style.u_out.bias' in a['state_dict']
    args = dataclass_from_dict(ns.DiffusionModelArgs, a['hparams']['diffusion_args'])
    ref = ns.DiffusionModel(a['hparams']['emb_dim'], a['hparams']['latent_args']['h_dim'], a['hparams']['style_dim'], args)
    ref.load_state_dict({k[len('diffusion.'):]: v for k, v in a['state_dict'].items() if k.startswith('diffusion.')}, strict=True)
    assert torch.equal(ref.proj_in.weight, O.make_state_dict(77)['proj_in.weight'])  # the EMA copy, not the live weights
    # resume round trip
    tr2 = build_trainer(cfg)
    load_checkpoint(str(tmp_path / 'denoiser.ckpt'), tr2)
    assert tr2.global_step == 123 and torch.equal(tr2.diffusion.proj_in.weight, tr.diffusion.proj_in.weight)


def test_flat_buffers_are_aligned():
    from osu_dreamer_b200.trainer import _flatten, _is_flat
    m = _new()
    ps = list(m.parameters())
    before = [p.detach().clone() for p in ps]
    flat = _flatten(ps)
    assert _is_flat(ps, flat) and all(p.data_ptr() % 256 == flat.data_ptr() % 256 for p in ps)
    assert all(torch.equal(a, b) for a, b in zip(before, ps))
    assert flat.numel() >= sum(p.numel() for p in ps)


def test_latent_cache_reader(tmp_path):
    from osu_dreamer_b200.data import LatentWindows, batches, split_mapsets
    rng = np.random.default_rng(0)
    for ms in range(4):
        d = tmp_path / f'set{ms}'
        d.mkdir()
        l = 400 + 37 * ms
        np.save(d / 'h.npy', rng.standard_normal((128, l)).astype(np.float32))
        for k in range(2):
            np.savez(d / f'm{k}.latent.npz', z=rng.standard_normal((6, l)).astype(np.float32),
                     s=rng.standard_normal(32).astype(np.float32), labels=rng.random(5).astype(np.float32))
    train, val = split_mapsets(tmp_path, '*.latent.npz', 2, 0.3)
    assert len(train) == 3 and len(val) == 1
    bs = list(batches(LatentWindows(train, 152, shuffle_buffer_size=4, max_per_map=2, seed=1), 4, pin=False))
    assert len(bs) >= 2 and bs[0][0].shape == (4, 128, 152) and bs[0][1].shape == (4, 6, 152) and bs[0][2].shape == (4, 32)
    full = list(LatentWindows(val, None))
    assert full[0].z.shape[0] == 6 and full[0].h.shape[1] == full[0].z.shape[1]


def _make_cache(root, n_sets=5, maps_per_set=3, seed=0, h_dtype=np.float32):
    rng = np.random.default_rng(seed)
    for ms in range(n_sets):
        d = root / f'set{ms}'
        d.mkdir()
        l = 700 + 53 * ms
        np.save(d / 'h.npy', rng.standard_normal((128, l)).astype(h_dtype))
        for k in range(maps_per_set):
            np.savez(d / f'm{k}.latent.npz', z=rng.standard_normal((6, l)).astype(np.float32),
                     s=rng.standard_normal(32).astype(np.float32), labels=rng.random(5).astype(np.float32))
    return sorted(p for p in root.iterdir() if p.is_dir())


@pytest.mark.skipif(not refimport.available(), reason='reference checkout not present on this host')
@pytest.mark.parametrize('seq_len,buf,mpm', [(160, 1, -1), (160, 7, 2), (300, 16, -1), (None, 1, -1)])
def test_latent_windows_match_reference_stream(tmp_path, seq_len, buf, mpm):
    """same cache, same global seeds -> the reader yields the reference LatentDataset's samples, in its order
    (osu_dreamer/data/modules/latent.py:86-149, single-process path)"""
    import sys
    refimport._install_stubs()
    if refimport.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, refimport.REFERENCE_ROOT)
    from osu_dreamer.data.modules.latent import LatentDataset
    from osu_dreamer_b200.data import LatentWindows
    sets = _make_cache(tmp_path)
    torch.manual_seed(123)
    ref = list(LatentDataset(sets, seq_len, buf, mpm))
    torch.manual_seed(123)
    ours = list(LatentWindows(sets, seq_len, buf, mpm, rng='global'))
    assert len(ref) == len(ours) and len(ref) > 0
    for a, b in zip(ref, ours):
        for x, y in zip(a, b):
            assert x.dtype == y.dtype and torch.equal(x, y)


def test_prefetcher_order_errors_and_pinned_batches(tmp_path):
    from osu_dreamer_b200.data import LatentWindows, Prefetcher, batches
    sets = _make_cache(tmp_path, h_dtype=np.float16)  # a half-precision cache is converted on the way out
    direct = list(batches(LatentWindows(sets, 128, 4, -1, seed=3), 4, pin=False))
    pre = list(Prefetcher(batches(LatentWindows(sets, 128, 4, -1, seed=3), 4, pin=False), depth=2))
    assert len(direct) == len(pre) > 2 and direct[0][0].dtype == torch.float32
    for a, b in zip(direct, pre):
        assert all(torch.equal(x, y) for x, y in zip(a, b))
    # rank sharding: rank r takes batch k*world + r; every rank takes the same number of steps (a trailing
    # incomplete round is dropped, otherwise the gradient all-reduce of the last step would hang)
    for world in (2, 3):
        per_rank = [list(batches(LatentWindows(sets, 128, 4, -1, seed=3), 4, rank=r, world=world, pin=False)) for r in range(world)]
        assert all(len(p) == len(direct) // world for p in per_rank)
        for r, p in enumerate(per_rank):
            for k, b in enumerate(p):
                assert torch.equal(b[0], direct[k * world + r][0]) and torch.equal(b[2], direct[k * world + r][2])
    # the feeder (host mode without a GPU: recycled buffers, background thread) yields the same batches
    from osu_dreamer_b200.data import DeviceFeeder
    fed = list(DeviceFeeder(LatentWindows(sets, 128, 4, -1, seed=3), 4, device=None, depth=2))
    assert len(fed) == len(direct)
    for a, b in zip(direct, fed):
        assert all(torch.equal(x, y) for x, y in zip(a, b))

    def boom():
        yield direct[0]
        raise RuntimeError('loader failed')
    p = Prefetcher(boom())
    next(p)
    with pytest.raises(RuntimeError, match='loader failed'):
        next(p)
    p.close()


def test_style_mirror_state_dict_contract_and_no_cpu_path(golden_dir):
    """the mirror of StyleModel carries the reference's parameter / buffer names, shapes and order
    (tests/golden/nb_spec.json, written from the reference module) and refuses to run without CUDA"""
    import json
    from osu_dreamer_b200 import lib
    from osu_dreamer_b200.style import StyleModel, StyleModelArgs
    spec = json.load(open(os.path.join(golden_dir, 'nb_spec.json')))['style']
    m = StyleModel(32, StyleModelArgs(label_features=128, h_dim=256, depth=8, expand=4))
    sd = m.state_dict()
    assert [(k, list(v.shape)) for k, v in sd.items()] == [(k, list(s)) for k, s in spec] and len(sd) == lib.STYLE_NUM_PARAMS
    assert abs(m.c0 - 64 * (1 - torch.tensor(2.3263478740408408).sigmoid().item()) ** 2) < 1e-12 and m.u_scale == 8.0
    assert float(m.u_out.bias.detach()) == pytest.approx(-0.4328) and float(m.proj_out[1].weight.detach().abs().max()) == 0.0
    with pytest.raises(lib.OsdError):
        StyleModel(16, StyleModelArgs(128, 256, 8, 4))
    if not torch.cuda.is_available():
        with torch.no_grad(), pytest.raises(lib.OsdError, match='no CPU path'):
            m(torch.randn(2, 32), torch.rand(2, 5))
        with pytest.raises(lib.OsdError, match='no CPU path'):  # the training path (autograd on) has no CPU fallback either
            m(torch.randn(2, 32), torch.rand(2, 5))
    with pytest.raises(lib.OsdError, match='not produced'):  # st / labels are data in fit-style
        m(torch.randn(2, 32, requires_grad=True), torch.rand(2, 5))


def test_latent_mirror_state_dict_contract(golden_dir):
    """the inference mirror of LatentModel carries the reference's full parameter tree -- names, shapes, order
    (tests/golden/nb_spec.json, written from the reference module) -- deep-copies, and refuses to run without CUDA"""
    import copy
    import json
    from osu_dreamer_b200 import lib
    from osu_dreamer_b200.latent import LatentModel, LatentModelArgs, LayerArgs
    spec = json.load(open(os.path.join(golden_dir, 'nb_spec.json')))['latent']
    m = LatentModel(6, 32, 3, 3, LatentModelArgs(128, LayerArgs(8, 4, 2), 64, 16))
    assert [(k, list(v.shape)) for k, v in m.state_dict().items()] == [(k, list(s)) for k, s in spec]
    assert m.chunk_size == 27 and m.a_dim == 128
    assert float(getattr(m.decoder.mixers, '0').gate.weight.abs().max()) == 0.0  # zero-initialised like the reference (unet.py:119)
    assert float(getattr(getattr(m.decoder.layers, '0').blocks, '0')._modules['1'].gamma[0]) == pytest.approx(1e-3)
    m2 = copy.deepcopy(m)
    assert m2.audio_encoder._owner[0] is m2
    # dict-form arguments, as rebuilt from an inference artifact's hparams (models/inference/artifact.py:52-71)
    LatentModel(6, 32, 3, 3, dict(h_dim=128, ae_args=dict(n_layers=8, expand=4, radius=2), style_head_dim=64, style_heads=16))
    if not torch.cuda.is_available():
        with pytest.raises(lib.OsdError, match='no CPU path'):
            m.audio_encoder(torch.randn(1, 72, 27))


@pytest.mark.skipif(not refimport.available(), reason='reference checkout not present on this host')
def test_install_swaps_every_model_of_ldm_and_artifacts_load(tmp_path):
    """after cli.install() the REFERENCE's own LDM (models/inference/model.py:27-32) is built from the three B200 mirrors,
    and an artifact written from reference-initialised weights loads into it strictly through the reference's
    load_inference (models/inference/artifact.py:44-49) -- the wiring `predict` relies on (no kernels run here)"""
    import sys
    refimport._install_stubs()
    if refimport.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, refimport.REFERENCE_ROOT)
    import importlib
    inf = importlib.import_module('osu_dreamer.models.inference.model')
    art = importlib.import_module('osu_dreamer.models.inference.artifact')
    dm = importlib.import_module('osu_dreamer.models.diffusion.model')
    bb = importlib.import_module('osu_dreamer.models.diffusion.backbone')
    saved = (inf.LatentModel, inf.StyleModel, inf.DiffusionModel, dm.DiffusionModel, dm.DiffusionModelArgs, bb.BackboneArgs)
    hparams = dict(emb_dim=6, style_dim=32, n_downs=3, stride=3,
                   latent_args=dict(h_dim=128, ae_args=dict(n_layers=8, expand=4, radius=2), style_head_dim=64, style_heads=16),
                   style_args=dict(label_features=128, h_dim=256, depth=8, expand=4),
                   diffusion_args=dict(global_cond_dim=512, backbone_dim=512, u_head_dim=64,
                                       backbone_args=dict(depth=8, expand=4, head_dim=64, n_heads=16, radius=2)))
    try:
        ref_ldm = inf.LDM(art.dataclass_from_dict(inf.LDMArgs, hparams))  # the unmodified reference, reference init
        torch.save({'hparams': hparams, 'state_dict': ref_ldm.state_dict()}, tmp_path / 'inference.pt')
        from osu_dreamer_b200 import cli, denoiser, latent, style
        cli.install()
        ours = art.load_inference(tmp_path / 'inference.pt')  # strict load_state_dict inside
        assert isinstance(ours.latent, latent.LatentModel) and isinstance(ours.style, style.StyleModel)
        assert isinstance(ours.diffusion, denoiser.DiffusionModel)
        a, b = ref_ldm.state_dict(), ours.state_dict()
        assert list(a.keys()) == list(b.keys()) and all(torch.equal(a[k], b[k]) for k in a)
        assert ours.latent.chunk_size == ref_ldm.latent.chunk_size == 27 and callable(ours.latent.audio_encoder)
        assert abs(ours.style.c0 - ref_ldm.style.c0) < 1e-12 and abs(ours.diffusion.c0 - ref_ldm.diffusion.c0) < 1e-12
        # the package's own pipeline class reads the same artifact without the reference
        from osu_dreamer_b200.ldm import load_artifact, pad_to_multiple
        own = load_artifact(tmp_path / 'inference.pt', device='cpu')
        c = own.state_dict()
        assert list(a.keys()) == list(c.keys()) and all(torch.equal(a[k], c[k]) for k in a)
        from osu_dreamer.data.modules.beatmap import pad_to_multiple as ref_pad
        x = torch.randn(72, 100)
        assert torch.equal(pad_to_multiple(x, 27), ref_pad(x, 27)) and pad_to_multiple(x, 27).shape[-1] == 108
        assert pad_to_multiple(x[:, :54], 27) is not None and pad_to_multiple(x[:, :54], 27).shape[-1] == 54
    finally:  # install() patches the imported reference modules: put the originals back for the other tests
        inf.LatentModel, inf.StyleModel, inf.DiffusionModel, dm.DiffusionModel, dm.DiffusionModelArgs, bb.BackboneArgs = saved


def test_rank_sharding_properties():
    """data.batch_refs over arbitrary stream lengths / batch sizes / world sizes: every rank takes the same number of
    steps, rank r's k-th batch is global batch k*world + r, nothing is duplicated, only a trailing incomplete round (and a
    trailing partial batch) is dropped"""
    from hypothesis import given, settings, strategies as st
    from osu_dreamer_b200.data import batch_refs

    @settings(max_examples=60, deadline=None)
    @given(n=st.integers(0, 200), bs=st.integers(1, 9), world=st.integers(1, 5))
    def check(n, bs, world):
        per_rank = [list(batch_refs(list(range(n)), bs, r, world)) for r in range(world)]
        full = n // bs
        rounds = full // world
        assert all(len(p) == rounds for p in per_rank)
        for r, p in enumerate(per_rank):
            for k, grp in enumerate(p):
                g = k * world + r
                assert grp == list(range(g * bs, (g + 1) * bs))
    check()


def test_style_trainer_from_reference_yaml_schema():
    """fit-style: the reference's own models/style/model.yml schema builds the mirror (no schedule_args there: constant LR);
    state-dict keys are the reference's (style.*, style_ema.module.*, style_ema.n_averaged)"""
    from osu_dreamer_b200.cli import build_style_trainer, check_config
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = yaml.safe_load(open(os.path.join(root, 'osu-dreamer_b200', 'style.yml')))
    assert check_config(cfg) == []
    cfg['model'].pop('schedule_args')
    tr = build_style_trainer(cfg)
    keys = list(tr.state_dict().keys())
    assert len(keys) == 121 and keys[0] == 'style.cond_proj_w' and 'style_ema.n_averaged' in keys
    assert sum(k.startswith('style_ema.module.') for k in keys) == 60
    assert tr.gradient_clip_val == 1.0 and tr.current_lr() == 3e-4 and tr.label_drop_prob == .2
    assert sum(p.numel() for p in tr.style.parameters()) == 5970721 or sum(p.numel() for p in tr.style.parameters()) > 5.9e6


def test_check_config_refuses_what_it_cannot_honour():
    import click
    from osu_dreamer_b200.cli import check_config
    base = {'trainer': {'gradient_clip_val': 1.0}, 'model': {'opt_args': {'lr': 1e-3}}}
    assert check_config(base) == []
    with pytest.raises(click.ClickException):
        check_config({'trainer': {'accumulate_grad_batches': 4}, 'model': {}})
    with pytest.raises(click.ClickException):
        check_config({'trainer': {'precision': '32-true'}, 'model': {}})
    with pytest.raises(click.ClickException):
        check_config({'trainer': {}, 'model': {'opt_args': {'amsgrad': True}}})
    assert any('val_check_interval' in w for w in check_config({'trainer': {'val_check_interval': 100}, 'model': {}}))
