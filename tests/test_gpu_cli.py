"""GPU: the product path end to end through the CLI -- cached-latent reader -> DeviceFeeder -> training steps ->
validation on the EMA weights -> checkpoint -> resume -> export-inference -> sampling from the exported weights
(reference flow: scripts/fit_denoiser.py:17-32, models/diffusion/train.py:120-139, scripts/export_inference.py:6-13,
models/inference/artifact.py:44-49)."""
import numpy as np
import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu


def _make_cache(root, n_sets=8, maps_per_set=3):
    rng = np.random.default_rng(0)
    for ms in range(n_sets):
        d = root / f'set{ms}'
        d.mkdir(parents=True)
        l = 600 + 37 * ms
        np.save(d / 'h.npy', rng.standard_normal((128, l)).astype(np.float32))
        for k in range(maps_per_set):
            z = rng.standard_normal((6, l)).astype(np.float32)
            z /= np.sqrt((z ** 2).mean(0, keepdims=True) + 1e-6)
            s = rng.standard_normal(32).astype(np.float32)
            s /= np.sqrt((s ** 2).mean() + 1e-6)
            np.savez(d / f'm{k}.latent.npz', z=z, s=s, labels=(10 * rng.random(5)).astype(np.float32))


def test_fit_resume_export_sample(tmp_path):
    import os
    from click.testing import CliRunner
    from osu_dreamer_b200.cli import main
    from osu_dreamer_b200.denoiser import DiffusionModel, DiffusionModelArgs, BackboneArgs
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = yaml.safe_load(open(os.path.join(root, 'osu-dreamer_b200', 'denoiser.yml')))
    _make_cache(tmp_path / 'data')
    cfg['data'].update(seq_len=128, batch_size=4, max_val_count=4, shuffle_buffer_size=8, max_per_map=2,
                       data_path=str(tmp_path / 'data'))
    cfg['trainer']['log_every_n_steps'] = 1
    cfg_path = tmp_path / 'cfg.yml'
    yaml.safe_dump(cfg, open(cfg_path, 'w'))
    ck1, ck2 = str(tmp_path / 'a.ckpt'), str(tmp_path / 'b.ckpt')
    r = CliRunner().invoke(main, ['fit-denoiser', '-c', str(cfg_path), '--max-steps', '4', '--out', ck1], catch_exceptions=False)
    assert r.exit_code == 0, r.output
    assert 'step 4 ' in r.output and 'val/loss' in r.output
    a = torch.load(ck1, weights_only=True)
    assert a['global_step'] == 4 and len(a['state_dict']) == 329
    assert all(torch.isfinite(v).all() for v in a['state_dict'].values())
    r = CliRunner().invoke(main, ['fit-denoiser', '-c', str(cfg_path), '--ckpt-path', ck1, '--max-steps', '7', '--out', ck2],
                           catch_exceptions=False)
    assert r.exit_code == 0, r.output
    b = torch.load(ck2, weights_only=True)
    assert b['global_step'] == 7 and 'step 5 ' in r.output and 'step 4 ' not in r.output
    w0, w1 = a['state_dict']['diffusion.net.layers.3.attn.qkv_proj.weight'], b['state_dict']['diffusion.net.layers.3.attn.qkv_proj.weight']
    assert not torch.equal(w0, w1)  # training moved the weights
    e1 = b['state_dict']['diffusion_ema.module.net.layers.3.attn.qkv_proj.weight']
    assert not torch.equal(e1, w1) and float((e1 - w1).abs().max()) < float((w1 - w0).abs().max()) * 50 + 1e-3
    assert int(b['state_dict']['diffusion_ema.n_averaged']) == 7
    # export + sample from the artifact's EMA weights
    torch.save({'hyper_parameters': dict(emb_dim=6, style_dim=32, n_downs=3, stride=3, latent_args=dict(h_dim=128)), 'state_dict': {}},
               tmp_path / 'latent.ckpt')
    torch.save({'hyper_parameters': dict(style_args={}), 'state_dict': {}}, tmp_path / 'style.ckpt')
    out = str(tmp_path / 'inference.pt')
    r = CliRunner().invoke(main, ['export-inference', '--latent-ckpt-path', str(tmp_path / 'latent.ckpt'), '--denoiser-ckpt-path', ck2,
                                  '--style-ckpt-path', str(tmp_path / 'style.ckpt'), '--output-path', out], catch_exceptions=False)
    assert r.exit_code == 0, r.output
    art = torch.load(out, weights_only=True)
    da = dict(art['hparams']['diffusion_args'])
    da['backbone_args'] = BackboneArgs(**da['backbone_args'])
    m = DiffusionModel(art['hparams']['emb_dim'], art['hparams']['latent_args']['h_dim'], art['hparams']['style_dim'],
                       DiffusionModelArgs(**da))
    m.load_state_dict({k[len('diffusion.'):]: v for k, v in art['state_dict'].items() if k.startswith('diffusion.')}, strict=True)
    m = m.cuda().eval()
    x = m.sample(torch.randn(1, 128, 200, device='cuda'), torch.randn(3, 32, device='cuda'), 4)
    assert x.shape == (3, 6, 200) and torch.isfinite(x).all()


def test_fit_style_resume_export(tmp_path):
    """fit-style through the CLI on a cached dataset (reference flow: scripts/fit_style.py:17-31, models/style/train.py):
    training steps, the whole-set validation (val/energy_dist selects the kept checkpoint), resume, and the exported
    artifact's style.* weights loading strictly into the style mirror and sampling"""
    import os
    from click.testing import CliRunner
    from osu_dreamer_b200.cli import main
    from osu_dreamer_b200.style import StyleModel, StyleModelArgs
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = yaml.safe_load(open(os.path.join(root, 'osu-dreamer_b200', 'style.yml')))
    _make_cache(tmp_path / 'data', n_sets=12, maps_per_set=4)
    cfg['data'].update(batch_size=8, max_val_count=6, shuffle_buffer_size=8, data_path=str(tmp_path / 'data'))
    cfg['trainer'].update(log_every_n_steps=1, max_epochs=2)
    cfg_path = tmp_path / 'style_cfg.yml'
    yaml.safe_dump(cfg, open(cfg_path, 'w'))
    ck1, ck2 = str(tmp_path / 's1.ckpt'), str(tmp_path / 's2.ckpt')
    r = CliRunner().invoke(main, ['fit-style', '-c', str(cfg_path), '--max-steps', '5', '--out', ck1], catch_exceptions=False)
    assert r.exit_code == 0, r.output
    assert 'step 5 ' in r.output and 'val/energy_dist' in r.output and 'train/u_mape' in r.output
    a = torch.load(ck1, weights_only=True)
    assert len(a['state_dict']) == 121 and a['optimizer_state']['step'] == a['global_step']
    assert all(torch.isfinite(v).all() for v in a['state_dict'].values())
    assert a['hyper_parameters']['style_args']['h_dim'] == 256
    r = CliRunner().invoke(main, ['fit-style', '-c', str(cfg_path), '--ckpt-path', ck1, '--max-steps', str(a['global_step'] + 3),
                                  '--out', ck2], catch_exceptions=False)
    assert r.exit_code == 0, r.output
    b = torch.load(ck2, weights_only=True) if os.path.exists(ck2) else None
    assert b is not None and b['global_step'] == a['global_step'] + 3 and b['epoch'] > a['epoch']
    assert not torch.equal(a['state_dict']['style.blocks.3.0.weight'], b['state_dict']['style.blocks.3.0.weight'])
    # the EMA weights are what export-inference ships as style.* (models/inference/artifact.py:31-35)
    sd = {k[len('style_ema.module.'):]: v for k, v in b['state_dict'].items() if k.startswith('style_ema.module.')}
    m = StyleModel(32, StyleModelArgs(**b['hyper_parameters']['style_args']))
    m.load_state_dict(sd, strict=True)
    s = m.cuda().eval().sample(10 * torch.rand(5, 5, device='cuda'), 16)
    assert s.shape == (5, 32) and torch.isfinite(s).all()
