"""`fit-denoiser` / `predict` entry points (reference: osu_dreamer/scripts/fit_denoiser.py:17-32,
osu_dreamer/scripts/predict.py:21-100, wired in osu_dreamer/__main__.py:19-29).

`fit-denoiser` reads the reference's own YAML schema (osu_dreamer/models/diffusion/model.yml: seed_everything /
trainer / data / model) and runs the training loop on the CUDA path; under torchrun it is data-parallel over
the GPUs of the box (one process per GPU, NCCL all-reduce of the gradients).  pytorch_lightning is not
required.  `predict` needs the reference's latent / style models and audio front end, which are outside this
package: `install()` swaps this package's DiffusionModel into an importable reference checkout so that
`python -m osu_dreamer predict ...` runs its `diffusion.sample` call (models/inference/model.py:50) on the
CUDA path while everything else stays the reference's.
"""
from __future__ import annotations

import os
import sys
import time
from pathlib import Path

import click
import torch
import yaml

from .data import DeviceFeeder, LatentWindows, split_mapsets, synthetic_batches
from .trainer import DiffusionTrainer


def install() -> None:
    """Make the reference (if importable) use this package's DiffusionModel / args classes."""
    from . import denoiser
    import importlib
    ref = importlib.import_module('osu_dreamer.models.diffusion.model')
    ref.DiffusionModel = denoiser.DiffusionModel
    ref.DiffusionModelArgs = denoiser.DiffusionModelArgs
    bb = importlib.import_module('osu_dreamer.models.diffusion.backbone')
    bb.BackboneArgs = denoiser.BackboneArgs
    for name in ('osu_dreamer.models.inference.model', 'osu_dreamer.models.diffusion.train'):
        if name in sys.modules:
            sys.modules[name].DiffusionModel = denoiser.DiffusionModel
    # the other three calls of LDM.sample (models/inference/model.py:47-51: latent.audio_encoder, style.sample,
    # latent.decode) have inference-only mirrors: they are swapped into the inference module, not into the trainers
    from . import latent, style
    try:
        inf = importlib.import_module('osu_dreamer.models.inference.model')
    except ImportError:  # a training-only environment without the inference module's dependencies
        inf = None
    if inf is not None:
        inf.StyleModel = style.StyleModel
        inf.LatentModel = latent.LatentModel  # inference half: audio_encoder / decode around diffusion.sample
        inf.DiffusionModel = denoiser.DiffusionModel


def build_trainer(cfg: dict) -> DiffusionTrainer:
    m = dict(cfg['model'])
    clip = (cfg.get('trainer') or {}).get('gradient_clip_val', 0.0) or 0.0
    return DiffusionTrainer(**m, gradient_clip_val=float(clip))


def _plain(x):
    """dataclasses -> dicts (recursively): the checkpoint must unpickle without this package, and the reference's
    `dataclass_from_dict` (models/inference/artifact.py:52-71) rebuilds its own args classes from dicts"""
    import dataclasses
    if dataclasses.is_dataclass(x) and not isinstance(x, type):
        return {f.name: _plain(getattr(x, f.name)) for f in dataclasses.fields(x)}
    if isinstance(x, dict):
        return {k: _plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_plain(v) for v in x)
    return x


def save_checkpoint(path: str, tr: DiffusionTrainer, epoch: int, best: float | None = None):
    """Lightning-shaped checkpoint (`state_dict` with the reference's keys diffusion.* / diffusion_ema.module.* /
    diffusion_ema.n_averaged, `hyper_parameters`, `global_step`, `epoch`): what the reference's export-inference
    reads (models/inference/artifact.py:15-35).  Plain tensors / dicts only.  `best_model_score` is what Lightning's
    ModelCheckpoint restores on resume (model.yml:15-21); `optimizer_state` carries the AdamW moments and their own
    step counter."""
    torch.save({'state_dict': {k: v.detach().cpu() for k, v in tr.state_dict().items()},
                'hyper_parameters': _plain(dict(tr.hparams)),
                'global_step': tr.global_step, 'epoch': epoch, 'best_model_score': best,
                'optimizer_state': ({'m': tr._opt['m'].cpu(), 'v': tr._opt['v'].cpu(), 'step': tr.adam_step}
                                    if tr._opt else None)}, path)


def load_checkpoint(path: str, tr: DiffusionTrainer):
    """resume: weights + EMA + the LR-schedule step; the AdamW moments and THEIR step count when the checkpoint carries
    them.  A checkpoint written by the reference's Lightning run has no `optimizer_state` in this shape: the moments then
    restart at zero and so does the bias-correction step (`tr.adam_step = 0`), exactly like a fresh torch AdamW -- with
    the bias corrections taken at `global_step` instead, the first update would be (1-b1) g / sqrt((1-b2) g^2) ~ 3.2x a
    normal Adam step at full learning rate.  Returns the checkpoint dict (`epoch`, `best_model_score` for the loop)."""
    ck = torch.load(path, map_location='cpu', weights_only=False)
    tr.load_state_dict(ck['state_dict'])
    tr.global_step = int(ck.get('global_step', 0))
    opt = ck.get('optimizer_state')
    if opt is not None:
        tr._opt = {'m': opt['m'].clone(), 'v': opt['v'].clone()}  # adopted by configure_optimizers when sizes match
        tr.adam_step = int(opt.get('step', tr.global_step))
    else:
        tr._opt = None
        tr.adam_step = 0
    return ck


_TRAINER_KEYS_HONOURED = {'gradient_clip_val', 'log_every_n_steps', 'max_epochs'}
_TRAINER_KEYS_NEUTRAL = {'accelerator', 'devices', 'logger', 'callbacks', 'enable_progress_bar', 'default_root_dir', 'strategy',
                         'num_nodes', 'enable_checkpointing', 'enable_model_summary', 'num_sanity_val_steps'}
_OPT_KEYS_HONOURED = {'lr', 'betas', 'eps', 'weight_decay'}


def check_config(cfg: dict) -> list[str]:
    """Reference config keys this loop does not implement must not pass silently: settings that would change the
    arithmetic raise, the rest are reported.  -> list of warnings"""
    warns = []
    t = cfg.get('trainer') or {}
    if int(t.get('accumulate_grad_batches', 1) or 1) != 1:
        raise click.ClickException('trainer.accumulate_grad_batches != 1 is not implemented (model.yml uses 1)')
    prec = str(t.get('precision', 'bf16-mixed'))
    if prec not in ('bf16-mixed', 'bf16', 'bf16-true'):
        raise click.ClickException(f'trainer.precision={prec!r}: training runs with bf16 tensor-core operands (fp32 accumulate, '
                                   'residual and statistics) -- the reference\'s bf16-mixed (model.yml:38)')
    for k in t:
        if k not in _TRAINER_KEYS_HONOURED | _TRAINER_KEYS_NEUTRAL | {'accumulate_grad_batches', 'precision'}:
            warns.append(f'trainer.{k} is ignored by this loop')
    for k, v in ((cfg.get('model') or {}).get('opt_args') or {}).items():
        if k not in _OPT_KEYS_HONOURED:
            if k == 'amsgrad' and not v:
                continue
            raise click.ClickException(f'model.opt_args.{k}={v!r} is not implemented by the fused AdamW (lr, betas, eps, weight_decay)')
    return warns


def save_inference(latent_ckpt_path: str, denoiser_ckpt_path: str, style_ckpt_path: str, output_path: str):
    """`export-inference` (scripts/export_inference.py:6-13, models/inference/artifact.py:9-42): one artifact
    {'hparams', 'state_dict'} with latent.* from the latent checkpoint and the EMA weights of the denoiser and the
    style model renamed to diffusion.* / style.*; the reference's `load_inference` / `predict` read it."""
    ck = {k: torch.load(p, map_location='cpu', weights_only=False)
          for k, p in (('latent', latent_ckpt_path), ('denoiser', denoiser_ckpt_path), ('style', style_ckpt_path))}
    hp = {k: ck['latent']['hyper_parameters'][k] for k in ('emb_dim', 'style_dim', 'n_downs', 'stride', 'latent_args')}
    hp['diffusion_args'] = ck['denoiser']['hyper_parameters']['diffusion_args']
    hp['style_args'] = ck['style']['hyper_parameters']['style_args']
    sd = {k: v for k, v in ck['latent']['state_dict'].items() if k.startswith('latent.')}
    for src, prefix, dst in (('denoiser', 'diffusion_ema.module.', 'diffusion.'), ('style', 'style_ema.module.', 'style.')):
        sd.update({dst + k[len(prefix):]: v for k, v in ck[src]['state_dict'].items() if k.startswith(prefix)})
    torch.save({'hparams': hp, 'state_dict': sd}, output_path)


@click.command('export-inference')
@click.option('--latent-ckpt-path', type=click.Path(exists=True, dir_okay=False), default='latent.ckpt', help='path to the latent checkpoint')
@click.option('--denoiser-ckpt-path', type=click.Path(exists=True, dir_okay=False), default='denoiser.ckpt', help='path to the denoiser checkpoint')
@click.option('--style-ckpt-path', type=click.Path(exists=True, dir_okay=False), default='style.ckpt', help='path to the style model checkpoint')
@click.option('--output-path', type=click.Path(exists=False, dir_okay=False), default='inference.pt', help='artifact output path')
def export_inference(latent_ckpt_path: str, denoiser_ckpt_path: str, style_ckpt_path: str, output_path: str):
    """export an inference model artifact from training checkpoints"""
    save_inference(latent_ckpt_path, denoiser_ckpt_path, style_ckpt_path, output_path)


@click.command('fit-denoiser')
@click.option('-c', '--config', type=click.Path(exists=True, dir_okay=False), required=True, help='config file')
@click.option('--ckpt-path', type=click.Path(exists=True, dir_okay=False), help='checkpoint from which to resume training')
@click.option('--synthetic', is_flag=True, help='train on synthetic latents instead of the cached dataset')
@click.option('--max-steps', type=int, default=None, help='stop after this many optimizer steps')
@click.option('--out', type=click.Path(dir_okay=False), default='denoiser.ckpt', help='checkpoint to write')
def fit_denoiser(config: str, ckpt_path: str | None, synthetic: bool, max_steps: int | None, out: str):
    """begin a training run for the diffusion model."""
    cfg = yaml.safe_load(open(config))
    for w in check_config(cfg):
        print('warning:', w, file=sys.stderr)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise click.ClickException('fit-denoiser needs a CUDA device: the B200 path has no CPU fallback')
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    seed = cfg.get('seed_everything', True)
    torch.manual_seed((0 if seed is True else int(seed)) + rank)
    tr = build_trainer(cfg)
    first_epoch, best = 0, float('inf')
    if ckpt_path:
        ck = load_checkpoint(ckpt_path, tr)
        # resume AFTER the stored epoch (its window order, seeded by the epoch number, was already trained on) and keep the
        # best validation loss, so a worse first validation does not overwrite the kept checkpoint
        first_epoch = int(ck.get('epoch', -1)) + 1
        if ck.get('best_model_score') is not None:
            best = float(ck['best_model_score'])
    else:  # every rank starts from rank 0's initialisation
        tr.diffusion_ema.module.load_state_dict(tr.diffusion.state_dict())
    tr = tr.cuda()
    if world > 1 and not ckpt_path:
        import torch.distributed as dist
        for p in list(tr.diffusion.parameters()) + list(tr.diffusion_ema.module.parameters()):
            dist.broadcast(p.data, 0)
    d = cfg['data']
    tcfg = cfg.get('trainer') or {}
    log_every = int(tcfg.get('log_every_n_steps', 50))
    max_epochs = int(tcfg.get('max_epochs', -1))
    if synthetic:
        def epochs():
            yield first_epoch, synthetic_batches(d['batch_size'], d['seq_len'], seed=rank + 7919 * first_epoch)
        val_sets = None
    else:
        train_sets, val_sets = split_mapsets(Path(d.get('data_path', './data')), '*.latent.npz',
                                             d.get('max_val_count', 512), d.get('max_val_frac', .3))
        def epochs():
            e = first_epoch
            while max_epochs < 0 or e < max_epochs:
                yield e, DeviceFeeder(LatentWindows(train_sets, d['seq_len'], d.get('shuffle_buffer_size', 1),
                                                 d.get('max_per_map', -1), seed=e), d['batch_size'], rank, world,
                                   device=torch.device('cuda', local))
                e += 1
    t0 = time.time()
    epoch = first_epoch
    for epoch, it in epochs():
        for batch in it:
            batch = tuple(t.cuda(non_blocking=True) for t in batch)
            loss, log = tr.training_step(batch, world_size=world)
            if rank == 0 and tr.global_step % log_every == 0:
                print(f'step {tr.global_step} lr {tr.current_lr():.2e} ' +
                      ' '.join(f'train/{k} {float(v):.4f}' for k, v in log.items()) + f' [{time.time() - t0:.0f}s]', flush=True)
            if max_steps is not None and tr.global_step >= max_steps:
                break
        if val_sets:
            # validation maps are sharded over the ranks (map i on rank i mod world) and the sums all-reduced: no rank waits
            # in the next epoch's gradient all-reduce while another validates, and every rank's generator advances alike
            acc = torch.zeros(2, dtype=torch.float64, device='cuda')
            for i, vb in enumerate(LatentWindows(val_sets, None)):
                if i % world == rank:
                    acc[0] += float(tr.validation_step(tuple(t[None].cuda() for t in vb))['val/loss'])
                    acc[1] += 1
            if world > 1:
                import torch.distributed as dist
                dist.all_reduce(acc)
            vl = float(acc[0]) / max(1.0, float(acc[1]))
            if rank == 0:
                print(f'epoch {epoch} val/loss {vl:.4f}', flush=True)
                if vl < best:  # ModelCheckpoint(monitor=val/loss, mode=min, save_top_k=1), model.yml:15-21
                    save_checkpoint(out, tr, epoch, vl)
            best = min(best, vl)
        if max_steps is not None and tr.global_step >= max_steps:
            break
    if rank == 0 and (not val_sets or not os.path.exists(out)):
        save_checkpoint(out, tr, epoch, best if best < float('inf') else None)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def build_style_trainer(cfg: dict):
    from .style_trainer import StyleTrainer
    clip = (cfg.get('trainer') or {}).get('gradient_clip_val', 0.0) or 0.0
    m = dict(cfg['model'])
    m.setdefault('schedule_args', {})  # the reference's style model.yml leaves it to LRScheduleArgs' defaults (constant LR)
    return StyleTrainer(**m, gradient_clip_val=float(clip))


@click.command('fit-style')
@click.option('-c', '--config', type=click.Path(exists=True, dir_okay=False), required=True, help='config file')
@click.option('--ckpt-path', type=click.Path(exists=True, dir_okay=False), help='checkpoint from which to resume training')
@click.option('--synthetic', is_flag=True, help='train on synthetic style codes instead of the cached dataset')
@click.option('--max-steps', type=int, default=None, help='stop after this many optimizer steps')
@click.option('--out', type=click.Path(dir_okay=False), default='style.ckpt', help='checkpoint to write')
def fit_style(config: str, ckpt_path: str | None, synthetic: bool, max_steps: int | None, out: str):
    """begin a training run for the style model (reference: scripts/fit_style.py:17-31, models/style/model.yml)."""
    cfg = yaml.safe_load(open(config))
    for w in check_config(cfg):
        print('warning:', w, file=sys.stderr)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise click.ClickException('fit-style needs a CUDA device: the B200 path has no CPU fallback')
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    seed = cfg.get('seed_everything', True)
    torch.manual_seed((0 if seed is True else int(seed)) + rank)
    tr = build_style_trainer(cfg)
    first_epoch, best = 0, float('inf')
    if ckpt_path:
        ck = load_checkpoint(ckpt_path, tr)
        first_epoch = int(ck.get('epoch', -1)) + 1
        if ck.get('best_model_score') is not None:
            best = float(ck['best_model_score'])
    else:
        tr.style_ema.module.load_state_dict(tr.style.state_dict())
    tr = tr.cuda()
    if world > 1 and not ckpt_path:
        import torch.distributed as dist
        for t in list(tr.style.state_dict().values()) + list(tr.style_ema.module.state_dict().values()):
            dist.broadcast(t.data, 0)
    d = cfg['data']
    tcfg = cfg.get('trainer') or {}
    log_every = int(tcfg.get('log_every_n_steps', 50))
    max_epochs = int(tcfg.get('max_epochs', -1))
    if synthetic:
        def epochs():
            yield first_epoch, synthetic_batches(d['batch_size'], d['seq_len'], seed=rank + 7919 * first_epoch)
        val_sets = None
    else:
        train_sets, val_sets = split_mapsets(Path(d.get('data_path', './data')), '*.latent.npz',
                                             d.get('max_val_count', 512), d.get('max_val_frac', .3))
        def epochs():
            e = first_epoch
            while max_epochs < 0 or e < max_epochs:
                yield e, DeviceFeeder(LatentWindows(train_sets, d['seq_len'], d.get('shuffle_buffer_size', 1),
                                                    d.get('max_per_map', -1), seed=e), d['batch_size'], rank, world,
                                      device=torch.device('cuda', local))
                e += 1
    t0 = time.time()
    epoch = first_epoch
    for epoch, it in epochs():
        for batch in it:
            batch = tuple(t.cuda(non_blocking=True) for t in batch)
            loss, log = tr.training_step(batch, world_size=world)
            if rank == 0 and tr.global_step % log_every == 0:
                print(f'step {tr.global_step} lr {tr.current_lr():.2e} ' +
                      ' '.join(f'train/{k} {float(v):.4f}' for k, v in log.items()) + f' [{time.time() - t0:.0f}s]', flush=True)
            if max_steps is not None and tr.global_step >= max_steps:
                break
        if val_sets and rank == 0:  # the style validation compares the whole validation set at once (train.py:120-150)
            tr.on_validation_epoch_start()
            for vb in LatentWindows(val_sets, None):  # full maps, one per item (data/modules/latent.py:52, seq_len=None)
                tr.validation_step(tuple(t[None].cuda() for t in vb))
            vals = tr.on_validation_epoch_end()
            print(f'epoch {epoch} ' + ' '.join(f'{k} {float(v):.4f}' for k, v in vals.items()), flush=True)
            ed = float(vals.get('val/energy_dist', float('inf')))
            if ed < best:  # ModelCheckpoint(monitor=val/energy_dist, mode=min, save_top_k=1), models/style/model.yml:13-18
                best = ed
                save_checkpoint(out, tr, epoch, best)
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        if max_steps is not None and tr.global_step >= max_steps:
            break
    if rank == 0 and (not val_sets or not os.path.exists(out)):
        save_checkpoint(out, tr, epoch, best if best < float('inf') else None)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


@click.command('predict', context_settings=dict(ignore_unknown_options=True, allow_extra_args=True))
@click.pass_context
def predict(ctx: click.Context):
    """generate osu!std maps from raw audio: the reference's own `predict` command (scripts/predict.py:21-100, same
    options) with its `diffusion.sample` call (models/inference/model.py:50) running on the B200 path.

    The audio front end, the latent / style models and the `.osu` codec are the reference's (outside this package), so
    an importable reference checkout (`osu_dreamer` on sys.path, e.g. PYTHONPATH=/path/to/osu-dreamer) is required."""
    if not torch.cuda.is_available():
        raise click.ClickException('predict needs a CUDA device: the B200 path has no CPU fallback')
    try:
        install()
        from osu_dreamer.scripts.predict import predict as ref_predict
    except ImportError as e:
        raise click.ClickException(f'predict needs an importable reference checkout (osu_dreamer on sys.path): {e}')
    ref_predict.main(args=list(ctx.args), standalone_mode=False)


@click.group()
def main():
    pass


main.add_command(fit_denoiser)
main.add_command(fit_style)
main.add_command(predict)
main.add_command(export_inference)

if __name__ == '__main__':
    main()
