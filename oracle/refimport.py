"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference from /root/reference.

Only usable where /root/reference exists (this container, never the GPU box).
`models/diffusion/model.py` imports as-is; `models/diffusion/train.py` needs
four third-party packages that are not installed here (pytorch_lightning,
rosu_pp_py, torchcodec, resonators -- SURVEY.md 8(c)); minimal stub modules are
inserted into sys.modules so the reference's own loss / optimizer / EMA code
runs verbatim.  No reference source is copied.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get('OSD_REFERENCE_ROOT', '/root/reference')


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'osu_dreamer'))


def _install_stubs():
    import torch.nn as nn

    if 'pytorch_lightning' not in sys.modules:
        pl = types.ModuleType('pytorch_lightning')

        class LightningModule(nn.Module):
            def save_hyperparameters(self, *a, **k):
                pass

            def log_dict(self, d=None, *a, **k):
                self.logged = dict(d) if d is not None else {}  # lets make_golden.py read what validation_step logs

            def log(self, *a, **k):
                pass

        class LightningDataModule:
            def __init__(self, *a, **k):
                pass

        pl.LightningModule = LightningModule
        pl.LightningDataModule = LightningDataModule
        sys.modules['pytorch_lightning'] = pl
    if 'rosu_pp_py' not in sys.modules:
        sys.modules['rosu_pp_py'] = types.ModuleType('rosu_pp_py')
    if 'torchcodec' not in sys.modules:
        tc = types.ModuleType('torchcodec')
        dec = types.ModuleType('torchcodec.decoders')
        ad = types.ModuleType('torchcodec.decoders._audio_decoder')
        ad.AudioDecoder = type('AudioDecoder', (), {})
        tc.decoders = dec
        dec._audio_decoder = ad
        sys.modules['torchcodec'] = tc
        sys.modules['torchcodec.decoders'] = dec
        sys.modules['torchcodec.decoders._audio_decoder'] = ad
    if 'resonators' not in sys.modules:
        rs = types.ModuleType('resonators')
        rs.ResonatorBank = type('ResonatorBank', (), {})
        sys.modules['resonators'] = rs


def import_reference(with_trainer: bool = False):
    """Returns a namespace with DiffusionModel, DiffusionModelArgs, BackboneArgs
    (+ DiffusionTrainer, frame_dist_sq, LRScheduleArgs when with_trainer)."""
    if not available():
        raise RuntimeError(f'reference not present at {REFERENCE_ROOT}')
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    from osu_dreamer.models.diffusion.model import DiffusionModel, DiffusionModelArgs
    from osu_dreamer.models.diffusion.backbone import BackboneArgs
    ns.DiffusionModel, ns.DiffusionModelArgs, ns.BackboneArgs = DiffusionModel, DiffusionModelArgs, BackboneArgs
    if with_trainer:
        _install_stubs()
        from osu_dreamer.models.diffusion.train import DiffusionTrainer, frame_dist_sq
        from osu_dreamer.common.lr_schedule import LRScheduleArgs, make_lr_schedule
        ns.DiffusionTrainer, ns.frame_dist_sq = DiffusionTrainer, frame_dist_sq
        ns.LRScheduleArgs, ns.make_lr_schedule = LRScheduleArgs, make_lr_schedule
    return ns


def default_args(ns):
    """model.yml:77-90."""
    return ns.DiffusionModelArgs(
        global_cond_dim=512, backbone_dim=512, u_head_dim=64,
        backbone_args=ns.BackboneArgs(depth=8, expand=4, head_dim=64, n_heads=16, radius=2))


def build_reference_model(ns, state_dict=None, dtype=None):
    m = ns.DiffusionModel(6, 128, 32, default_args(ns))
    if state_dict is not None:
        m.load_state_dict(state_dict, strict=True)
    if dtype is not None:
        m = m.to(dtype)
    return m.eval()
