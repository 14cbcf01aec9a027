"""`LDM` -- the reference's inference pipeline (osu_dreamer/models/inference/model.py:17-52) assembled from the three B200
mirrors: `latent` (audio_encoder / decode), `style` (sphere-tracing sampler) and `diffusion` (the denoiser hot path).
Same constructor argument, attribute names and state-dict keys (`latent.*`, `style.*`, `diffusion.*`) as the reference,
so `load_artifact` reads the files written by `export-inference` (models/inference/artifact.py:9-49) and `sample(audio,
labels, num_steps)` is the call `predict` makes (scripts/predict.py:71-75).  Inference only, CUDA only.
"""
from __future__ import annotations

from dataclasses import dataclass, fields, is_dataclass

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from . import lib
from .denoiser import BackboneArgs, DiffusionModel, DiffusionModelArgs
from .latent import A_DIM, LatentModel, LatentModelArgs, LayerArgs
from .style import StyleModel, StyleModelArgs


@dataclass
class LDMArgs:  # models/inference/model.py:17-25
    emb_dim: int
    style_dim: int
    n_downs: int
    stride: int
    latent_args: LatentModelArgs
    style_args: StyleModelArgs
    diffusion_args: DiffusionModelArgs


_NESTED = {'latent_args': LatentModelArgs, 'ae_args': LayerArgs, 'style_args': StyleModelArgs,
           'diffusion_args': DiffusionModelArgs, 'backbone_args': BackboneArgs}


def _from_dict(cls, data):
    """dicts -> the args dataclasses, recursively (the job of artifact.py:52-71 `dataclass_from_dict`)"""
    if not isinstance(data, dict):
        return data
    names = {f.name for f in fields(cls)}
    return cls(**{k: (_from_dict(_NESTED[k], v) if k in _NESTED else v) for k, v in data.items() if k in names})


def pad_to_multiple(x: Tensor, chunk_size: int) -> Tensor:
    """right-pad the time axis by replication to a multiple of chunk_size (data/modules/beatmap.py:26-30)"""
    pad = (chunk_size - x.size(-1) % chunk_size) % chunk_size
    return F.pad(x, (0, pad), mode='replicate') if pad > 0 else x


class LDM(nn.Module):
    def __init__(self, args: LDMArgs):
        super().__init__()
        if isinstance(args, dict):
            args = _from_dict(LDMArgs, args)
        self.latent = LatentModel(args.emb_dim, args.style_dim, args.n_downs, args.stride, args.latent_args)
        self.style = StyleModel(args.style_dim, args.style_args)
        self.diffusion = DiffusionModel(args.emb_dim, args.latent_args.h_dim, args.style_dim, args.diffusion_args)

    @torch.no_grad()
    def sample(self, audio: Tensor, labels: Tensor, num_steps: int, show_progress: bool = False):
        """models/inference/model.py:34-52: audio [72, L], labels [B, 5] -> (chart [B, 9, L], labels [B, 5])"""
        if audio.dim() != 2 or audio.size(0) != A_DIM:
            raise lib.OsdError(f'audio must be [{A_DIM}, L]')
        L = audio.size(-1)
        audio = pad_to_multiple(audio, self.latent.chunk_size)
        skips, h = self.latent.audio_encoder(audio[None])
        s = self.style.sample(labels)
        z = self.diffusion.sample(h, s, num_steps, show_progress=show_progress)
        chart, out_labels = self.latent.decode(z, s, skips=skips)
        return chart[..., :L], out_labels


def load_artifact(path, device='cuda') -> LDM:
    """models/inference/artifact.py:44-49 `load_inference` for the B200 pipeline"""
    art = torch.load(path, map_location='cpu', weights_only=False)
    model = LDM(_from_dict(LDMArgs, art['hparams']))
    model.load_state_dict(art['state_dict'])
    return model.eval().to(device)
