"""B200-native osu-dreamer denoiser hot path (DiffusionModel forward/backward + sampler).

Host side: a Python mirror of the reference's `osu_dreamer/models/diffusion` interface.
Device side: hand-written sm_100a CUDA behind the C ABI declared in include/osd_b200.h
(libosd_b200.so, built in-tree by `__graft_entry__.build()` / `make -C osu-dreamer_b200/csrc`).
"""
from . import lib  # noqa: F401

__all__ = ['lib']
