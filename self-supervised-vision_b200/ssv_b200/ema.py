"""Multi-tensor EMA of the momentum (key / target / teacher) network — the per-step update of the reference's
models/moco.py:108-111, byol.py:120-123, relic.py:119-122 and dino.py:129-134:

    for t_param, s_param in zip(target.parameters(), source.parameters()):
        t_param.data = m * t_param.data + (1.0 - m) * s_param.data

One kernel launch for the whole network (the reference launches three elementwise kernels per parameter tensor),
bit-exact with the eager expression (products and sum rounded separately, scalars rounded like torch's).
"""
from __future__ import annotations

import weakref

import numpy as np
import torch

from . import _cabi as C


class EmaUpdater:
    """`step(m)` runs target = m * target + (1 - m) * source over a parameter list, in place, in ONE launch.

    The updater keeps the Parameter / tensor OBJECTS and re-derives the device chunk table from their *live*
    `data_ptr()`s on every step (a few microseconds of host work), so parameters that were re-pointed after
    construction (`module.to()`, `p.data = ...` as in the reference's own loop, FSDP-style re-allocation) are followed
    instead of silently updating dead storage.  Pairs the kernel cannot take (non-fp32, non-contiguous such as
    channels_last, CPU) fall back to the reference's three eager ops for that pair only; they do not raise."""

    def __init__(self, target_params, source_params):
        self.targets = list(target_params)
        self.sources = list(source_params)
        if len(self.targets) != len(self.sources):
            raise ValueError("target and source parameter lists differ in length")
        for t, s in zip(self.targets, self.sources):
            if t.shape != s.shape:
                raise ValueError("EMA needs parameter pairs of equal shape")
        self._key = None
        self._table = None
        self._n = 0
        self._slow = []
        self.device = None

    @staticmethod
    def _live(p):
        return p.data if isinstance(p, torch.nn.Parameter) else p

    @staticmethod
    def _fast_ok(t, s):
        return (t.is_cuda and s.is_cuda and t.device == s.device and t.dtype == torch.float32 and s.dtype == torch.float32
                and t.is_contiguous() and s.is_contiguous() and t.data_ptr() % 4 == 0 and s.data_ptr() % 4 == 0)

    def _build(self):
        live = [(self._live(t), self._live(s)) for t, s in zip(self.targets, self.sources)]
        key = tuple((t.data_ptr(), s.data_ptr(), t.numel(), t.dtype, t.device, t.is_contiguous() and s.is_contiguous())
                    for t, s in live)
        if key == self._key:
            return live
        fast = [(t, s) for t, s in live if t.numel() and self._fast_ok(t, s)]
        devs = {t.device for t, _ in fast}
        if len(devs) > 1:   # one launch serves one device: keep the majority device on the kernel, the rest eager
            dev0 = max(devs, key=lambda d: sum(t.numel() for t, _ in fast if t.device == d))
            fast = [(t, s) for t, s in fast if t.device == dev0]
        fast_ids = {id(t) for t, _ in fast}
        self._slow = [i for i, (t, _) in enumerate(live) if t.numel() and id(t) not in fast_ids]
        self.device = fast[0][0].device if fast else None
        chunk = int(C.lib().ssvb_ema_chunk_elems()) if fast else 1
        rows = []
        for t, s in fast:
            tp, sp, n = t.data_ptr(), s.data_ptr(), t.numel()
            for off in range(0, n, chunk):
                rows.append((tp + 4 * off, sp + 4 * off, min(chunk, n - off)))
        tab = np.asarray(rows, dtype=np.uint64).reshape(-1, 3)
        self._table = torch.from_numpy(tab.view(np.int64)).to(self.device) if fast else None
        self._n = tab.shape[0]
        self._key = key
        return live

    @torch.no_grad()
    def step(self, m):
        live = self._build()
        if self._n:
            with C.on_device(self.device):
                C.check(C.lib().ssvb_ema_update(C.ptr(self._table), self._n, float(m), float(1.0 - m),
                                                C.stream_ptr(self.device)), "ssvb_ema_update")
        for i in self._slow:     # the reference's own expression (models/moco.py:108-111) for pairs the kernel cannot take
            t, s = live[i]
            t.copy_(m * t + (1.0 - m) * s.to(t.dtype))


_UPDATERS = weakref.WeakKeyDictionary()   # target module -> {id(source module): (weakref(source), EmaUpdater)}


@torch.no_grad()
def momentum_update(target_module, source_module, m):
    """Drop-in body for the reference's `momentum_update()` / `update_teacher_model()`:
    target = m * target + (1 - m) * source over all parameters, in place, one launch.
    The updater cache is keyed weakly on the target module (no id() reuse after garbage collection, nothing pinned)."""
    per_target = _UPDATERS.setdefault(target_module, {})
    entry = per_target.get(id(source_module))
    if entry is None or entry[0]() is not source_module:
        up = EmaUpdater(list(target_module.parameters()), list(source_module.parameters()))
        per_target[id(source_module)] = (weakref.ref(source_module), up)
    else:
        up = entry[1]
    up.step(m)
