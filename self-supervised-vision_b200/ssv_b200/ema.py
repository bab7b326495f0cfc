"""Multi-tensor EMA of the momentum (key / target / teacher) network — the per-step update of the reference's
models/moco.py:108-111, byol.py:120-123, relic.py:119-122 and dino.py:129-134:

    for t_param, s_param in zip(target.parameters(), source.parameters()):
        t_param.data = m * t_param.data + (1.0 - m) * s_param.data

One kernel launch for the whole network (the reference launches three elementwise kernels per parameter tensor),
bit-exact with the eager expression (products and sum rounded separately, scalars rounded like torch's).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _cabi as C


class EmaUpdater:
    """Caches the device chunk table for a fixed (target, source) parameter list; `step(m)` runs the update in place."""

    def __init__(self, target_params, source_params):
        self.targets = [p.data if isinstance(p, torch.nn.Parameter) else p for p in target_params]
        self.sources = [p.data if isinstance(p, torch.nn.Parameter) else p for p in source_params]
        if len(self.targets) != len(self.sources):
            raise ValueError("target and source parameter lists differ in length")
        C.require_cuda(*self.targets, *self.sources)
        for t, s in zip(self.targets, self.sources):
            if t.shape != s.shape or t.dtype != torch.float32 or s.dtype != torch.float32 \
                    or not t.is_contiguous() or not s.is_contiguous():
                raise ValueError("EMA needs contiguous fp32 parameter pairs of equal shape")
        self.device = self.targets[0].device if self.targets else torch.device("cuda")
        self._key = None
        self._table = None
        self._n = 0

    def _build(self):
        key = tuple((t.data_ptr(), s.data_ptr(), t.numel()) for t, s in zip(self.targets, self.sources))
        if key == self._key:
            return
        chunk = int(C.lib().ssvb_ema_chunk_elems())
        rows = []
        for tp, sp, n in key:
            for off in range(0, n, chunk):
                rows.append((tp + 4 * off, sp + 4 * off, min(chunk, n - off)))
        tab = np.asarray(rows, dtype=np.uint64).reshape(-1, 3)
        self._table = torch.from_numpy(tab.view(np.int64)).to(self.device)
        self._n = tab.shape[0]
        self._key = key

    @torch.no_grad()
    def step(self, m):
        self._build()
        if self._n == 0:
            return
        with C.on_device(self.device):
            C.check(C.lib().ssvb_ema_update(C.ptr(self._table), self._n, float(m), float(1.0 - m),
                                            C.stream_ptr(self.device)), "ssvb_ema_update")


_UPDATERS = {}


@torch.no_grad()
def momentum_update(target_module, source_module, m):
    """Drop-in body for the reference's `momentum_update()` / `update_teacher_model()`:
    target = m * target + (1 - m) * source over all parameters, in place, one launch."""
    key = (id(target_module), id(source_module))
    up = _UPDATERS.get(key)
    if up is None:
        up = EmaUpdater(list(target_module.parameters()), list(source_module.parameters()))
        _UPDATERS[key] = up
    up.step(m)
