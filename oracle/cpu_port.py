"""CPU baseline port of the NT-Xent hot path (TEST / BENCH INFRASTRUCTURE ONLY — never on the product path).

Used by bench.py's `cpu_baseline` leg and by `bench.py --impl reference`.  The reference itself is Python
and cannot travel to the GPU box (/root/reference does not exist there), so the CPU arm times this port:
torch CPU, fp32 (the reference's dtype), all host threads, forward + backward of SimclrLoss
(reference utils/losses.py:15-46) in closed form.

Bounded sample: the reference materialises > 150 GB of N x N intermediates at N = 32768 and
cannot run BASELINE's headline config at all; this port instead evaluates a ROW SLAB of the same problem —
`rows` of the M = 2N similarity rows against all M columns, loss terms plus both gradient contributions of
those rows (dZ_R += G Z and dZ += G^T Z_R) — which is exactly rows/M of the full job's 6*M^2*d FLOPs and of its
exp / mask / reduction work.  samples/s = (N * rows / M) / seconds.
"""
from __future__ import annotations

import time

import torch


def ntxent_row_slab(zhat: torch.Tensor, n: int, r0: int, rows: int, temperature: float):
    """Loss sum and gradient contributions (w.r.t. the normalised rows) of similarity rows [r0, r0+rows)."""
    m = zhat.shape[0]
    zr = zhat[r0:r0 + rows]
    s = (zr @ zhat.t()) / temperature                      # rows x M logits   (utils/losses.py:27-30)
    idx = torch.arange(r0, r0 + rows)
    ar = torch.arange(rows)
    partner = (idx + n) % m
    pos = s[ar, partner].clone()                           # positives          (:32-33)
    s[ar, idx] = float("-inf")                             # drop self-similarity (:34-37 mask)
    lse = torch.logsumexp(s, dim=1)                        # cross entropy, label 0 (:45)
    loss_sum = (lse - pos).sum()
    g = torch.exp(s - lse[:, None])                        # softmax rows
    g[ar, partner] -= 1.0
    g /= (m * temperature)
    d_rows = g @ zhat                                      # dZ_R += G Z
    d_cols = g.t() @ zr                                    # dZ   += G^T Z_R
    return loss_sum, d_rows, d_cols


def time_ntxent_sample(n: int, d: int, temperature: float, rows: int, reps: int = 1, seed: int = 0, threads=None):
    """Returns dict(seconds per slab, samples/s equivalent, threads used)."""
    if threads:
        torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(seed)
    zi = torch.randn(n, d, generator=g)
    zj = torch.randn(n, d, generator=g)
    m = 2 * n
    rows = min(rows, m)
    times = []
    for it in range(reps + 1):  # first pass is the warm-up
        t0 = time.perf_counter()
        zhat = torch.nn.functional.normalize(torch.cat([zi, zj]), dim=-1)   # (:20-25)
        r0 = ((max(it - 1, 0)) * rows) % (m - rows + 1) if m > rows else 0   # a different slab of rows every repetition
        loss_sum, d_rows, d_cols = ntxent_row_slab(zhat, n, r0, rows, temperature)
        d_cols[r0:r0 + rows] += d_rows
        float(loss_sum)
        if it > 0 or reps == 0:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"seconds": sec, "samples_per_s": (n * rows / m) / sec, "threads": torch.get_num_threads(),
            "rows": rows, "m": m, "reps": len(times), "total_seconds": sum(times)}
