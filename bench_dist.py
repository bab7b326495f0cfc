#!/usr/bin/env python
"""Multi-GPU companion of bench_losses.py: the partitioned losses of SURVEY.md §8(e) at BASELINE.json's configs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench_dist.py [--reps 20]          (N = 1 runs without torchrun as well)

STRONG scaling on the fixed global problem (the BASELINE config), one process per GPU over NCCL:
  * Barlow Twins cfg3: global batch 2048 x 8192, rows sharded, cross-correlation all-reduced
    (reduce-scatter fp32 -> fused loss/dC slab epilogue -> all-gather bf16);
  * Sinkhorn cfg4: 4096 x 3000 scores, rows sharded, K-vector marginals exchanged per pass;
  * SwAV loss: 4096 global rows x 3000 prototypes x 128;
  * MoCo cfg2: 256 global queries x 65536-entry queue sharded over the ranks, + sharded enqueue.
Every number: forward + backward through the public Distributed* API, inputs resident, CUDA events bracketed by a
barrier, L2 flushed between repetitions, median of `reps`, MAX over ranks.  Rank 0 prints one JSON object per line."""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "self-supervised-vision_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def randn(seed, *shape):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float32)


def unit(x):
    return torch.nn.functional.normalize(x, dim=-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from ssv_b200.dist import (DistributedBarlowLoss, DistributedMocoLoss, DistributedSwavLoss, ShardedMemoryBank)

    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def timed(fn, reps):
        for _ in range(4):
            fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(t.item())
        return statistics.median(ts)

    def emit(name, cfg, ms, samples, flops=None, bytes_=None, note=""):
        if rank != 0:
            return
        r = {"loss": name, "config": cfg, "n_gpus": world, "ms": ms, "samples_per_s": samples / (ms * 1e-3),
             "scaling": "strong", "note": note}
        if flops:
            r.update(bound="tensor", achieved_tflops=flops / (ms * 1e-3) / 1e12,
                     frac_of_aggregate_peak=flops / (ms * 1e-3) / 1e12 / (peaks["bf16_tflops"] * world))
        if bytes_:
            r.update(bound="hbm", achieved_gbs=bytes_ / (ms * 1e-3) / 1e9,
                     frac_of_aggregate_peak=bytes_ / (ms * 1e-3) / 1e9 / (peaks["hbm_gbs"] * world))
        print(json.dumps(r), flush=True)

    def fwd_bwd(loss_fn, *tensors):
        def run():
            for t in tensors:
                t.grad = None
            loss_fn().backward()
        return run

    # ---- Barlow cfg3: global 2048 x 8192
    n, d = 2048, 8192
    nl = n // world
    g = torch.Generator().manual_seed(7)
    sig, mu = torch.rand(d, generator=g) * 1.5 + 0.5, torch.randn(d, generator=g)
    zi = randn(10 + rank, nl, d) * sig + mu
    zj = 0.7 * zi + 0.3 * (randn(50 + rank, nl, d) * sig + mu)
    a, b = zi.to(dev).requires_grad_(True), zj.to(dev).requires_grad_(True)
    fn_b = DistributedBarlowLoss(False, 0.005)
    ms = timed(fwd_bwd(lambda: fn_b(a, b), a, b), max(5, args.reps // 2))
    emit("DistributedBarlowLoss[allreduce]", f"cfg3 global 2048x8192 ({nl} rows/rank), C all-reduced", ms, n,
         flops=6 * n * d * d, note="reduce-scatter fp32 C (256 MiB) + all-gather bf16 dC (128 MiB) per step")
    del fn_b
    torch.cuda.empty_cache()
    if d % (8 * world) == 0:
        fn_c = DistributedBarlowLoss(False, 0.005, mode="colshard")
        ms = timed(fwd_bwd(lambda: fn_c(a, b), a, b), max(5, args.reps // 2))
        emit("DistributedBarlowLoss[colshard]", f"cfg3 global 2048x8192 ({nl} rows/rank), column slab of C and C^T per rank",
             ms, n, flops=6 * n * d * d,
             note="all-gather bf16 standardised rows (64 MiB), all-to-all of the gradient slabs; C never crosses NVLink")
        del fn_c
    del a, b
    torch.cuda.empty_cache()

    # ---- Sinkhorn cfg4: global 4096 x 3000
    bsz, kp = 4096, 3000
    bl = bsz // world
    scores = (unit(randn(100 + rank, bl, 128)) @ unit(randn(1, kp, 128)).t()).contiguous().to(dev)
    fn_s = DistributedSwavLoss(0.1, 0.05, 3)
    ms = timed(lambda: fn_s.compute_codes_sinkhorn(scores), args.reps)
    emit("DistributedSwavLoss.compute_codes_sinkhorn", f"cfg4 global 4096x3000 ({bl} rows/rank) 3 iters", ms, bsz,
         bytes_=2 * bsz * kp * 4, note="one all-gather of (K+1) floats per rank per pass")

    # ---- SwAV loss: global 4096 rows x 3000 prototypes x 128
    z1, z2 = unit(randn(200 + rank, bl, 128)), unit(randn(300 + rank, bl, 128))
    c = unit(randn(2, kp, 128))
    a, b, pc = z1.to(dev).requires_grad_(True), z2.to(dev).requires_grad_(True), c.to(dev).requires_grad_(True)
    ms = timed(fwd_bwd(lambda: fn_s(a, b, pc), a, b, pc), args.reps)
    emit("DistributedSwavLoss", f"global 4096 rows ({bl}/rank) x 3000 prototypes x 128", ms, bsz,
         bytes_=6 * bsz * kp * 4 * 2, note="marginals all-gather x3, loss + dC all-reduce")

    # ---- MoCo cfg2: 256 global queries x 65536 queue (sharded) + enqueue
    n, k, d = 256, 65536, 128
    nl = n // world
    bank = ShardedMemoryBank(k, d)
    for i in range(4):
        bank.add_batch(randn(400 + 10 * i + rank, 16384 // world, d).to(dev))
    q, kk = randn(500 + rank, nl, d), randn(600 + rank, nl, d)
    a, b = q.to(dev).requires_grad_(True), kk.to(dev).requires_grad_(True)
    fn_m = DistributedMocoLoss(True, 0.07)
    mem = bank.get_vectors()
    ms = timed(fwd_bwd(lambda: fn_m(a, b, mem), a, b), args.reps)
    emit("DistributedMocoLoss", f"cfg2 256 global queries x 65536 queue ({k // world} rows/rank) tau=0.07", ms, n,
         bytes_=2 * k * d * 4, note="all-gather q (bf16), all-gather shard partials, reduce-scatter dq")
    kd = b.detach()
    ms = timed(lambda: bank.add_batch(kd), args.reps)
    emit("ShardedMemoryBank.add_batch", "cfg2 enqueue 256 global keys into the sharded 65536 ring", ms, n,
         bytes_=2 * n * d * 4, note="all-gather keys + one enqueue kernel per rank")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
