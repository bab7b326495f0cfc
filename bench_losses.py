#!/usr/bin/env python
"""Per-loss throughput + roofline fractions for every row of SURVEY.md §8(a) at BASELINE.json's configs
(the headline NT-Xent metric lives in bench.py; this is the companion table).

    python bench_losses.py [--reps 30] [--cpu]       -> one JSON object per line + a markdown table on stderr

Each measurement: forward + backward through the public drop-in API, inputs resident in HBM, CUDA events on the
current stream, L2 flushed (256 MiB write) between repetitions, median of `reps`; measured twice: eager (includes the
Python / launch cost per call, which dominates the small losses) and as a CUDA-graph replay of the same step (device
time; the roofline fraction is taken from it).  `--cpu` also times the oracle
port (numpy fp64 closed form, oracle/ssl_oracle.py) once per config on the host cores as the reported baseline.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "self-supervised-vision_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        pk = json.load(open(path))
        return pk["bf16_tflops"], pk["hbm_gbs"], "measured"
    return 1590.0, 6650.0, "fallback"


def randn(seed, *shape):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float32)


def unit(x):
    return torch.nn.functional.normalize(x, dim=-1)


def time_gpu(fn, reps, flush):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def time_graph(fn, reps, flush):
    """Same step captured once in a CUDA graph and replayed: device time without the per-launch Python / driver cost
    (what a graph-captured training step pays).  Returns None if capture is not possible."""
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)
    except Exception as e:  # noqa: BLE001
        print(f"[graph capture failed: {type(e).__name__}: {str(e)[:120]}]", file=sys.stderr)
        torch.cuda.synchronize()
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--cpu", action="store_true")
    args = ap.parse_args()
    import ssv_b200 as S
    O = None
    if args.cpu:  # the oracle is test / baseline infrastructure: only the --cpu leg may load it
        from oracle import ssl_oracle as O

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    tf_peak, hbm_peak, src = peaks()
    rows = []

    def record(name, cfg, ms, samples, flops=None, bytes_=None, cpu_s=None, note="", graph_ms=None):
        r = {"loss": name, "config": cfg, "ms": ms, "samples_per_s": samples / (ms * 1e-3), "graph_ms": graph_ms}
        t = graph_ms if graph_ms else ms   # roofline fraction from the device time (graph replay) when available
        if flops:
            r.update(bound="tensor", achieved_tflops=flops / (t * 1e-3) / 1e12, frac=flops / (t * 1e-3) / 1e12 / tf_peak)
        if bytes_:
            r.update(bound="hbm", achieved_gbs=bytes_ / (t * 1e-3) / 1e9, frac=bytes_ / (t * 1e-3) / 1e9 / hbm_peak)
        if cpu_s is not None:
            r.update(cpu_port_s=cpu_s, cpu_samples_per_s=samples / cpu_s, cpu_cores=os.cpu_count())
        r["peak_source"] = src
        r["note"] = note
        rows.append(r)
        print(json.dumps(r), flush=True)

    def cpu_time(fn):
        if not args.cpu:
            return None
        t0 = time.perf_counter()
        fn()
        return time.perf_counter() - t0

    def fwd_bwd(loss_fn, *tensors):
        def run():
            for t in tensors:
                if t.requires_grad:
                    t.grad = None
            loss_fn().backward()
        return run

    # ---- cfg1: SimCLR NT-Xent 2 x 256 x 128, tau 0.5 (the reference's own CPU-runnable case)
    zi, zj = randn(0, 256, 128), randn(1, 256, 128)
    a, b = zi.to(dev).requires_grad_(True), zj.to(dev).requires_grad_(True)
    fn = S.SimclrLoss(True, 0.5)
    _f = fwd_bwd(lambda: fn(a, b), a, b)
    ms = time_gpu(_f, args.reps, flush)
    gms = time_graph(_f, args.reps, flush)
    record("SimclrLoss", "cfg1 2x256x128 tau=0.5", ms, 256, flops=6 * 512 ** 2 * 128, graph_ms=gms,
           cpu_s=cpu_time(lambda: O.ntxent(zi.numpy(), zj.numpy(), True, 0.5)), note="latency-bound (4 CTAs)")

    # ---- NT-Xent mid sizes (scaling series)
    for n in (2048, 8192):
        zi, zj = randn(0, n, 128), randn(1, n, 128)
        a, b = zi.to(dev).requires_grad_(True), zj.to(dev).requires_grad_(True)
        _f = fwd_bwd(lambda: fn(a, b), a, b)
        ms = time_gpu(_f, args.reps, flush)
        gms = time_graph(_f, args.reps, flush)
        record("SimclrLoss", f"{n}x128 tau=0.5", ms, n, flops=6 * (2 * n) ** 2 * 128, graph_ms=gms)

    # ---- cfg2: MoCo 256 queries x 65536-entry queue x 128 + enqueue (device-resident bank, bf16 shadow)
    n, k, d = 256, 65536, 128
    bank = S.MemoryBank(k, d)
    fill = randn(5, k, d).to(dev)
    for i in range(4):
        bank.add_batch(fill[i * 16384:(i + 1) * 16384])
    q, kk = randn(0, n, d), randn(1, n, d)
    a, b = q.to(dev).requires_grad_(True), kk.to(dev).requires_grad_(True)
    fn_m = S.MocoLoss(True, 0.07)
    mem = bank.get_vectors()
    _f = fwd_bwd(lambda: fn_m(a, b, mem), a, b)
    ms = time_gpu(_f, args.reps, flush)
    gms = time_graph(_f, args.reps, flush)
    mem_np = mem.cpu().numpy()
    record("MocoLoss", "cfg2 256x65536x128 tau=0.07", ms, n, bytes_=2 * k * d * 4, graph_ms=gms,
           cpu_s=cpu_time(lambda: O.moco(q.numpy(), kk.numpy(), mem_np, True, 0.07)),
           note="algorithmic bytes = queue read twice as fp32 (67.1 MB); kernels read the bf16 shadow")
    kd = b.detach()
    _f = lambda: bank.add_batch(kd)
    ms = time_gpu(_f, args.reps, flush)
    gms = time_graph(_f, args.reps, flush)
    record("MemoryBank.add_batch", "cfg2 enqueue 256x128 into 65536", ms, n, bytes_=2 * n * d * 4, graph_ms=gms,
           cpu_s=cpu_time(lambda: O.ring_enqueue(mem_np, 0, kk.numpy(), True)), note="launch-latency bound")

    # ---- cfg3: Barlow Twins 2048 x 8192
    n, d = 2048, 8192
    g = torch.Generator().manual_seed(7)
    sig, mu = torch.rand(d, generator=g) * 1.5 + 0.5, torch.randn(d, generator=g)
    zi = randn(0, n, d) * sig + mu
    zj = 0.7 * zi + 0.3 * (randn(1, n, d) * sig + mu)
    a, b = zi.to(dev).requires_grad_(True), zj.to(dev).requires_grad_(True)
    fn_b = S.BarlowLoss(False, 0.005)
    _f = fwd_bwd(lambda: fn_b(a, b), a, b)
    ms = time_gpu(_f, max(5, args.reps // 3), flush)
    gms = time_graph(_f, max(5, args.reps // 3), flush)
    record("BarlowLoss", "cfg3 2048x8192 lambda=0.005", ms, n, flops=6 * n * d * d, graph_ms=gms,
           cpu_s=cpu_time(lambda: O.barlow(zi.numpy(), zj.numpy(), False, 0.005)))
    del a, b

    # ---- cfg4: Sinkhorn-Knopp 4096 x 3000, 3 iters, eps 0.05
    bsz, kp = 4096, 3000
    scores = (unit(randn(0, bsz, 128)) @ unit(randn(1, kp, 128)).t()).contiguous()
    sd = scores.to(dev)
    fn_s = S.SwavLoss(0.1, 0.05, 3)
    _f = lambda: fn_s.compute_codes_sinkhorn(sd)
    ms = time_gpu(_f, args.reps, flush)
    gms = time_graph(_f, args.reps, flush)
    record("SwavLoss.compute_codes_sinkhorn", "cfg4 4096x3000 3 iters eps=0.05", ms, bsz, bytes_=2 * bsz * kp * 4,
           graph_ms=gms,
           cpu_s=cpu_time(lambda: O.sinkhorn(scores.numpy(), 0.05, 3)))

    # ---- SwAV full loss (reference shapes: 512 live + 3000 bank rows, 3000 prototypes, d = 128)
    z1, z2 = unit(randn(0, 512, 128)), unit(randn(1, 512, 128))
    c, bk = unit(randn(2, 3000, 128)), unit(randn(3, 3000, 128))
    a, b, pc = z1.to(dev).requires_grad_(True), z2.to(dev).requires_grad_(True), c.to(dev).requires_grad_(True)
    bkd = bk.to(dev)
    _f = fwd_bwd(lambda: fn_s(a, b, pc, bkd), a, b, pc)
    ms = time_gpu(_f, args.reps, flush)
    gms = time_graph(_f, args.reps, flush)
    bp = 3512
    record("SwavLoss", "512+3000 bank x 3000 prototypes x 128", ms, 512, bytes_=6 * bp * 3000 * 4 * 2, graph_ms=gms,
           cpu_s=cpu_time(lambda: O.swav(z1.numpy(), z2.numpy(), c.numpy(), bk.numpy())))

    # ---- BYOL MSE / SimSiam
    for (n, d) in ((32768, 128), (4096, 1024)):
        o, t = unit(randn(0, n, d)), unit(randn(1, n, d))
        a, b = o.to(dev).requires_grad_(True), t.to(dev)
        fn_e = S.MSELoss()
        _f = fwd_bwd(lambda: fn_e(a, b), a)
        ms = time_gpu(_f, args.reps, flush)
        gms = time_graph(_f, args.reps, flush)
        record("MSELoss (BYOL)", f"{n}x{d}", ms, n, bytes_=(2 + 3) * n * d * 4, graph_ms=gms,
               cpu_s=cpu_time(lambda: O.mse(o.numpy(), t.numpy())))
        b2 = t.to(dev).requires_grad_(True)
        fn_ss = S.SimSiamLoss()
        _f = fwd_bwd(lambda: fn_ss(a, b2), a, b2)
        ms = time_gpu(_f, args.reps, flush)
        gms = time_graph(_f, args.reps, flush)
        record("SimSiamLoss", f"{n}x{d}", ms, n, bytes_=(2 + 4) * n * d * 4, graph_ms=gms,
               cpu_s=cpu_time(lambda: O.simsiam(o.numpy(), t.numpy())))

    # ---- ReLIC
    for n in (512, 4096):
        zi, zj, zo = randn(0, n, 128), randn(1, n, 128), randn(2, n, 128)
        a, b, c3 = (x.to(dev).requires_grad_(True) for x in (zi, zj, zo))
        fn_r = S.RelicLoss(True, 1.0, 0.5)
        _f = fwd_bwd(lambda: fn_r(a, b, c3), a, b, c3)
        ms = time_gpu(_f, args.reps, flush)
        gms = time_graph(_f, args.reps, flush)
        record("RelicLoss", f"{n}x128 tau=1 alpha=0.5", ms, n, flops=6 * (2 * n) ** 2 * 128, graph_ms=gms,
               cpu_s=cpu_time(lambda: O.relic(zi.numpy(), zj.numpy(), zo.numpy(), True, 1.0, 0.5)),
               note="FLOPs of the contrastive part only; the KL term adds 9*N*d*4 bytes")

    # ---- DinoLoss (reference shape bs 64, 2+6 views, K 1024; and a large one) and the parameter EMA
    for (bs, nv, k) in ((64, 8, 1024), (1024, 8, 4096)):
        teacher, student, center = randn(0, bs, 2, k), randn(1, bs, nv, k), 0.1 * randn(2, k)
        t, st, c = teacher.to(dev), student.to(dev).requires_grad_(True), center.to(dev)
        fn_d = S.DinoLoss()
        _f = fwd_bwd(lambda: fn_d(t, st, 0.1, 0.04, c), st)
        ms = time_gpu(_f, args.reps, flush)
        gms = time_graph(_f, args.reps, flush)
        record("DinoLoss", f"bs {bs} x (2 teacher, {nv} student views) x K {k}", ms, bs,
               bytes_=(2 * (2 + nv) + nv) * bs * k * 4, graph_ms=gms,
               cpu_s=cpu_time(lambda: O.dino(teacher.numpy(), student.numpy(), 0.1, 0.04, center.numpy())),
               note="bytes: fwd reads teacher + student, bwd reads them again and writes dstudent")
    # PirlLoss at the reference's shape (bs 256, 1000 negatives, d 128, tau 0.07) and with a 65536-row negative set
    for (n, k) in ((256, 1000), (256, 65536)):
        img, patch = randn(0, n, 128), randn(1, n, 128)
        mp, mn = unit(0.6 * unit(img) + 0.4 * unit(randn(2, n, 128))), unit(randn(3, k, 128))
        a, b, mpd, mnd = img.to(dev).requires_grad_(True), patch.to(dev).requires_grad_(True), mp.to(dev), mn.to(dev)
        fn_p = S.PirlLoss(True, 0.07, 0.5)
        _f = fwd_bwd(lambda: fn_p(a, b, mpd, mnd), a, b)
        ms = time_gpu(_f, args.reps, flush)
        gms = time_graph(_f, args.reps, flush)
        record("PirlLoss", f"{n} x {k} negatives x 128 tau=0.07", ms, n, bytes_=k * 128 * 4 + 5 * n * 128 * 4, graph_ms=gms,
               cpu_s=cpu_time(lambda: O.pirl(img.numpy(), patch.numpy(), mp.numpy(), mn.numpy(), True, 0.07, 0.5)),
               note="bytes: negatives read once (shared by both heads) + the row-wise inputs / gradients")
    torch.manual_seed(0)
    n_params = 11_200_000   # ~ResNet-18 sized network split into 62 tensors of mixed sizes
    sizes = [64 * 3 * 9, 64, 64] + [n_params // 60] * 58 + [512 * 1000]
    ema_tgt = [torch.randn(n, device=dev) for n in sizes]
    ema_src = [torch.randn(n, device=dev) for n in sizes]
    up = S.EmaUpdater(ema_tgt, ema_src)
    tot = sum(sizes)
    _f = lambda: up.step(0.99)
    ms = time_gpu(_f, args.reps, flush)
    gms = time_graph(_f, args.reps, flush)
    record("EmaUpdater.step (momentum_update)", f"{len(sizes)} tensors, {tot / 1e6:.1f} M parameters", ms, tot,
           bytes_=3 * tot * 4, graph_ms=gms,
           cpu_s=cpu_time(lambda: [O.ema_update(a.cpu().numpy(), b.cpu().numpy(), 0.99) for a, b in zip(ema_tgt[:3], ema_src[:3])]),
           note="bytes: read target + source, write target; one launch for the whole network (samples = parameters)")

    hdr = (f"| loss | config | fwd+bwd ms (eager) | ms (CUDA-graph replay) | samples/s | bound | achieved | "
           f"frac of {src} peak | CPU port s |")
    print(hdr, file=sys.stderr)
    print("|---|---|---|---|---|---|---|---|---|", file=sys.stderr)
    for r in rows:
        ach = f"{r.get('achieved_tflops', 0):.1f} TFLOP/s" if r.get("bound") == "tensor" else f"{r.get('achieved_gbs', 0):.0f} GB/s"
        gm = f"{r['graph_ms']:.4f}" if r.get("graph_ms") else "-"
        print(f"| {r['loss']} | {r['config']} | {r['ms']:.4f} | {gm} | {r['samples_per_s']:.3g} | {r.get('bound')} | {ach} | "
              f"{r.get('frac', 0):.3f} | {r.get('cpu_port_s', float('nan')):.3g} |", file=sys.stderr)


if __name__ == "__main__":
    main()
