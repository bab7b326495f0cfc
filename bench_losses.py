#!/usr/bin/env python
"""Per-config throughput + roofline fractions for every row of SURVEY.md §8(a)/(f) at BASELINE.json's configs.

    python bench_losses.py [--reps 20] [--no-cpu] [--no-ref-gpu] [--only cfg1,cfg2,...]
        -> one JSON object per line + a markdown table on stderr

`run_all()` is also what `bench.py` embeds as the `per_config` array of its JSON line (N = 1).

Each config is measured four ways on the SAME seeded inputs:
  ms / graph_ms        ours: forward + backward through the public drop-in API, inputs resident in HBM, CUDA events on
                       the current stream, L2 flushed (256 MiB write) between repetitions, median; eager (includes the
                       Python / launch cost per call) and as a CUDA-graph replay of the same step (device time; the
                       roofline fraction is taken from it)
  torch_eager_b200_ms  the REFERENCE's own classes (utils/losses.py, unmodified, staged by oracle/make_ref.py) run
                       eagerly in fp32 on the same B200 - the honest GPU comparator SURVEY §2 / §8(d) names
  cpu_baseline         the reference's own classes on the box's host cores (all threads), fwd + bwd, 1 warm-up +
                       a few timed repetitions; kind = "reference".  Without a staged reference: the numpy-fp64
                       oracle port, kind = "port".
The oracle / reference are test and baseline infrastructure: only these comparison legs load them.
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


_FLUSH_READ = {}


def l2_flush(flush):
    """Evict the L2 between timed repetitions: write a 256 MiB buffer (> the 126 MB L2), then READ a second 256 MiB
    buffer so that the cache is left cold AND clean.  After the write alone every L2 line is dirty, and the first
    ~126 MB a kernel reads pay for the write-back of the flush buffer on top of their own traffic (measured: the
    leading read-only kernel of a row ran at 3-3.5 TB/s after a write-only flush) - an artefact of the flush, not of
    the kernel.  BENCH_FLUSH=write restores the write-only flush."""
    flush.zero_()
    if os.environ.get("BENCH_FLUSH", "clean") == "write":
        return
    key = flush.device.index
    buf = _FLUSH_READ.get(key)
    if buf is None:
        buf = (torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=flush.device), torch.empty(1, dtype=torch.float32, device=flush.device))
        _FLUSH_READ[key] = buf
    torch.sum(buf[0], dim=(0,), keepdim=True, out=buf[1])


def time_gpu(fn, reps, flush, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        l2_flush(flush)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def time_host(fn, n=50):
    """Host (Python + driver) microseconds per step: n back-to-back calls with no synchronisation in between, timed
    on the host clock - the issue cost a training loop pays per call when the GPU is not the bottleneck.  (n is small
    enough that the launch queue never fills for the rows where the device time exceeds the host time.)"""
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / n * 1e6


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
            l2_flush(flush)
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


PROFILE_ONLY = False   # --profile: minimal launch count per row (for ncu)
TIMELINE = None   # file object: when set, legs() also prints the kernel timeline of one graph replay of each row


def graph_timeline(fn, flush, title, out):
    """Per-kernel start / gap / duration of ONE CUDA-graph replay of `fn` (torch.profiler = CUPTI; L2 flushed before)."""
    from torch.profiler import ProfilerActivity, profile
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
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                l2_flush(flush)
                g.replay()
                torch.cuda.synchronize()
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        evs.sort(key=lambda e: e.time_range.start)
        # the last replay = everything after the last flush (256 MiB FillFunctor<unsigned char> launch [+ the read-back
        # reduce_kernel of l2_flush])
        cut = max(i for i, e in enumerate(evs) if "FillFunctor<unsigned char>" in e.name)
        for i in range(cut + 1, min(cut + 4, len(evs))):   # [memset of the reduce output,] the 256 MiB read-back
            if "reduce_kernel" in evs[i].name and evs[i].time_range.end - evs[i].time_range.start > 20:
                cut = i
        step = evs[cut + 1:]
        if not step:
            return
        t0 = step[0].time_range.start
        print(f"## {title}: {len(step)} device activities, {step[-1].time_range.end - t0:.1f} us first start -> last end",
              file=out)
        prev = t0
        for e in step:
            s_, dur = e.time_range.start - t0, e.time_range.end - e.time_range.start
            print(f"  +{s_:8.1f} us  gap {e.time_range.start - prev:6.1f}  dur {dur:8.1f}  {e.name[:110]}", file=out)
            prev = e.time_range.end
        out.flush()
    except Exception as e:  # noqa: BLE001
        print(f"## {title}: timeline failed: {type(e).__name__}: {str(e)[:160]}", file=out)
        torch.cuda.synchronize()


def time_cpu(fn, budget_s=4.0, max_reps=5):
    """1 warm-up + up to max_reps timed repetitions within ~budget_s; mean seconds."""
    fn()
    ts = []
    t_start = time.perf_counter()
    while len(ts) < max_reps and (not ts or time.perf_counter() - t_start < budget_s):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return sum(ts) / len(ts), len(ts)


def fwd_bwd(loss_fn, *tensors):
    def run():
        for t in tensors:
            if t.requires_grad:
                t.grad = None
        loss_fn().backward()
    return run


def run_all(dev=None, reps=20, cpu=True, ref_gpu=True, only=None, emit=None):
    """Returns the list of per-config rows (dicts).  `only`: iterable of row tags to run (None = all)."""
    import ssv_b200 as S
    dev = dev or torch.device("cuda", torch.cuda.current_device())
    R = None
    O = None
    if cpu or ref_gpu:
        from oracle import ref_loader
        R = ref_loader.load()
    if cpu and R is None:
        from oracle import ssl_oracle as O  # noqa: N811
    ncores = os.cpu_count()
    if cpu:
        torch.set_num_threads(ncores)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    tf_peak, hbm_peak, src = peaks()
    rows = []
    want = set(only) if only else None

    def record(tag, name, cfg, ms, samples, flops=None, bytes_=None, graph_ms=None, cpu_res=None, ref_ms=None, note="",
               host_us=None, ref_host_us=None):
        t = graph_ms if graph_ms else ms   # roofline fraction from the device time (graph replay) when available
        if flops:
            ach = flops / (t * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                    "traffic": None, "algorithmic": flops, "peak_source": f"{src} burst bf16"}
        else:
            ach = bytes_ / (t * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                    "traffic": None, "algorithmic": bytes_, "peak_source": f"{src} copy bandwidth"}
        r = {"tag": tag, "loss": name, "config": cfg, "ms": ms, "graph_ms": graph_ms,
             "host_us_per_step": host_us, "torch_eager_host_us_per_step": ref_host_us,
             "samples_per_s": samples / (t * 1e-3), "roofline": roof, "torch_eager_b200_ms": ref_ms,
             "speedup_vs_torch_eager_b200": (ref_ms / ms) if ref_ms else None, "cpu_baseline": None, "note": note}
        if cpu_res is not None:
            sec, nrep, kind = cpu_res
            r["cpu_baseline"] = {"value": samples / sec, "unit": "samples/s", "seconds": sec, "reps": nrep,
                                 "cores": ncores, "kind": kind,
                                 "sample": "the full config, fwd+bwd, fp32, all host threads"
                                 if kind == "reference" else "the full config, numpy fp64 oracle port"}
        rows.append(r)
        if emit:
            emit(r)
        return r

    def legs(ours, samples, ref_cpu=None, port_cpu=None, ref_cuda=None, reps_=None, **kw):
        rp = reps_ or reps
        if PROFILE_ONLY:   # under ncu: two eager steps per row and nothing else (every launch is replayed by the profiler)
            ours()
            l2_flush(flush)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ours()
            e1.record()
            e1.synchronize()
            return dict(ms=e0.elapsed_time(e1), graph_ms=None, ref_ms=None, cpu_res=None, samples=samples, **kw)
        ms = time_gpu(ours, rp, flush)
        gms = time_graph(ours, rp, flush)
        host_us = time_host(ours)
        ref_host_us = None
        if TIMELINE is not None:
            legs.count = getattr(legs, "count", 0) + 1
            graph_timeline(ours, flush, f"row {legs.count} (eager {ms * 1e3:.1f} us, graph "
                           f"{(gms or 0) * 1e3:.1f} us)", TIMELINE)
        ref_ms = None
        if ref_gpu and R is not None and ref_cuda is not None:
            try:
                ref_ms = time_gpu(ref_cuda, max(3, rp // 4), flush, warm=2)
                if ref_ms < 5.0:
                    ref_host_us = time_host(ref_cuda, n=20)
            except Exception as e:  # noqa: BLE001
                print(f"[reference on cuda failed: {type(e).__name__}: {str(e)[:160]}]", file=sys.stderr)
                torch.cuda.synchronize()
        cpu_res = None
        if cpu:
            if R is not None and ref_cpu is not None:
                sec, nrep = time_cpu(ref_cpu)
                cpu_res = (sec, nrep, "reference")
            elif port_cpu is not None and O is not None:
                sec, nrep = time_cpu(port_cpu, budget_s=2.0, max_reps=2)
                cpu_res = (sec, nrep, "port")
        return dict(ms=ms, graph_ms=gms, ref_ms=ref_ms, cpu_res=cpu_res, samples=samples, host_us=host_us,
                    ref_host_us=ref_host_us, **kw)

    def on(tag):
        return want is None or tag in want

    def req(*ts):
        return [t.clone().requires_grad_(True) for t in ts]

    # ---- cfg1: SimCLR NT-Xent 2 x 256 x 128, tau 0.5 (the reference's own CPU-runnable case)
    if on("cfg1"):
        zi, zj = randn(0, 256, 128), randn(1, 256, 128)
        a, b = zi.to(dev).requires_grad_(True), zj.to(dev).requires_grad_(True)
        fn = S.SimclrLoss(True, 0.5)
        ca, cb = req(zi, zj)
        ga, gb = req(zi.to(dev), zj.to(dev))
        rf = R.losses.SimclrLoss(True, 0.5) if R else None
        record("cfg1", "SimclrLoss", "cfg1 2x256x128 tau=0.5", flops=6 * 512 ** 2 * 128, note="latency-bound (4 CTAs)",
               **legs(fwd_bwd(lambda: fn(a, b), a, b), 256,
                      ref_cpu=(fwd_bwd(lambda: rf(ca, cb), ca, cb) if R else None),
                      port_cpu=lambda: O.ntxent(zi.numpy(), zj.numpy(), True, 0.5),
                      ref_cuda=(fwd_bwd(lambda: rf(ga, gb), ga, gb) if R else None)))

    # ---- NT-Xent mid sizes (scaling series; the reference still fits on the GPU here)
    for n in (2048, 8192):
        if not on(f"ntx{n}"):
            continue
        zi, zj = randn(0, n, 128), randn(1, n, 128)
        a, b = zi.to(dev).requires_grad_(True), zj.to(dev).requires_grad_(True)
        fn = S.SimclrLoss(True, 0.5)
        ga, gb = req(zi.to(dev), zj.to(dev))
        ca, cb = req(zi, zj)
        rf = R.losses.SimclrLoss(True, 0.5) if R else None
        record(f"ntx{n}", "SimclrLoss", f"{n}x128 tau=0.5", flops=6 * (2 * n) ** 2 * 128,
               **legs(fwd_bwd(lambda: fn(a, b), a, b), n,
                      ref_cpu=(fwd_bwd(lambda: rf(ca, cb), ca, cb) if (R and n <= 2048) else None),
                      ref_cuda=(fwd_bwd(lambda: rf(ga, gb), ga, gb) if R else None)))
        del ga, gb

    # ---- cfg2: MoCo 256 queries x 65536-entry queue x 128 + enqueue (device-resident bank, bf16 shadow)
    if on("cfg2"):
        n, k, d = 256, 65536, 128
        bank = S.MemoryBank(k, d)
        fill = randn(5, k, d).to(dev)
        for i in range(4):
            bank.add_batch(fill[i * 16384:(i + 1) * 16384])
        q, kk = randn(0, n, d), randn(1, n, d)
        a, b = q.to(dev).requires_grad_(True), kk.to(dev).requires_grad_(True)
        fn_m = S.MocoLoss(True, 0.07)
        kd = b.detach()

        def ours_moco():
            a.grad = None
            b.grad = None
            fn_m(a, b, bank.get_vectors()).backward()
            bank.add_batch(kd)

        ref_cpu = ref_cuda = None
        mem_np = None
        if R and R.MemoryBank is not None:
            rbank = R.MemoryBank(k, d)
            rbank.bank = unit(fill.cpu())          # the reference bank lives on the host (models/moco.py:25-29)
            rfm = R.losses.MocoLoss(True, 0.07)
            cq, ck = req(q, kk)
            gq, gk = req(q.to(dev), kk.to(dev))

            def ref_step(qq, kq, device):
                def run():
                    qq.grad = None
                    kq.grad = None
                    rfm(qq, kq, rbank.get_vectors().to(device)).backward()   # models/moco.py:117
                    rbank.add_batch(kq.detach())                             # models/moco.py:124
                return run
            ref_cpu, ref_cuda = ref_step(cq, ck, torch.device("cpu")), ref_step(gq, gk, dev)
        else:
            mem_np = bank.get_vectors().cpu().numpy()
        record("cfg2", "MocoLoss + MemoryBank.add_batch", "cfg2 256x65536x128 tau=0.07, enqueue 256 keys",
               bytes_=2 * k * d * 4 + 2 * n * d * 4,
               note="algorithmic bytes = queue read twice as fp32 (67.1 MB) + the enqueue; kernels read the bf16 shadow; "
                    "the reference step includes its per-step full-queue H2D copy and per-row enqueue loop",
               **legs(ours_moco, n, ref_cpu=ref_cpu, ref_cuda=ref_cuda,
                      port_cpu=lambda: (O.moco(q.numpy(), kk.numpy(), mem_np, True, 0.07),
                                        O.ring_enqueue(mem_np, 0, kk.numpy(), True))))
        mem = bank.get_vectors()
        record("cfg2-loss", "MocoLoss", "cfg2 256x65536x128 tau=0.07 (loss only)", bytes_=2 * k * d * 4,
               **legs(fwd_bwd(lambda: fn_m(a, b, mem), a, b), n))
        record("cfg2-enqueue", "MemoryBank.add_batch", "cfg2 enqueue 256x128 into 65536", bytes_=2 * n * d * 4,
               note="launch-latency bound", **legs(lambda: bank.add_batch(kd), n))
        del fill

    # ---- cfg3: Barlow Twins 2048 x 8192
    if on("cfg3"):
        n, d = 2048, 8192
        g = torch.Generator().manual_seed(7)
        sig, mu = torch.rand(d, generator=g) * 1.5 + 0.5, torch.randn(d, generator=g)
        zi = randn(0, n, d) * sig + mu
        zj = 0.7 * zi + 0.3 * (randn(1, n, d) * sig + mu)
        a, b = zi.to(dev).requires_grad_(True), zj.to(dev).requires_grad_(True)
        fn_b = S.BarlowLoss(False, 0.005)
        rfb = R.losses.BarlowLoss(False, 0.005) if R else None
        ca, cb = req(zi, zj)
        ga, gb = req(zi.to(dev), zj.to(dev))
        record("cfg3", "BarlowLoss", "cfg3 2048x8192 lambda=0.005", flops=6 * n * d * d,
               **legs(fwd_bwd(lambda: fn_b(a, b), a, b), n, reps_=max(5, reps // 3),
                      ref_cpu=(fwd_bwd(lambda: rfb(ca, cb), ca, cb) if R else None),
                      port_cpu=lambda: O.barlow(zi.numpy(), zj.numpy(), False, 0.005),
                      ref_cuda=(fwd_bwd(lambda: rfb(ga, gb), ga, gb) if R else None)))
        del a, b, ga, gb, ca, cb

    # ---- cfg4: Sinkhorn-Knopp 4096 x 3000, 3 iters, eps 0.05
    if on("cfg4"):
        bsz, kp = 4096, 3000
        scores = (unit(randn(0, bsz, 128)) @ unit(randn(1, kp, 128)).t()).contiguous()
        sd = scores.to(dev)
        fn_s = S.SwavLoss(0.1, 0.05, 3)
        ref_cpu = ref_cuda = None
        if R:
            rs_cpu = R.losses.SwavLoss(0.1, 0.05, 3)
            rs_cpu.device = torch.device("cpu")      # utils/losses.py:208 picks cuda at construction on a GPU box
            rs_gpu = R.losses.SwavLoss(0.1, 0.05, 3)
            rs_gpu.device = dev
            ref_cpu = lambda: rs_cpu.compute_codes_sinkhorn(scores)  # noqa: E731
            ref_cuda = lambda: rs_gpu.compute_codes_sinkhorn(sd)     # noqa: E731
        record("cfg4", "SwavLoss.compute_codes_sinkhorn", "cfg4 4096x3000 3 iters eps=0.05", bytes_=2 * bsz * kp * 4,
               **legs(lambda: fn_s.compute_codes_sinkhorn(sd), bsz, ref_cpu=ref_cpu, ref_cuda=ref_cuda,
                      port_cpu=lambda: O.sinkhorn(scores.numpy(), 0.05, 3)))

    # ---- SwAV full loss (reference shapes: 512 live + 3000 bank rows, 3000 prototypes, d = 128)
    if on("swav"):
        z1, z2 = unit(randn(0, 512, 128)), unit(randn(1, 512, 128))
        c, bk = unit(randn(2, 3000, 128)), unit(randn(3, 3000, 128))
        a, b, pc = z1.to(dev).requires_grad_(True), z2.to(dev).requires_grad_(True), c.to(dev).requires_grad_(True)
        bkd = bk.to(dev)
        fn_s = S.SwavLoss(0.1, 0.05, 3)
        ref_cpu = ref_cuda = None
        if R:
            rs_cpu = R.losses.SwavLoss(0.1, 0.05, 3)
            rs_cpu.device = torch.device("cpu")
            rs_gpu = R.losses.SwavLoss(0.1, 0.05, 3)
            rs_gpu.device = dev
            c1, c2, cc = req(z1, z2, c)
            g1, g2, gc = req(z1.to(dev), z2.to(dev), c.to(dev))
            ref_cpu = fwd_bwd(lambda: rs_cpu(c1, c2, cc, bk), c1, c2, cc)
            ref_cuda = fwd_bwd(lambda: rs_gpu(g1, g2, gc, bkd), g1, g2, gc)
        bp = 3512
        record("swav", "SwavLoss", "512+3000 bank x 3000 prototypes x 128", bytes_=6 * bp * 3000 * 4 * 2,
               **legs(fwd_bwd(lambda: fn_s(a, b, pc, bkd), a, b, pc), 512, ref_cpu=ref_cpu, ref_cuda=ref_cuda,
                      port_cpu=lambda: O.swav(z1.numpy(), z2.numpy(), c.numpy(), bk.numpy())))

    # ---- BYOL MSE / SimSiam
    for (n, d) in ((32768, 128), (4096, 1024), (524288, 128)):
        # the last shape is 268 MB per operand (> the 126 MB L2): the streaming rate of the kernels without the
        # launch / ramp share that dominates the 17 MB reference-sized rows
        big = n * d * 4 > 128 * 1024 * 1024
        rtag = "rowdot_big" if big else f"rowdot{d}"
        if not on(rtag):
            continue
        o, t = unit(randn(0, n, d)), unit(randn(1, n, d))
        a, b = o.to(dev).requires_grad_(True), t.to(dev)
        fn_e = S.MSELoss()
        rmse = torch.nn.MSELoss()                 # BYOL's loss IS torch.nn.MSELoss (models/byol.py:89)
        co, = req(o)
        go, = req(o.to(dev))
        record(rtag, "MSELoss (BYOL)", f"{n}x{d}", bytes_=(2 + 3) * n * d * 4,
               **legs(fwd_bwd(lambda: fn_e(a, b), a), n,
                      ref_cpu=fwd_bwd(lambda: rmse(co, t), co) if (R and not big) else None,
                      port_cpu=lambda: O.mse(o.numpy(), t.numpy()),
                      ref_cuda=fwd_bwd(lambda: rmse(go, b), go) if R else None))
        b2 = t.to(dev).requires_grad_(True)
        fn_ss = S.SimSiamLoss()
        rss = R.losses.SimSiamLoss() if R else None
        co2, ct2 = req(o, t)
        go2, gt2 = req(o.to(dev), t.to(dev))
        record(rtag, "SimSiamLoss", f"{n}x{d}", bytes_=(2 + 4) * n * d * 4,
               **legs(fwd_bwd(lambda: fn_ss(a, b2), a, b2), n,
                      ref_cpu=fwd_bwd(lambda: rss(co2, ct2), co2, ct2) if (R and not big) else None,
                      port_cpu=lambda: O.simsiam(o.numpy(), t.numpy()),
                      ref_cuda=fwd_bwd(lambda: rss(go2, gt2), go2, gt2) if R else None))

    # ---- ReLIC
    for n in (512, 4096):
        if not on(f"relic{n}"):
            continue
        zi, zj, zo = randn(0, n, 128), randn(1, n, 128), randn(2, n, 128)
        a, b, c3 = (x.to(dev).requires_grad_(True) for x in (zi, zj, zo))
        fn_r = S.RelicLoss(True, 1.0, 0.5)
        rr = R.losses.RelicLoss(True, 1.0, 0.5) if R else None
        ci, cj, co = req(zi, zj, zo)
        gi, gj, go = req(zi.to(dev), zj.to(dev), zo.to(dev))
        record(f"relic{n}", "RelicLoss", f"{n}x128 tau=1 alpha=0.5", flops=6 * (2 * n) ** 2 * 128,
               note="FLOPs of the contrastive part only; the KL term adds 9*N*d*4 bytes",
               **legs(fwd_bwd(lambda: fn_r(a, b, c3), a, b, c3), n,
                      ref_cpu=fwd_bwd(lambda: rr(ci, cj, co), ci, cj, co) if R else None,
                      port_cpu=lambda: O.relic(zi.numpy(), zj.numpy(), zo.numpy(), True, 1.0, 0.5),
                      ref_cuda=fwd_bwd(lambda: rr(gi, gj, go), gi, gj, go) if R else None))

    # ---- DinoLoss, PirlLoss, the parameter EMA
    for (bs, nv, k) in ((64, 8, 1024), (1024, 8, 4096)):
        if not on(f"dino{bs}"):
            continue
        teacher, student, center = randn(0, bs, 2, k), randn(1, bs, nv, k), 0.1 * randn(2, k)
        t, st, c = teacher.to(dev), student.to(dev).requires_grad_(True), center.to(dev)
        fn_d = S.DinoLoss()
        rd = R.losses.DinoLoss() if R else None
        cs, = req(student)
        gs, = req(student.to(dev))
        record(f"dino{bs}", "DinoLoss", f"bs {bs} x (2 teacher, {nv} student views) x K {k}",
               bytes_=(2 * (2 + nv) + nv) * bs * k * 4,
               note="bytes: fwd reads teacher + student, bwd reads them again and writes dstudent",
               **legs(fwd_bwd(lambda: fn_d(t, st, 0.1, 0.04, c), st), bs,
                      ref_cpu=fwd_bwd(lambda: rd(teacher, cs, 0.1, 0.04, center), cs) if (R and bs <= 64) else None,
                      port_cpu=(lambda: O.dino(teacher.numpy(), student.numpy(), 0.1, 0.04, center.numpy())) if bs <= 64 else None,
                      ref_cuda=fwd_bwd(lambda: rd(t, gs, 0.1, 0.04, c), gs) if R else None))
    for (n, k) in ((256, 1000), (256, 65536)):
        if not on(f"pirl{k}"):
            continue
        img, patch = randn(0, n, 128), randn(1, n, 128)
        mp, mn = unit(0.6 * unit(img) + 0.4 * unit(randn(2, n, 128))), unit(randn(3, k, 128))
        a, b, mpd, mnd = img.to(dev).requires_grad_(True), patch.to(dev).requires_grad_(True), mp.to(dev), mn.to(dev)
        fn_p = S.PirlLoss(True, 0.07, 0.5)
        rp_ = R.losses.PirlLoss(True, 0.07, 0.5) if R else None
        ci, cp = req(img, patch)
        gi, gp = req(img.to(dev), patch.to(dev))
        record(f"pirl{k}", "PirlLoss", f"{n} x {k} negatives x 128 tau=0.07", bytes_=k * 128 * 4 + 5 * n * 128 * 4,
               note="bytes: negatives read once (shared by both heads) + the row-wise inputs / gradients",
               **legs(fwd_bwd(lambda: fn_p(a, b, mpd, mnd), a, b), n,
                      ref_cpu=fwd_bwd(lambda: rp_(ci, cp, mp, mn), ci, cp) if R else None,
                      port_cpu=lambda: O.pirl(img.numpy(), patch.numpy(), mp.numpy(), mn.numpy(), True, 0.07, 0.5),
                      ref_cuda=fwd_bwd(lambda: rp_(gi, gp, mpd, mnd), gi, gp) if R else None))
    for ema_tag, n_params, n_mid in (("ema", 11_200_000, 58), ("ema_big", 25_600_000, 158)):
        if not on(ema_tag):
            continue
        torch.manual_seed(0)
        # ~ResNet-18 (62 tensors) / ~ResNet-50 (162 tensors, 307 MB of traffic > L2) sized networks, mixed tensor sizes
        sizes = [64 * 3 * 9, 64, 64] + [n_params // (n_mid + 2)] * n_mid + [512 * 1000]
        ema_tgt = [torch.randn(n, device=dev) for n in sizes]
        ema_src = [torch.randn(n, device=dev) for n in sizes]
        up = S.EmaUpdater(ema_tgt, ema_src)
        tot = sum(sizes)
        r_tgt = [x.clone() for x in ema_tgt]

        def ref_ema():  # the reference's loop: models/moco.py:108-111 (three eager ops per tensor)
            for tp, sp in zip(r_tgt, ema_src):
                tp.data = 0.99 * tp.data + (1.0 - 0.99) * sp.data
        c_tgt, c_src = [x.cpu() for x in ema_tgt], [x.cpu() for x in ema_src]

        def ref_ema_cpu():
            for tp, sp in zip(c_tgt, c_src):
                tp.data = 0.99 * tp.data + (1.0 - 0.99) * sp.data
        record(ema_tag, "EmaUpdater.step (momentum_update)", f"{len(sizes)} tensors, {tot / 1e6:.1f} M parameters",
               bytes_=3 * tot * 4,
               note="bytes: read target + source, write target; one launch for the whole network (samples = parameters)",
               **legs(lambda: up.step(0.99), tot, ref_cpu=ref_ema_cpu if R else None, ref_cuda=ref_ema if R else None))
    return rows


def table(rows, file=sys.stderr):
    print("| loss | config | ours eager ms | ours graph ms | ours host us / step | reference eager on B200 ms | "
          "speed-up (eager/eager) | bound | achieved | frac of measured peak | CPU arm (kind, cores) s |", file=file)
    print("|---|---|---|---|---|---|---|---|---|---|---|", file=file)
    for r in rows:
        rf = r["roofline"]
        ach = f"{rf['achieved']:.1f} {rf['unit']}"
        gm = f"{r['graph_ms']:.4f}" if r.get("graph_ms") else "-"
        rm = f"{r['torch_eager_b200_ms']:.3f}" if r.get("torch_eager_b200_ms") else "-"
        sp = f"{r['speedup_vs_torch_eager_b200']:.1f}x" if r.get("speedup_vs_torch_eager_b200") else "-"
        cb = r.get("cpu_baseline")
        cs = f"{cb['seconds']:.4g} ({cb['kind']}, {cb['cores']})" if cb else "-"
        hu = f"{r['host_us_per_step']:.0f}" if r.get("host_us_per_step") else "-"
        print(f"| {r['loss']} | {r['config']} | {r['ms']:.4f} | {gm} | {hu} | {rm} | {sp} | {rf['bound']} | {ach} | "
              f"{rf['frac']:.3f} | {cs} |", file=file)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true")
    ap.add_argument("--only", default=None, help="comma-separated row tags (cfg1,cfg2,cfg3,cfg4,swav,rowdot128,...)")
    ap.add_argument("--timeline", default=None, help="also write the per-kernel timeline of one graph replay per row here")
    ap.add_argument("--profile", action="store_true", help="two eager steps per row and nothing else (run under ncu)")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    if args.profile:
        global PROFILE_ONLY
        PROFILE_ONLY = True
    if args.timeline:
        global TIMELINE
        TIMELINE = open(args.timeline, "w")
    rows = run_all(torch.device("cuda", 0), reps=args.reps, cpu=not args.no_cpu, ref_gpu=not args.no_ref_gpu,
                   only=args.only.split(",") if args.only else None,
                   emit=lambda r: print(json.dumps(r), flush=True))
    table(rows)


if __name__ == "__main__":
    main()
