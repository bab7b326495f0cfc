#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: NT-Xent fwd+bwd samples/s (and fraction of the tensor roofline)
at global batch 32768 x 128-d, tau = 0.5, normalize = True, on 1/2/4/8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path
    python bench.py --impl reference [...]                          # CPU arm (oracle port, host cores)
    torchrun --nproc-per-node N bench.py --gpus N ...               # N > 1: one rank per GPU, NCCL

One "step" = one forward + backward of SimclrLoss over the global batch (strong scaling: the global batch
is fixed, rank r owns 32768/N rows per view).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "self-supervised-vision_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

N_GLOBAL = 32768          # images per step ("B=32768"): M = 65536 embedding rows
DIM = 128
TAU = 0.5
METRIC = "ntxent_fwd_bwd_samples_per_s"
WORKLOAD = "SimCLR global-batch NT-Xent 32768x128-d (M=65536 rows), tau=0.5, normalize=True (BASELINE configs[4])"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            pk = json.load(f)
        return {"bf16_tflops": pk.get("bf16_tflops", 1590.0), "bf16_tflops_sustained": pk.get("bf16_tflops_sustained"),
                "hbm_gbs": pk.get("hbm_gbs", 6650.0), "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


# ----------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        clocks, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                clocks.append(float(parts[1]))
                mx = float(parts[2])
            except ValueError:
                continue
            for nm, val in zip(names, parts[4:8]):
                if val.lower() == "active":
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if clocks:
            # median over the upper half of the samples == "under load" (idle gaps between steps drop out)
            top = sorted(clocks)[len(clocks) // 2:]
            out.update(sm_mhz=statistics.median(top), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(clocks))
        return out


# ----------------------------------------------------------------------------------------- CPU arm
def cpu_sample(rows=4096, reps=1):
    from oracle import cpu_port
    r = cpu_port.time_ntxent_sample(N_GLOBAL, DIM, TAU, rows=rows, reps=reps, threads=os.cpu_count())
    return r


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import cpu_port
    torch.set_num_threads(os.cpu_count())
    rows = 4096
    g = torch.Generator().manual_seed(0)
    zi = torch.randn(N_GLOBAL, DIM, generator=g)
    zj = torch.randn(N_GLOBAL, DIM, generator=g)
    m = 2 * N_GLOBAL

    def step():
        zhat = torch.nn.functional.normalize(torch.cat([zi, zj]), dim=-1)
        loss_sum, d_rows, d_cols = cpu_port.ntxent_row_slab(zhat, N_GLOBAL, 0, rows, TAU)
        d_cols[:rows] += d_rows
        return float(loss_sum)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = (N_GLOBAL * rows / m) / dt
    sample = (f"row slab: {rows} of the {m} similarity rows x all {m} columns per step (= {rows}/{m} of the full "
              f"6*M^2*d job), torch CPU fp32, fwd + both gradient GEMMs")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": N_GLOBAL, "dim": DIM},
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import ssv_b200
    from ssv_b200 import _cabi
    from ssv_b200.dist import DistributedSimclrLoss

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torchrun (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _cabi.lib()
    _cabi.check(L.ssvb_device_check(), "ssvb_device_check")

    n_local = N_GLOBAL // world
    m = 2 * N_GLOBAL
    # synthetic embeddings: rank r rows seeded 100+2r / 101+2r, made on the host
    gi = torch.Generator().manual_seed(100 + 2 * rank)
    gj = torch.Generator().manual_seed(101 + 2 * rank)
    h_zi = torch.randn(n_local, DIM, generator=gi).pin_memory()
    h_zj = torch.randn(n_local, DIM, generator=gj).pin_memory()
    zi = h_zi.to(dev).requires_grad_(True)
    zj = h_zj.to(dev).requires_grad_(True)
    loss_fn = ssv_b200.SimclrLoss(True, TAU) if world == 1 else DistributedSimclrLoss(True, TAU)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_resident():
        zi.grad = None
        zj.grad = None
        loss = loss_fn(zi, zj)
        loss.backward()
        return loss

    # e2e: what a training loop does — the next step's inputs travel host->device on a side stream (pinned memory,
    # double-buffered) while the current step computes.  EVERY step's H2D copy is issued inside a timed region (the
    # copy for step t+1 is enqueued right after step t's start event), and every step ends with the D2H read of its
    # loss and a stream synchronize, like `loss.item()` in the reference loop (models/simclr.py:95).
    copy_stream = torch.cuda.Stream(device=dev)
    d_in = [(torch.empty(n_local, DIM, device=dev), torch.empty(n_local, DIM, device=dev)) for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    h_loss = torch.empty((), dtype=torch.float32).pin_memory()
    st = {"t": 0}

    def issue_copy(slot, first_use):
        with torch.cuda.stream(copy_stream):
            if not first_use:
                copy_stream.wait_event(consumed[slot])  # the step that read this buffer pair has finished with it
            d_in[slot][0].copy_(h_zi, non_blocking=True)
            d_in[slot][1].copy_(h_zj, non_blocking=True)
            copied[slot].record(copy_stream)

    def step_e2e():
        t = st["t"]
        slot = t & 1
        if t == 0:
            issue_copy(slot, True)          # the very first step waits for its own copy
        issue_copy(slot ^ 1, t == 0)        # inputs of step t+1: overlap with this step's kernels
        torch.cuda.current_stream().wait_event(copied[slot])
        a = d_in[slot][0].detach().requires_grad_(True)
        b = d_in[slot][1].detach().requires_grad_(True)
        loss = loss_fn(a, b)
        loss.backward()
        h_loss.copy_(loss.detach(), non_blocking=True)
        consumed[slot].record()
        torch.cuda.current_stream().synchronize()  # the caller reads the loss (loss.item() in the reference loop)
        st["t"] = t + 1
        return float(h_loss)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, profile=False):
        for _ in range(warmup):
            fn()
        barrier()
        evs = []
        if profile:
            L.ssvb_launch_count(1)
            L.ssvb_profile_enable(1)
        for _ in range(steps):
            flush.zero_()  # evict L2 between timed iterations (not timed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        launches = L.ssvb_launch_count(0) if profile else 0
        prof = {}
        if profile:
            for kind, name in ((0, "sim_fwd"), (1, "sim_bwd")):
                ms, cnt = ctypes.c_double(0), ctypes.c_longlong(0)
                L.ssvb_profile_summary(kind, ctypes.cast(ctypes.pointer(ms), ctypes.c_void_p),
                                       ctypes.cast(ctypes.pointer(cnt), ctypes.c_void_p))
                prof[name] = (ms.value, cnt.value)
            L.ssvb_profile_enable(0)
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / steps, launches, prof

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_step, launches, prof = timed(step_resident, args.steps, args.warmup, profile=True)
    clocks = sampler.stop()
    ms_e2e, _, _ = timed(step_e2e, args.steps, max(3, args.warmup // 2))

    peaks = load_peaks()
    value = N_GLOBAL / (ms_step * 1e-3)
    e2e_value = N_GLOBAL / (ms_e2e * 1e-3)
    # algorithmic FLOPs: 6*M^2*d per step over all ranks; the dominant kernel (sim_bwd_kernel,
    # recompute S + dZ GEMM) carries 4*(M/world)*M*d per launch, sim_fwd_kernel 2*(M/world)*M*d.
    bwd_ms = prof["sim_bwd"][0] / max(prof["sim_bwd"][1], 1)
    fwd_ms = prof["sim_fwd"][0] / max(prof["sim_fwd"][1], 1)
    bwd_flops = 4.0 * (m / world) * m * DIM
    fwd_flops = 2.0 * (m / world) * m * DIM
    achieved = bwd_flops / (bwd_ms * 1e-3) / 1e12
    # denominator: the burst cuBLAS figure, unless the clock record shows the power cap biting during the timed
    # region (a long back-to-back run) - then the sustained figure is the like-for-like one (B200_PROFILING.md)
    capped = "sw_power_cap" in clocks.get("reasons", []) and peaks.get("bf16_tflops_sustained")
    peak = peaks["bf16_tflops_sustained"] if capped else peaks["bf16_tflops"]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if world == 1 and os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("sim_bwd_kernel")  # bytes per launch from the committed ncu capture
    roofline = {"bound": "tensor", "kernel": "sim_bwd_kernel<2,0>", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": f"{peaks['source']} {'sustained (sw_power_cap active)' if capped else 'burst'} bf16 "
                               f"(MEASURED_PEAKS.json)",
                "frac_of_burst": achieved / peaks["bf16_tflops"],
                "kernel_ms": bwd_ms, "flops_per_launch": bwd_flops,
                "fwd_kernel": {"name": "sim_fwd_kernel<2,0>", "ms": fwd_ms,
                               "achieved": fwd_flops / (fwd_ms * 1e-3) / 1e12,
                               "frac": fwd_flops / (fwd_ms * 1e-3) / 1e12 / peak},
                "step": {"flops": 6.0 * m * m * DIM, "achieved": 6.0 * m * m * DIM / (ms_step * 1e-3) / 1e12 / world,
                         "frac": 6.0 * m * m * DIM / (ms_step * 1e-3) / 1e12 / world / peak},
                "frac_of_sustained": (achieved / peaks["bf16_tflops_sustained"]) if peaks.get("bf16_tflops_sustained") else None}

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            r = cpu_sample(rows=4096, reps=8)
            cpu = {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["threads"], "kind": "port",
                   "sample": (f"{r['reps']} row slabs of {r['rows']} of the {r['m']} similarity rows x all columns "
                              f"(each {r['rows']}/{r['m']} of the full job; {r['reps'] * r['rows']}/{r['m']} in total), "
                              f"torch CPU fp32 fwd+bwd, {r['seconds']:.2f} s per slab, {r['total_seconds']:.1f} s timed")}
        line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": WORKLOAD, "global_batch": N_GLOBAL, "dim": DIM, "rows_per_rank": 2 * n_local,
                           "parallelism": f"row-sharded x{world}, all-gather zhat + lse", "l2": "flushed (256 MiB write) between timed iterations"},
                "roofline": roofline, "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": 2 * n_local * DIM * 4 * world, "d2h_bytes_per_step": 4 * world,
                        "pipeline": "H2D of step t+1 (pinned, side stream, double-buffered) overlaps the kernels of step t; "
                                    "every copy is issued inside a timed step; each step ends with the loss D2H + sync"},
                "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
