"""Turn ncu artefacts brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python profiles/summarize_ncu.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches.txt
    python profiles/summarize_ncu.py report   gpurun_out/prof_bwd_r1.ncu-rep > profiles/r1_sim_bwd.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_active.avg", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "smsp__inst_executed.sum",
]


def launches(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            v = float(d["Metric Value"].replace(",", ""))
            u = d["Metric Unit"]
            v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
            agg.setdefault(d["Kernel Name"].split("(")[0][:70], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none : {path}")
    print(f"# total device time of listed launches: {tot / 1e3:.3f} ms (cold-cache, serialised: compare SHARES)")
    print(f"{'kernel':72s} {'launches':>8s} {'mean_us':>10s} {'share_%':>8s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:72s} {len(v):8d} {sum(v) / len(v):10.1f} {sum(v) / tot * 100:8.2f}")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        print(f"# ncu --set full --clock-control none : {path}")
        print(f"kernel: {d.get('Kernel Name', ('', '?'))[1][:100]}")
        for k in KEYS:
            if k in d:
                print(f"{k:90s} {d[k][0]:14s} {d[k][1]}")
        stalls = []
        for h, (u, v) in d.items():
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
                stalls.append((float(v.replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        print("top warp-stall reasons (warps stalled per issue-active cycle):")
        for v, h in sorted(stalls, reverse=True)[:8]:
            print(f"    {h:30s} {v:8.2f}")


def metrics(path):
    """Per-kernel means of every metric in a `ncu --metrics a,b,c --csv` log (one line per launch and metric).  Launches
    of one kernel on very different problem sizes (durations more than 1.8x apart) are listed as separate groups."""
    rows = list(csv.reader(open(path, errors="ignore")))
    hdr = None
    per = collections.OrderedDict()   # (kernel, launch id) -> metric -> value
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            u = d["Metric Unit"]
            scale = {"ns": 1e-3, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3, "Kbyte": 1e-3, "Gbyte": 1e3, "byte": 1e-6}.get(u, 1.0)
            per.setdefault((d["Kernel Name"].split("(")[0][:60], d["ID"]), collections.OrderedDict())[d["Metric Name"]] = v * scale
    names = []
    for m in per.values():
        for k in m:
            if k not in names:
                names.append(k)
    short = {"gpu__time_duration.sum": "us", "dram__bytes_read.sum": "rd_MB", "dram__bytes_write.sum": "wr_MB",
             "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_%", "lts__t_sector_hit_rate.pct": "l2hit_%",
             "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_%", "launch__grid_size": "grid",
             "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_%"}
    by_kernel = collections.OrderedDict()
    for (kname, _), m in per.items():
        by_kernel.setdefault(kname, []).append(m)
    groups = []
    for kname, launches_ in by_kernel.items():
        launches_.sort(key=lambda m: m.get("gpu__time_duration.sum", 0.0))
        cur = [launches_[0]]
        for m in launches_[1:]:
            if m.get("gpu__time_duration.sum", 0.0) > 1.8 * cur[-1].get("gpu__time_duration.sum", 1e-9):
                groups.append((kname, cur))
                cur = []
            cur.append(m)
        groups.append((kname, cur))
    print(f"# ncu --metrics ... --clock-control none : {path}   (means over the launches of each kernel / size group; time in us, "
          "bytes in MB, GB/s = (rd + wr) / time; cold cache and serialised launches: compare with the in-graph timelines)")
    print(f"{'kernel':62s} {'n':>4s} " + " ".join(f"{short.get(k, k[:10]):>9s}" for k in names) + f" {'GB/s':>9s}")
    mean = lambda g, k: sum(m.get(k, 0.0) for m in g) / len(g)  # noqa: E731
    for kname, g in sorted(groups, key=lambda kg: -mean(kg[1], "gpu__time_duration.sum") * len(kg[1])):
        t = mean(g, "gpu__time_duration.sum")
        bw = (mean(g, "dram__bytes_read.sum") + mean(g, "dram__bytes_write.sum")) / t * 1e3 if t > 0 else float("nan")
        print(f"{kname:62s} {len(g):4d} " + " ".join(f"{mean(g, k):9.2f}" for k in names) + f" {bw:9.0f}")


if __name__ == "__main__":
    {"launches": launches, "report": report, "metrics": metrics}[sys.argv[1]](sys.argv[2])
