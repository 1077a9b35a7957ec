#!/usr/bin/env python
"""Round-2 profile of one training step of the default bench workload, run ON THE GPU BOX:

  python tools/profile_r2.py <tag>          -> gpurun_out/<tag>_*

 1. launch list of `bench.py --steps 2 --warmup 1` (gpu__time_duration.sum per launch; shares, not absolutes)
 2. ncu --set full over three windows of ONE step that together contain every kernel of the hot path: the start of the
    forward pass (embedding, mask plan, first layers), the forward/backward boundary (last layer, loss head, head backward,
    first backward layers) and the end of the step (embedding gradient, clipping norm, AdamW)
 3. one table: per distinct kernel — launches per step, share of the step, tensor-pipe %, DRAM GB/s and % of peak, issue
    slots busy, the two largest stall reasons.
The windows are found from the launch list itself, so the script keeps working when the launch sequence changes."""
import csv
import gzip
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
BENCH = ["python", os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--batches", "1", "--no-e2e", "--no-fwd-only",
         "--no-cpu-baseline"]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("ggpt::", "")


def launch_list(tag):
    path = os.path.join(OUT, f"{tag}_launches.csv")
    subprocess.run(["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "--csv", "--log-file", path] + BENCH,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=False)
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    for r in rd:
        if r[im] == "gpu__time_duration.sum":
            rows.append((short(r[ik]), float(r[iv].replace(",", ""))))
    return rows


def capture(tag, win, skip, count):
    rep = f"/tmp/{tag}_{win}"
    subprocess.run(["ncu", "--set", "full", "--clock-control", "none", "-s", str(skip), "-c", str(count), "-f", "-o", rep] + BENCH,
                   stdout=open(os.path.join(OUT, f"{tag}_ncu_{win}.log"), "w"), stderr=subprocess.STDOUT, check=False)
    raw = subprocess.run(["ncu", "-i", rep + ".ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    with gzip.open(os.path.join(OUT, f"{tag}_ncu_{win}_raw.csv.gz"), "wt") as f:
        f.write(raw)
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return []
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    if not stall:
        stall = [h for h in hdr if "issue_stalled" in h and h.endswith(".pct")]
    out = []
    for r in rows[2:]:
        def g(k, default=0.0):
            try:
                return float(r[idx[k]].replace(",", ""))
            except Exception:
                return default
        dur_ns = g("gpu__time_duration.sum")
        unit = rows[1][idx["gpu__time_duration.sum"]] if "gpu__time_duration.sum" in idx else "ns"
        dur_us = dur_ns * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1e-3)

        def bytes_of(k):
            v, u = g(k), rows[1][idx[k]] if k in idx else "byte"
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        dram = bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum")
        st = sorted(((g(h), h) for h in stall), reverse=True)[:2]
        clean = lambda h: re.sub(r"smsp__average_warps_issue_stalled_|_per_issue_active.ratio|smsp__warp_issue_stalled_|_per_warp_active.pct", "", h)
        out.append({"kernel": short(r[idx["Kernel Name"]]), "us": dur_us, "dram_bytes": dram,
                    "dram_gbs": dram / (dur_us * 1e-6) / 1e9 if dur_us > 0 else 0.0,
                    "dram_pct": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                    "tensor_pct": g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                    "issue_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "regs": g("launch__registers_per_thread"),
                    "stalls": ", ".join(f"{clean(h)} {v:.1f}" for v, h in st)})
    return out


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    os.makedirs(OUT, exist_ok=True)
    ll = launch_list(tag)
    names = [n for n, _ in ll]
    starts = [i for i, n in enumerate(names) if n.startswith("embed_fwd_kernel")]
    if len(starts) < 2:
        print("could not locate the steps in the launch list", len(ll))
        return
    s2 = starts[1]                                           # first launch of the timed step
    e2 = starts[2] if len(starts) > 2 else len(names)        # (a third step = the per-kernel timing pass of bench.py)
    step = ll[s2:e2]
    total = sum(t for _, t in step)
    agg = {}
    for n, t in step:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
    with open(os.path.join(OUT, f"{tag}_launches_summary.txt"), "w") as f:
        f.write("# one training step of `bench.py` (c2 packed), ncu --metrics gpu__time_duration.sum, cold-cache / serialised: SHARES\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{n:60s} launches {c:4d}  {t / 1e6:9.3f} ms  share {t / total:6.3f}\n")
        f.write(f"total {total / 1e6:.3f} ms over {len(step)} launches\n")
    ce = next(i for i in range(s2, e2) if names[i].startswith("ce_fwd"))
    windows = {"fwd": (s2, 18), "boundary": (ce - 13, 46), "tail": (e2 - 8, 8)}
    only = [a for a in sys.argv[2:] if a in windows]          # e.g. `profile_r2.py <tag> boundary`: that window only
    if only:
        windows = {k: v for k, v in windows.items() if k in only}
    rows = []
    for w, (skip, count) in windows.items():
        rows += capture(tag, w, skip, count)
    best = {}
    for r in rows:                                           # one row per distinct kernel: the longest instance
        if r["kernel"] not in best or r["us"] > best[r["kernel"]]["us"]:
            best[r["kernel"]] = r
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    with open(os.path.join(OUT, f"{tag}_ncu_step_summary.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none, one instance (the longest) of every kernel of one c2-packed training step\n")
        f.write(f"# DRAM GB/s = (dram__bytes_read + dram__bytes_write) / duration; HBM peak {hbm:.0f} GB/s (MEASURED_PEAKS.json)\n")
        f.write(f"{'kernel':58s} {'per step':>8s} {'share':>6s} {'us':>9s} {'tensor%':>8s} {'DRAM GB/s':>10s} {'of peak':>8s} {'issue%':>7s} {'regs':>5s}  top stalls (warps per issue)\n")
        for k, r in sorted(best.items(), key=lambda kv: -agg.get(kv[0], [0, 0.0])[1]):
            c, t = agg.get(k, [0, 0.0])
            f.write(f"{k:58s} {c:8d} {t / total if total else 0:6.3f} {r['us']:9.1f} {r['tensor_pct']:8.1f} {r['dram_gbs']:10.0f} "
                    f"{r['dram_gbs'] / hbm:8.2f} {r['issue_pct']:7.1f} {int(r['regs']):5d}  {r['stalls']}\n")
    # GEMM DRAM traffic per launch (roofline.traffic of bench.py)
    gem = [r for r in rows if r["kernel"].startswith("gemm_kernel")]
    if gem:
        json.dump({"dram_bytes_per_launch": sum(r["dram_bytes"] for r in gem) / len(gem), "launches_sampled": len(gem),
                   "how": "mean of dram__bytes_read.sum + dram__bytes_write.sum over the gemm_kernel launches captured with ncu "
                          "--set full in tools/profile_r2.py (forward, loss head and backward instances of one step)"},
                  open(os.path.join(OUT, f"{tag}_gemm_traffic.json"), "w"))
    print(open(os.path.join(OUT, f"{tag}_ncu_step_summary.txt")).read())


if __name__ == "__main__":
    main()
