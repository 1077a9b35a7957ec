"""profiles/<tag>_gemm_traffic.csv (ncu dram__bytes_read/write + duration per GEMM launch of one training step)
-> profiles/r1_gemm_traffic.json, the file bench.py reads roofline.traffic from.   usage: python tools/gemm_traffic_json.py <csv>"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) >= 15 and r[0].isdigit()]
    per = {}
    for r in rows:
        d = per.setdefault(r[0], {"name": r[4]})
        d[r[12]] = float(r[14].replace(",", ""))
    n = len(per)
    rd = sum(v.get("dram__bytes_read.sum", 0.0) for v in per.values())
    wr = sum(v.get("dram__bytes_write.sum", 0.0) for v in per.values())
    by_kernel = {}
    for v in per.values():
        k = by_kernel.setdefault(v["name"].split("(")[0], {"launches": 0, "dram_bytes": 0.0, "ns": 0.0})
        k["launches"] += 1
        k["dram_bytes"] += v.get("dram__bytes_read.sum", 0.0) + v.get("dram__bytes_write.sum", 0.0)
        k["ns"] += v.get("gpu__time_duration.sum", 0.0)
    out = {"source": f"profiles/{os.path.basename(path)} (ncu dram__bytes_read.sum + dram__bytes_write.sum over the {n} GEMM "
                     "launches of one training step)",
           "gemm_launches_per_step": n, "dram_bytes_per_step": rd + wr, "dram_bytes_per_launch": (rd + wr) / max(n, 1),
           "by_instantiation": by_kernel}
    with open(os.path.join(ROOT, "profiles", "r1_gemm_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: out[k] for k in ("gemm_launches_per_step", "dram_bytes_per_step", "dram_bytes_per_launch")}))


if __name__ == "__main__":
    main(sys.argv[1])
