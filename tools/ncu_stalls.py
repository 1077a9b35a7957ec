#!/usr/bin/env python
"""Summarise `ncu -i rep --page source --csv --kernel-name ...` output: stall-reason totals and the hottest SASS lines.
Usage: python tools/ncu_stalls.py file.csv [top_n]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = {h: 0 for h in stalls}
    total, recs, inst = 0, [], 0
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[0] == "Address":
            continue
        n = int(r[idx["# Samples"]] or 0)
        total += n
        ie = int(r[idx["Instructions Executed"]] or 0)
        inst += ie
        st = {h: int(r[idx[h]] or 0) for h in stalls}
        for h in stalls:
            tot[h] += st[h]
        recs.append((n, r[idx["Source"]].strip(), {h: v for h, v in st.items() if v}, ie))
    print(f"kernel: {rows[0][1] if rows[0] else ''}")
    print(f"total samples {total}, warp instructions executed {inst}")
    for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
        print(f"  {h:28s} {v:7d} {100 * v / max(total, 1):5.1f}%")
    print("hottest instructions:")
    for n, src, st, ie in sorted(recs, key=lambda x: -x[0])[:top_n]:
        top = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{n:7d} {100 * n / max(total, 1):5.1f}% ie={ie:9d}  {src[:64]:64s} {top}")


if __name__ == "__main__":
    main()
