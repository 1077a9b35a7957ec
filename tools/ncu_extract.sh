#!/bin/bash
# usage: tools/ncu_extract.sh <report.ncu-rep> <out-prefix>   — writes compact CSV summaries next to the report
rep=$1; out=$2
ncu -i "$rep" --page raw --csv > "${out}_raw.csv" 2>/dev/null
python - "$out" <<'PY'
import csv, sys
out = sys.argv[1]
keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_sleeping_per_warp_active.pct", "sm__pipe_tensor_subunit_cycles_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(open(out + "_raw.csv")))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
sel = [h for h in hdr if h in keep or "tensor" in h or h.startswith("dram__bytes") or "pipe_xu" in h]
with open(out + "_summary.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(sel)
    for r in rows[2:]:
        w.writerow([r[idx[h]] for h in sel])
print("kernels:", len(rows) - 2, "metrics kept:", len(sel))
PY
gzip -f "${out}_raw.csv"
