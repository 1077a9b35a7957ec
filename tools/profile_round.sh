#!/bin/bash
# Round profile on the GPU box: bench lines, ncu launch list of the bench command, GEMM DRAM traffic, ncu --set full mix.
# usage: tools/profile_round.sh <tag>      outputs -> gpurun_out/<tag>_*
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
python bench.py --steps 10 --warmup 3 --fwd-only > $out/${tag}_bench_packed.json 2> $out/${tag}_bench_packed.err
python bench.py --steps 6 --warmup 3 --layout dense --fwd-only --no-cpu-baseline > $out/${tag}_bench_dense.json 2> $out/${tag}_bench_dense.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2> /dev/null
# launch list of the bench command (2 timed steps; per-launch times are cold-cache / serialised -> compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1
# DRAM traffic of every GEMM launch of one training step (151 launches after the warm-up step)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --kernel-name-base demangled -k regex:gemm_kernel -s 151 -c 151 --csv --log-file $out/${tag}_gemm_traffic.csv \
    python bench.py --steps 1 --warmup 1 --batches 1 --no-e2e --no-cpu-baseline > $out/${tag}_ncu_traffic.log 2>&1
python tools/gemm_traffic_json.py $out/${tag}_gemm_traffic.csv > $out/${tag}_gemm_traffic_summary.json && cp profiles/r1_gemm_traffic.json $out/r1_gemm_traffic.json
# full-set capture around the forward/backward boundary of the timed step
ncu --set full --clock-control none --kernel-name-base demangled \
    -k "regex:gemm_kernel|attn_diag|attn_fwd_kernel|attn_bwd_kernel|rmsnorm_bwd|add_rmsnorm_fwd_kernel|rmsnorm_fwd_kernel" \
    -s 262 -c 36 -o /tmp/${tag}_full python bench.py --steps 1 --warmup 1 --batches 1 --no-e2e --no-cpu-baseline > $out/${tag}_ncu_full.log 2>&1
tools/ncu_extract.sh /tmp/${tag}_full.ncu-rep $out/${tag}_ncu_full
sz=$(stat -c %s /tmp/${tag}_full.ncu-rep 2>/dev/null || echo 999999999)
if [ "$sz" -lt 30000000 ]; then cp /tmp/${tag}_full.ncu-rep $out/; fi
ls -la $out | tail -20
