#!/bin/bash
# Same-box A/B of a compile-time variant of the library: tools/gemm_ab.sh "<nvcc defines of variant B>" [gemm_bench args]
# builds variant B into /tmp/libggpt_ab.so on the GPU box and runs tools/gemm_bench.py alternately with both builds.
defs=$1; shift
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC $defs -I include -shared -o /tmp/libggpt_ab.so graph-gpt_b200/csrc/*.cu -lcudart || exit 1
for rep in 1 2; do
  echo "== A (tree build)"; python tools/gemm_bench.py ab_a pair 2>&1 | grep -v "^case\|^$"
  echo "== B ($defs)"; GGPT_LIB_PATH=/tmp/libggpt_ab.so python tools/gemm_bench.py ab_b pair 2>&1 | grep -v "^case\|^$"
done
