#!/bin/bash
# One `ncu --set full` capture of the kernels matching REGEX in the rasterizer step at V views per call.
# usage: tools/ncu_full.sh REGEX V TAG [SKIP] [COUNT]   (SKIP / COUNT count launches that MATCH the regex: to capture
#        one launch of each of n kernel types after two warm-up steps use SKIP = 2 n, COUNT = n)
RE=${1:-k_composite}; V=${2:-1}; TAG=${3:-full}; SKIP=${4:-6}; COUNT=${5:-2}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$RE -s $SKIP -c $COUNT -f -o gpurun_out/ncu_${TAG}_V${V} \
    python bench.py --views $V --no-graph --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --streams 1 --batched-views 1 \
    > gpurun_out/ncu_${TAG}_V${V}.log 2>&1
