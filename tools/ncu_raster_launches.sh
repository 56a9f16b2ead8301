#!/bin/bash
# Launch list (device time + occupancy / issue / DRAM per launch) of the rasterizer step at V target views
# per call, no CUDA graph so that every kernel is a separate launch.  usage: tools/ncu_raster_launches.sh V TAG
V=${1:-1}; TAG=${2:-v}
M=gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__grid_size,launch__block_size
mkdir -p gpurun_out
ncu --metrics $M --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}_V${V}.csv \
    python bench.py --views $V --no-graph --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --streams 1 --batched-views 1 \
    > gpurun_out/launches_${TAG}_V${V}.log 2>&1
