#!/bin/bash
# A/B of the compositor's list-run count: bench value + stage times at V = 1 and the batched V = 4 leg.
# usage: tools/ab_segments.sh "1 4"
mkdir -p gpurun_out
for K in ${1:-1 2 4}; do
  PIXELSPLAT_B200_SEGMENTS=$K timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/ab_seg$K.json 2> gpurun_out/ab_seg$K.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_seg$K.json"))
print("K=$K value", round(d["value"],1), "ms", round(d["ms_per_step"],4), {k: round(v,4) for k,v in d["stage_ms"].items()}, "V4", round(d["throughput_batched_views"]["value"],1), "streams", round(d["throughput_concurrent_streams"]["value"],1))
PY
done
