#!/bin/bash
# configs[3]: the training step at N GPUs of this box, whole-step CUDA graph and eager.  usage: tools/ddp_sweep.sh N
N=${1:-2}
mkdir -p gpurun_out
for mode in "--graph" ""; do
  tag=$([ -z "$mode" ] && echo eager || echo graph)
  if [ "$N" = "1" ]; then
    timeout 150 python tools/train_step.py --batch 7 --steps 10 --warmup 3 $mode > gpurun_out/train_${N}gpu_${tag}.json 2> gpurun_out/train_${N}gpu_${tag}.err
  else
    timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      tools/train_step.py --batch 7 --steps 10 --warmup 3 $mode > gpurun_out/train_${N}gpu_${tag}.json 2> gpurun_out/train_${N}gpu_${tag}.err
  fi
  tail -1 gpurun_out/train_${N}gpu_${tag}.json | cut -c1-700
  grep -i "error\|Traceback" gpurun_out/train_${N}gpu_${tag}.err | head -3
done
