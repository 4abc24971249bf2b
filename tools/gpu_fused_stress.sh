#!/bin/bash
# repeat the full-size bench with a fused kernel (M=1 single CTA, M=2 cluster) and show failures
export PVS_FV_FUSED=${M:-2}
for i in 1 2 3; do
  timeout 300 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline --e2e-images 64 > gpurun_out/stress_$i.json 2> gpurun_out/stress_$i.err
  echo "run $i rc=$? $(cut -c1-120 gpurun_out/stress_$i.json)"; grep -m2 "Error\|error\|failed" gpurun_out/stress_$i.err | cut -c1-300
done
CUDA_LAUNCH_BLOCKING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline --e2e-images 64 > gpurun_out/stress_b.json 2> gpurun_out/stress_b.err
echo "blocking rc=$? $(cut -c1-120 gpurun_out/stress_b.json)"; grep -m3 "Error\|error\|failed" gpurun_out/stress_b.err | cut -c1-300
