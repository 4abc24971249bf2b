#!/bin/bash
python bench.py --steps 1 --warmup 3 --images 1024 --no-cpu-baseline > gpurun_out/plain_fv2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc_kernel|tc2_kernel|fv_finalize" -s 24 -c 4 -f -o gpurun_out/prof_fv2_r01 \
    python bench.py --steps 1 --warmup 3 --images 1024 --no-cpu-baseline > gpurun_out/ncu_fv2.log 2>&1
echo "ncu rc=$?"
