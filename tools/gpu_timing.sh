#!/bin/bash
export PVS_TIMING_PRINT=1
python bench.py --steps 1 --warmup 1 --images 1024 --no-cpu-baseline --no-extra --e2e-images 64 2>&1 | grep "tc2 timing" | sed 's/void pvs::tc2::launch_tc2//' | cut -c1-330 | tail -8
python tools/bench_vlad.py --shape c3 --images 4096 --reps 1 2>&1 | grep "tc2 timing" | cut -c1-330 | tail -2
python tools/bench_vlad.py --shape c1 --images 1024 --reps 1 2>&1 | grep "tc2 timing" | cut -c1-330 | tail -2
