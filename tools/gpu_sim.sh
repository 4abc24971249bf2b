#!/bin/bash
timeout 300 python -m pytest tests -m gpu -x -q -k "topk or all_pairs or eval" 2>&1 | tail -15
timeout 120 python tools/bench_sim.py --n 16384 --d 32768 --k 100
timeout 120 python tools/bench_sim.py --n 65536 --nq 16384 --d 32768 --k 100 --check 16
timeout 120 python tools/bench_sim.py --n 32768 --d 4096 --k 100
timeout 120 python tools/bench_sim.py --n 8192 --d 8192 --k 100 --reps 2
if [ "${NCU:-0}" = "1" ]; then
python tools/bench_sim.py --n 16384 --d 32768 --reps 1 --check 0 > gpurun_out/plain_sim2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc2_kernel -s 1 -c 1 -f -o gpurun_out/prof_sim2_r01 \
    python tools/bench_sim.py --n 16384 --d 32768 --reps 1 --check 0 > gpurun_out/ncu_sim2.log 2>&1
echo "ncu rc=$?"
fi
