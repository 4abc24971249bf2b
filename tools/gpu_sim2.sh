#!/bin/bash
timeout 300 python -m pytest tests -m gpu -x -q -k "topk or all_pairs or similarity_topk" 2>&1 | tail -3
for e in "" "PVS_SIM_NOSYNC=1"; do
echo "== $e"
env $e python tools/bench_sim.py --n 16384 --d 32768 --k 100 --check 0 | cut -c1-140
env $e python tools/bench_sim.py --n 65536 --nq 16384 --d 32768 --k 100 --check 0 | cut -c1-140
env $e python tools/bench_sim.py --n 262144 --nq 32768 --d 32768 --k 100 --reps 1 --check 0 | cut -c1-140
done
