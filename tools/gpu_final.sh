#!/bin/bash
# final check of the tree: all GPU tests, smoke, default bench, and the two fused modes on the same box
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -1
for M in 0 2 1; do
  if [ $M = 0 ]; then unset PVS_FV_FUSED; else export PVS_FV_FUSED=$M; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > $O/bench_fused_mode$M.json 2> $O/bench_fused_mode$M.err
  echo "mode $M rc=$? $(python -c "import json; d=json.load(open('$O/bench_fused_mode$M.json')); print(round(d['value']), round(d['ms_per_step'],3), d['stages_ms'])")"
done
