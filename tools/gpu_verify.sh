#!/bin/bash
# re-entry check: GPU tests, smoke, default bench
set -u
O=gpurun_out; R=${R:-r01c}; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$R.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_$R.log
timeout 200 python __graft_entry__.py smoke > $O/smoke_$R.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_$R.log
timeout 400 python bench.py > $O/bench_$R.json 2> $O/bench_$R.err; echo "bench rc=$?"; cut -c1-600 $O/bench_$R.json
