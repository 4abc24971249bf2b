#!/bin/bash
# fp16x2 VLAD assignment: parity tests + C1 / C3 throughput
O=gpurun_out; R=${R:-r01c}; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q -k "vlad or pipeline or determin" 2>&1 | tail -8
python tools/bench_vlad.py --shape c1 --images 4096 > $O/vlad_c1_$R.json 2>&1; cat $O/vlad_c1_$R.json
python tools/bench_vlad.py --shape c3 --images 16384 > $O/vlad_c3_$R.json 2>&1; cat $O/vlad_c3_$R.json
