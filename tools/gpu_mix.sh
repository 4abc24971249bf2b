#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 120 python tools/bench_vlad.py --shape c1
timeout 120 python tools/bench_vlad.py --shape c3
timeout 120 python tools/bench_sim.py --n 16384 --d 32768 --k 100
timeout 120 python tools/bench_sim.py --n 65536 --nq 16384 --d 32768 --k 100 --check 16
timeout 120 python tools/bench_sim.py --n 32768 --d 4096 --k 100
