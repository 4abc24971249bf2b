#!/bin/bash
timeout 300 python -m pytest tests -m gpu -x -q -k "vlad or pipeline or tensor_path_refuses" 2>&1 | tail -15
timeout 120 python tools/bench_vlad.py --shape c1
timeout 120 python tools/bench_vlad.py --shape c3
timeout 120 python tools/bench_vlad.py --shape c3 --images-per-call 16384
