#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-images 1024 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','stages_ms')}); print(d['extra'])"
