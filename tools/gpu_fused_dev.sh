#!/bin/bash
# development loop for the fused kernels: parity tests, then phase timing (timing build) and bench with PVS_FV_FUSED=$M
M=${M:-2}
timeout 200 python -m pytest tests -m gpu -x -q -k "fused" 2>&1 | tail -12
export PVS_FV_FUSED=$M
timeout 300 python -m pytest tests -m gpu -x -q -k "fv_golden or fp16x2_path or ragged or independence or device_resident_equals_host_and_oracle" 2>&1 | tail -4
PVS_TIMING_PRINT=1 timeout 200 python bench.py --steps 1 --warmup 1 --images 1184 --no-cpu-baseline --no-extra --e2e-images 64 2>&1 | grep "fused.* timing" | head -2
timeout 300 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline --e2e-images 64 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','stages_ms')})"
