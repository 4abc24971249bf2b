#!/bin/bash
# fp16x2 FV path: parity tests, accuracy probe, stage timings
set -u
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q ${K:+-k "$K"} 2>&1 | tail -15
timeout 120 python tools/probe_fv_err.py 2>&1 | tail -8
timeout 300 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline --e2e-images 64 > $O/bench_fv16.json 2> $O/bench_fv16.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_fv16.json"))
print({k:d[k] for k in ("value","ms_per_step","stages_ms","gpu_launches")}); print(d["roofline"])
PY
