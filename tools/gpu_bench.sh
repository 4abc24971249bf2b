#!/bin/bash
timeout 600 python bench.py > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; echo "rc=$?"; tail -3 gpurun_out/bench_now.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_now.json'))
for k in ('value','ms_per_step','e2e','gpu_launches','roofline','stages_ms','cpu_baseline','extra','clocks'):
    print(k, '=', d.get(k))
PY
