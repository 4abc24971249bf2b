#!/bin/bash
# One gpurun call: GPU test suite, smoke, bench lines, then the ncu launch list and one
# `--set full` capture per hot kernel (each only after its plain command exited 0).
# Everything lands in gpurun_out/ (summaries are copied to profiles/ by hand afterwards).
set -u
O=gpurun_out
mkdir -p $O
R=${ROUND_TAG:-r01}
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $O/clocks_$R.csv &
SMI=$!
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$R.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_$R.log
python __graft_entry__.py smoke > $O/smoke_$R.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_$R.log
python bench.py > $O/bench_$R.json 2> $O/bench_$R.err; echo "bench rc=$?"; cat $O/bench_$R.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_$R.json 2>> $O/bench_$R.err; echo "bench ref rc=$?"; cat $O/bench_ref_$R.json
python tools/bench_vlad.py --shape c1 > $O/vlad_c1_$R.json 2>&1; cat $O/vlad_c1_$R.json
python tools/bench_vlad.py --shape c3 > $O/vlad_c3_$R.json 2>&1; cat $O/vlad_c3_$R.json
python tools/bench_sim.py --n 16384 --d 32768 --k 100 > $O/sim_16k_$R.json 2>&1; cat $O/sim_16k_$R.json
python tools/bench_sim.py --n 65536 --nq 16384 --d 32768 --k 100 --check 16 > $O/sim_64k_$R.json 2>&1; cat $O/sim_64k_$R.json
kill $SMI

if [ "${SKIP_NCU:-0}" = "0" ]; then
# launch list of the bench command
python bench.py --steps 2 --warmup 3 > $O/plain_$R.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_$R.csv \
    python bench.py --steps 2 --warmup 3 > $O/ncu_launches_$R.log 2>&1
echo "ncu launches rc=$?"
# full capture: the three FV tensor-core kernels + finalize (one call of 512 images each)
python bench.py --steps 1 --warmup 3 --images 1024 --no-cpu-baseline > $O/plain_fv_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc_kernel|fv_finalize" -s 24 -c 4 -f -o $O/prof_fv_$R \
    python bench.py --steps 1 --warmup 3 --images 1024 --no-cpu-baseline > $O/ncu_fv_$R.log 2>&1
echo "ncu fv rc=$?"
python tools/bench_vlad.py --shape c3 --images 4096 --reps 1 > $O/plain_vlad_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_nt|argmin|vlad_aggregate|tc_kernel" -s 6 -c 3 -f -o $O/prof_vlad_c3_$R \
    python tools/bench_vlad.py --shape c3 --images 4096 --reps 1 > $O/ncu_vlad_$R.log 2>&1
echo "ncu vlad rc=$?"
python tools/bench_sim.py --n 8192 --d 8192 --reps 1 --check 0 > $O/plain_sim_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_kernel -s 1 -c 1 -f -o $O/prof_sim_$R \
    python tools/bench_sim.py --n 8192 --d 8192 --reps 1 --check 0 > $O/ncu_sim_$R.log 2>&1
echo "ncu sim rc=$?"
fi
ls -la $O
