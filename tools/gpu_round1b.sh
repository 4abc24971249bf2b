#!/bin/bash
# final round-1 evidence: tests, smoke, bench (+reference arm), launch list, full ncu captures
set -u
O=gpurun_out; R=r01b; mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $O/clocks_$R.csv &
SMI=$!
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$R.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu_$R.log
python __graft_entry__.py smoke > $O/smoke_$R.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_$R.log
python bench.py > $O/bench_$R.json 2> $O/bench_$R.err; echo "bench rc=$?"; cut -c1-400 $O/bench_$R.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_$R.json 2>> $O/bench_$R.err; echo "ref rc=$?"
python tools/bench_vlad.py --shape c1 --images 4096 > $O/vlad_c1_$R.json 2>&1; cat $O/vlad_c1_$R.json
python tools/bench_vlad.py --shape c3 --images 16384 > $O/vlad_c3_$R.json 2>&1; cat $O/vlad_c3_$R.json
python tools/bench_sim.py --n 16384 --d 32768 --k 100 > $O/sim_16k_$R.json 2>&1; cat $O/sim_16k_$R.json
python tools/bench_sim.py --n 65536 --nq 16384 --d 32768 --k 100 --check 16 > $O/sim_64k_$R.json 2>&1; cat $O/sim_64k_$R.json
kill $SMI
python bench.py --steps 2 --warmup 3 --no-extra > $O/plain_$R.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_$R.csv \
    python bench.py --steps 2 --warmup 3 --no-extra > $O/ncu_launches_$R.log 2>&1
echo "ncu launches rc=$?"
python bench.py --steps 1 --warmup 3 --images 1184 --no-cpu-baseline --no-extra --e2e-images 64 > $O/plain_fv_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc_kernel|tc2_kernel|fv_finalize" -s 24 -c 4 -f -o $O/prof_fv_$R \
    python bench.py --steps 1 --warmup 3 --images 1184 --no-cpu-baseline --no-extra --e2e-images 64 > $O/ncu_fv_$R.log 2>&1
echo "ncu fv rc=$?"
for S in c3 c1; do
python tools/bench_vlad.py --shape $S --images 4096 --reps 1 > $O/plain_vlad_${S}_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc2_kernel|vlad_aggregate" -s 4 -c 2 -f -o $O/prof_vlad_${S}_$R \
    python tools/bench_vlad.py --shape $S --images 4096 --reps 1 > $O/ncu_vlad_${S}_$R.log 2>&1
echo "ncu vlad $S rc=$?"
done
python tools/bench_sim.py --n 16384 --d 32768 --reps 1 --check 0 > $O/plain_sim_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc2_kernel -s 1 -c 1 -f -o $O/prof_sim_$R \
    python tools/bench_sim.py --n 16384 --d 32768 --reps 1 --check 0 > $O/ncu_sim_$R.log 2>&1
echo "ncu sim rc=$?"
