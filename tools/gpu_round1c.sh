#!/bin/bash
# round-1c evidence: tests, smoke, bench (+reference arm), launch list, full ncu captures
set -u
O=gpurun_out; R=r01c; mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/clocks_$R.csv &
SMI=$!
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$R.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu_$R.log
python __graft_entry__.py smoke > $O/smoke_$R.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_$R.log
python bench.py > $O/bench_$R.json 2> $O/bench_$R.err; echo "bench rc=$?"; cut -c1-300 $O/bench_$R.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_$R.json 2>> $O/bench_$R.err; echo "ref rc=$?"
kill $SMI
python bench.py --steps 2 --warmup 3 --no-extra > $O/plain_$R.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_$R.csv \
    python bench.py --steps 2 --warmup 3 --no-extra > $O/ncu_launches_$R.log 2>&1
echo "ncu launches rc=$?"
python bench.py --steps 1 --warmup 3 --images 1184 --no-cpu-baseline --no-extra --e2e-images 64 > $O/plain_fv_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc_kernel|tc2_kernel|fv_finalize" -s 36 -c 6 -f -o $O/prof_fv_$R \
    python bench.py --steps 1 --warmup 3 --images 1184 --no-cpu-baseline --no-extra --e2e-images 64 > $O/ncu_fv_$R.log 2>&1
echo "ncu fv rc=$?"
for S in c3 c1; do
python tools/bench_vlad.py --shape $S --images 4096 --reps 1 > $O/plain_vlad_${S}_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc2_kernel|vlad_aggregate" -s 6 -c 3 -f -o $O/prof_vlad_${S}_$R \
    python tools/bench_vlad.py --shape $S --images 4096 --reps 1 > $O/ncu_vlad_${S}_$R.log 2>&1
echo "ncu vlad $S rc=$?"
done
