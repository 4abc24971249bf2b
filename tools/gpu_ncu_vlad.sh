#!/bin/bash
S=${SHAPE:-c3}; N=${N:-4096}
python tools/bench_vlad.py --shape $S --images $N --reps 1 > gpurun_out/plain_vlad_$S.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc2_kernel|vlad_aggregate" -s 4 -c 2 -f -o gpurun_out/prof_vlad2_$S \
    python tools/bench_vlad.py --shape $S --images $N --reps 1 > gpurun_out/ncu_vlad_$S.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_vlad_$S.log
