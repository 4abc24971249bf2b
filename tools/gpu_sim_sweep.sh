#!/bin/bash
for s in 3 5 9 12 18 32; do
  echo "stripes=$s"; PVS_SIM_STRIPES=$s timeout 120 python tools/bench_sim.py --n 16384 --d 32768 --k 100 --check 0 | cut -c1-150
  PVS_SIM_STRIPES=$s timeout 120 python tools/bench_sim.py --n 65536 --nq 16384 --d 32768 --k 100 --check 0 | cut -c1-150
done
