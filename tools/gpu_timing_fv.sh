#!/bin/bash
# role / phase timing of the FV kernels (library must be built with PVS_NVCC_EXTRA=-DPVS_TIMING)
export PVS_TIMING_PRINT=1
python bench.py --steps 1 --warmup 1 --images 1184 --no-cpu-baseline --no-extra --e2e-images 64 2>&1 | grep -A1 "timing" | sed 's/void pvs::tc2::launch_tc2//; s/void pvs::tc::launch_tc//' | grep -v "^--" | cut -c1-400 | head -${LINES_MAX:-12}
