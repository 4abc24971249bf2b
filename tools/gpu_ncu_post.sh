#!/bin/bash
# full ncu capture of the FV kernels of one chunk (posterior / stats / project / finalize)
O=gpurun_out; R=${R:-r01c}
python bench.py --steps 1 --warmup 3 --images 1184 --no-cpu-baseline --no-extra --e2e-images 64 > $O/plain_fv_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc_kernel|tc2_kernel|fv_finalize" -s 36 -c 6 -f -o $O/prof_fv_$R \
    python bench.py --steps 1 --warmup 3 --images 1184 --no-cpu-baseline --no-extra --e2e-images 64 > $O/ncu_fv_$R.log 2>&1
echo "ncu fv rc=$?"; tail -3 $O/ncu_fv_$R.log | cut -c1-300
