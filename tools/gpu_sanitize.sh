#!/bin/bash
# compute-sanitizer memcheck over the small GPU parity tests (one tool per call)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 \
    python -m pytest tests -m gpu -x -q -k "${K:-golden or odd_shapes or ragged or ties or merge or quirk or pair_bf16 or resident or fp16x2}" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|error" gpurun_out/sanitize_memcheck.log | tail -15
