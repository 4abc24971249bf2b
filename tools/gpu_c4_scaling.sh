#!/bin/bash
# BASELINE.json configs[3] at full size (262,144 x 32,768-D, top-100) on N GPUs of one box, plus the FV bench at N
N=${N:-4}; O=gpurun_out; mkdir -p $O
if [ "$N" = "1" ]; then
  timeout 900 python tools/bench_retrieval_mgpu.py --rows 262144 --dim 32768 --topk 100 --reps 1 > $O/retrieval_c4_n$N.json 2> $O/retrieval_c4_n$N.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
      tools/bench_retrieval_mgpu.py --rows 262144 --dim 32768 --topk 100 --reps 2 --native-comm > $O/retrieval_c4_n$N.json 2> $O/retrieval_c4_n$N.err
fi
echo "c4 n$N rc=$?"; tail -1 $O/retrieval_c4_n$N.json; tail -2 $O/retrieval_c4_n$N.err | cut -c1-200
if [ "$N" != "1" ] && [ -z "${NOBENCH:-}" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 \
      bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err
  echo "bench n$N rc=$?"; cut -c1-200 $O/bench_n$N.json
fi
