#!/bin/bash
# 8-GPU evidence: BASELINE.json configs[3] at full size (262,144 VLAD-32768 vectors, top-100) and the FV bench
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv | head -9
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
    tools/bench_retrieval_mgpu.py --rows 262144 --dim 32768 --topk 100 --reps 2 --native-comm > $O/retrieval_c4_n8.json 2> $O/retrieval_c4_n8.err
echo "c4 n8 rc=$?"; tail -1 $O/retrieval_c4_n8.json; tail -2 $O/retrieval_c4_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err
echo "bench n8 rc=$?"; cut -c1-260 $O/bench_n8.json
