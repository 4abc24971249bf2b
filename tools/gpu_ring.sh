#!/bin/bash
# ring-pass retrieval (database sharded) at N GPUs vs the replicated path
N=${N:-2}; O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 \
    tools/bench_retrieval_mgpu.py --rows ${ROWS:-65536} --dim 32768 --topk 100 --reps 2 --ring > $O/retrieval_ring_n$N.json 2> $O/retrieval_ring_n$N.err
echo "ring n$N rc=$?"; tail -1 $O/retrieval_ring_n$N.json; grep -i "error" $O/retrieval_ring_n$N.err | head -5 | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 \
    tools/bench_retrieval_mgpu.py --rows ${ROWS:-65536} --dim 32768 --topk 100 --reps 2 > $O/retrieval_repl_n$N.json 2> $O/retrieval_repl_n$N.err
echo "replicated n$N rc=$?"; tail -1 $O/retrieval_repl_n$N.json
