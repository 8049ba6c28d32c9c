#!/usr/bin/env bash
# Round-2 8-GPU call: plane-layout records (C3 at 8 and 4 GPUs).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29000 + RANDOM % 2000)) bench.py --gpus 8 --steps 30 --warmup 5 > $O/c20_bench_c3_n8.json 2> $O/c20_bench_c3_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((35000 + RANDOM % 2000)) bench.py --gpus 4 --steps 30 --warmup 5 > $O/c20_bench_c3_n4.json 2> $O/c20_bench_c3_n4.err
echo CALL20_DONE
