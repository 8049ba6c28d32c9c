#!/usr/bin/env bash
# Round-2 8-GPU call: training bench with the autotuned exchange, C4, C5 extraction (view- and tet-sharded).
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $TR --master-port $((29000 + RANDOM % 2000)) bench.py --gpus 8 --steps 30 --warmup 5 > $O/c8_bench_c3_n8.json 2> $O/c8_bench_c3_n8.err
timeout 600 $TR --master-port $((31000 + RANDOM % 2000)) bench.py --gpus 8 --config C4 --steps 15 --warmup 4 > $O/c8_bench_c4_n8.json 2> $O/c8_bench_c4_n8.err
timeout 900 $TR --master-port $((33000 + RANDOM % 2000)) bench.py --gpus 8 --config C5 > $O/c8_bench_c5_n8.json 2> $O/c8_bench_c5_n8.err
timeout 600 $TR4 --master-port $((35000 + RANDOM % 2000)) bench.py --gpus 4 --steps 30 --warmup 5 > $O/c8_bench_c3_n4.json 2> $O/c8_bench_c3_n4.err
ls -la $O > $O/c8_ls.txt
echo CALL8_DONE
