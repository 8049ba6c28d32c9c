#!/usr/bin/env bash
# Round-2 8-GPU call: factored SH exchange with the prefetching expansion kernel (C3 at 8 GPUs with and without the overlapped
# expansion, C3 at 4 GPUs, C4 at 8).
set -u
mkdir -p gpurun_out
O=gpurun_out
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 400 $TR8 --master-port $((29000 + RANDOM % 2000)) bench.py --gpus 8 --steps 30 --warmup 5 > $O/c18_bench_c3_n8.json 2> $O/c18_bench_c3_n8.err
GOF_DP_OVERLAP=1 timeout 400 $TR8 --master-port $((27000 + RANDOM % 2000)) bench.py --gpus 8 --steps 30 --warmup 5 > $O/c18_bench_c3_n8_overlap.json 2> $O/c18_bench_c3_n8_overlap.err
timeout 400 $TR4 --master-port $((35000 + RANDOM % 2000)) bench.py --gpus 4 --steps 30 --warmup 5 > $O/c18_bench_c3_n4.json 2> $O/c18_bench_c3_n4.err
timeout 400 $TR8 --master-port $((31000 + RANDOM % 2000)) bench.py --gpus 8 --config C4 --steps 15 --warmup 4 > $O/c18_bench_c4_n8.json 2> $O/c18_bench_c4_n8.err
ls -la $O > $O/c18_ls.txt
echo CALL18_DONE
