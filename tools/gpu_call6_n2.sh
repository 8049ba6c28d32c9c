#!/usr/bin/env bash
# Round-2 multi-GPU call (gpurun --gpus 2): the exchange kernels against NCCL, the training bench with every exchange mode,
# the extraction bench view- / tet-sharded over two ranks.
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi topo -m > $O/c6_topo.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_p2p_exchange.py -q -p no:cacheprovider 2>&1 | tail -30 ) > $O/c6_pytest_exchange.log 2>&1
for ex in auto nccl p2p nvls; do
  timeout 600 $TR --master-port $((29000 + RANDOM % 2000)) bench.py --gpus 2 --steps 30 --warmup 5 --exchange $ex > $O/c6_bench_c3_n2_$ex.json 2> $O/c6_bench_c3_n2_$ex.err
done
timeout 600 $TR --master-port $((31000 + RANDOM % 2000)) bench.py --gpus 2 --impl reference --steps 10 --warmup 3 > $O/c6_bench_c3_n2_ref.json 2> $O/c6_bench_c3_n2_ref.err
timeout 600 $TR --master-port $((33000 + RANDOM % 2000)) bench.py --gpus 2 --config C4 --steps 15 --warmup 4 > $O/c6_bench_c4_n2.json 2> $O/c6_bench_c4_n2.err
timeout 1200 $TR --master-port $((35000 + RANDOM % 2000)) bench.py --gpus 2 --config C5 > $O/c6_bench_c5_n2.json 2> $O/c6_bench_c5_n2.err
timeout 600 $TR --master-port $((37000 + RANDOM % 2000)) bench.py --gpus 2 --mode train_step --steps 15 --warmup 4 > $O/c6_bench_train_n2.json 2> $O/c6_bench_train_n2.err
ls -la $O > $O/c6_ls.txt
echo CALL6_DONE
