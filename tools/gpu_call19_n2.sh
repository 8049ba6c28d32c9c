#!/usr/bin/env bash
# Round-2 2-GPU call: plane-layout records (tests + C3 bench).
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_p2p_exchange.py tests/test_gpu_bucket.py -q -p no:cacheprovider > $O/c19_pytest.log 2>&1
timeout 400 $TR --master-port $((29000 + RANDOM % 2000)) bench.py --gpus 2 --steps 30 --warmup 5 > $O/c19_bench_c3_n2.json 2> $O/c19_bench_c3_n2.err
echo CALL19_DONE
