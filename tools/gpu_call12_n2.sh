#!/usr/bin/env bash
# Round-2 2-GPU call (repeat of call 11 after the single-run checks): factored SH exchange tests + bench, exchange sweep.
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( time timeout 900 python -m pytest tests/test_gpu_p2p_exchange.py tests/test_gpu_bucket.py -q -p no:cacheprovider 2>&1 | tail -40 ) > $O/c12_pytest.log 2>&1
timeout 600 $TR --master-port $((29000 + RANDOM % 2000)) bench.py --gpus 2 --steps 30 --warmup 5 > $O/c12_bench_c3_n2.json 2> $O/c12_bench_c3_n2.err
timeout 600 $TR --master-port $((31000 + RANDOM % 2000)) bench.py --gpus 2 --steps 30 --warmup 5 --no-factor-sh > $O/c12_bench_c3_n2_plain.json 2> $O/c12_bench_c3_n2_plain.err
timeout 600 $TR --master-port $((33000 + RANDOM % 2000)) tools/exchange_sweep.py > $O/c12_sweep.log 2>&1
timeout 600 $TR --master-port $((35000 + RANDOM % 2000)) bench.py --gpus 2 --config C4 --steps 15 --warmup 4 > $O/c12_bench_c4_n2.json 2> $O/c12_bench_c4_n2.err
timeout 300 python tools/appearance_profile.py 2>&1 | grep -v "^$" | cut -c1-260 > $O/c12_appearance.log
ls -la $O > $O/c12_ls.txt
echo CALL12_DONE
