#!/usr/bin/env bash
# Round-2 2-GPU call: factored SH exchange after the explicit-rounding weights (diag, tests, bench), wgrad kernel timing.
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python tools/factored_diag.py > $O/c14_diag.log 2>&1
timeout 900 python -m pytest tests/test_gpu_p2p_exchange.py tests/test_gpu_bucket.py -q -p no:cacheprovider > $O/c14_pytest.log 2>&1
timeout 600 $TR --master-port $((29000 + RANDOM % 2000)) bench.py --gpus 2 --steps 30 --warmup 5 > $O/c14_bench_c3_n2.json 2> $O/c14_bench_c3_n2.err
timeout 600 $TR --master-port $((35000 + RANDOM % 2000)) bench.py --gpus 2 --config C4 --steps 15 --warmup 4 > $O/c14_bench_c4_n2.json 2> $O/c14_bench_c4_n2.err
timeout 300 python tools/appearance_profile.py 2>&1 | grep -v "^$" | cut -c1-260 > $O/c14_appearance.log
ls -la $O > $O/c14_ls.txt
echo CALL14_DONE
