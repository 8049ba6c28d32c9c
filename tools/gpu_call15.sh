#!/usr/bin/env bash
# Round-2 GPU call 15 (1 GPU): full GPU suite after the K8 / exchange changes, pipelined wgrad kernel timing, C3 + C4 bench lines.
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 ) > $O/c15_pytest.log 2>&1
timeout 300 python tools/appearance_profile.py 2>&1 | grep -v "^$" | cut -c1-260 > $O/c15_appearance.log
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/c15_bench_c3.json 2> $O/c15_bench_c3.err
timeout 900 python bench.py --config C4 --steps 20 --warmup 5 --no-cpu-baseline > $O/c15_bench_c4.json 2> $O/c15_bench_c4.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_conv3x3_wgrad" -c 3 -o $O/c15_wgrad -f python tools/appearance_profile.py > $O/c15_ncu.out 2>&1
ls -la $O > $O/c15_ls.txt
echo CALL15_DONE
