#!/usr/bin/env bash
# Round-2 GPU call 4: defaults (bulk forward / register-staged backward, one-sweep v3, pixel-sorted query points), new tests,
# C4 and train_step bench lines of both arms, C5 again.
set -u
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/step_time.jsonl
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 ) > $O/c4_pytest.log 2>&1
timeout 300 python tools/step_time.py C3 30 "defaults" >> $O/c4_ab.log 2>&1
GOF_BINNING=legacy timeout 300 python tools/step_time.py C3 30 "GOF_BINNING=legacy" >> $O/c4_ab.log 2>&1
python tools/timeline.py > $O/c4_timeline.log 2>&1
timeout 900 python bench.py --steps 30 --warmup 5 > $O/c4_bench_c3.json 2> $O/c4_bench_c3.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/c4_bench_c3_ref.json 2> $O/c4_bench_c3_ref.err
timeout 900 python bench.py --config C4 --steps 20 --warmup 5 --no-cpu-baseline > $O/c4_bench_c4.json 2> $O/c4_bench_c4.err
timeout 900 python bench.py --config C4 --impl reference --steps 10 --warmup 3 > $O/c4_bench_c4_ref.json 2> $O/c4_bench_c4_ref.err
timeout 900 python bench.py --mode train_step --steps 20 --warmup 5 > $O/c4_bench_train.json 2> $O/c4_bench_train.err
timeout 900 python bench.py --mode train_step --impl reference --steps 10 --warmup 3 > $O/c4_bench_train_ref.json 2> $O/c4_bench_train_ref.err
timeout 1200 python bench.py --config C5 --no-cpu-baseline > $O/c4_bench_c5.json 2> $O/c4_bench_c5.err
ncu --set full --clock-control none -k regex:"k_onesweep|k_scan_emit|k_integrate" -s 6 -c 7 -o $O/c4_binning_full -f python tools/one_iter.py C3 ours 2 > $O/c4_ncu_binning.out 2>&1
ls -la $O > $O/c4_ls.txt
echo CALL4_DONE
