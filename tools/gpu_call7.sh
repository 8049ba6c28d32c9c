#!/usr/bin/env bash
# Round-2 GPU call 7 (1 GPU): full test suite on the current defaults, A/B of the sort chunk size, bench lines, ncu evidence.
set -u
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/step_time.jsonl
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 ) > $O/c7_pytest.log 2>&1
for cfg in "GOF_SORT_KEYS=8" "GOF_SORT_KEYS=16" "GOF_SUBWARP_BWD=0"; do
  env $cfg timeout 300 python tools/step_time.py C3 30 "$cfg" >> $O/c7_ab.log 2>&1
done
python tools/timeline.py > $O/c7_timeline.log 2>&1
timeout 900 python bench.py --steps 30 --warmup 5 > $O/c7_bench_c3.json 2> $O/c7_bench_c3.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/c7_bench_c3_ref.json 2> $O/c7_bench_c3_ref.err
timeout 900 python bench.py --config C4 --steps 20 --warmup 5 --no-cpu-baseline > $O/c7_bench_c4.json 2> $O/c7_bench_c4.err
timeout 900 python bench.py --mode train_step --steps 20 --warmup 5 > $O/c7_bench_train.json 2> $O/c7_bench_train.err
timeout 900 python bench.py --config C2 --steps 30 --warmup 5 --no-cpu-baseline > $O/c7_bench_c2.json 2> $O/c7_bench_c2.err
timeout 900 python bench.py --config C2 --impl reference --steps 20 --warmup 5 > $O/c7_bench_c2_ref.json 2> $O/c7_bench_c2_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/c7_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/c7_ncu_launches.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_render_forward|k_render_backward|k_preprocess" -s 4 -c 4 -o $O/c7_render_full -f python tools/one_iter.py C3 ours 2 > $O/c7_ncu_render.out 2>&1
ls -la $O > $O/c7_ls.txt
echo CALL7_DONE
