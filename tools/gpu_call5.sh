#!/usr/bin/env bash
# Round-2 GPU call 5: quarter-warp lists in the blend kernels (A/B against the warp-wide walk, 3 vs 4 CTAs/SM), float64-generated
# scenes, channels-last appearance network.
set -u
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/step_time.jsonl
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -40 ) > $O/c5_pytest.log 2>&1
for cfg in "GOF_SUBWARP=1" "GOF_SUBWARP=0" "GOF_SUBWARP=1 GOF_FWD_OCC=3" "GOF_SUBWARP=1 GOF_FWD_OCC=3 GOF_BWD_OCC=3" "GOF_SUBWARP=1 GOF_STAGE_FWD=cpasync" "GOF_SUBWARP=1 GOF_STAGE_BWD=cpasync"; do
  env $cfg timeout 300 python tools/step_time.py C3 30 "$cfg" >> $O/c5_ab.log 2>&1
done
GOF_STATS=1 python tools/stats_probe.py C3 > $O/c5_stats_probe.log 2>&1
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/c5_bench_c3.json 2> $O/c5_bench_c3.err
timeout 900 python bench.py --config C4 --steps 20 --warmup 5 --no-cpu-baseline > $O/c5_bench_c4.json 2> $O/c5_bench_c4.err
timeout 900 python bench.py --mode train_step --steps 20 --warmup 5 > $O/c5_bench_train.json 2> $O/c5_bench_train.err
timeout 600 compute-sanitizer --tool racecheck --print-limit 100 python tools/sanitize_run.py render > $O/c5_sanitizer_racecheck.full.log 2>&1
grep -E "^=========" $O/c5_sanitizer_racecheck.full.log | grep -vE "^=========\s*$" | cut -c1-260 | head -60 > $O/c5_sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 100 python tools/sanitize_run.py render integrate > $O/c5_sanitizer_memcheck.full.log 2>&1
grep -E "^=========" $O/c5_sanitizer_memcheck.full.log | grep -vE "^=========\s*$" | cut -c1-260 | head -60 > $O/c5_sanitizer_memcheck.log
gzip -f $O/c5_sanitizer_racecheck.full.log $O/c5_sanitizer_memcheck.full.log
ncu --set full --clock-control none --import-source on -k regex:"k_render_forward|k_render_backward" -s 2 -c 2 -o $O/c5_render_full -f python tools/one_iter.py C3 ours 2 > $O/c5_ncu_render.out 2>&1
ls -la $O > $O/c5_ls.txt
echo CALL5_DONE
