#!/usr/bin/env bash
# Round-2 GPU call 2: one-sweep binning + staged slabs (bulk / cp.async / regs) -- tests, A/B step times, bench, sanitizers with
# full logs, ncu launch list and full captures of the blend kernels.
set -u
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/step_time.jsonl
( time timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 ) > $O/c2_pytest.log 2>&1
for cfg in "GOF_STAGE=bulk" "GOF_STAGE=cpasync" "GOF_STAGE=regs" "GOF_STAGE=bulk GOF_BINNING=legacy" "GOF_STAGE=cpasync GOF_FWD_OCC=3 GOF_BWD_OCC=3" "GOF_STAGE=bulk GOF_FWD_OCC=3 GOF_BWD_OCC=3"; do
  env $cfg timeout 300 python tools/step_time.py C3 30 "$cfg" >> $O/c2_ab.log 2>&1
done
python bench.py --steps 30 --warmup 5 > $O/c2_bench_ours.json 2> $O/c2_bench_ours.err
python tools/timeline.py > $O/c2_timeline.log 2>&1
python tools/mask_stats.py C3 1 > $O/c2_mask_stats.log 2>&1
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 100 python tools/sanitize_run.py > $O/c2_sanitizer_$tool.full.log 2>&1
  grep -E "^=========" $O/c2_sanitizer_$tool.full.log | grep -vE "^=========\s*$" | cut -c1-260 | head -150 > $O/c2_sanitizer_$tool.log
  grep -E "SUMMARY|SANITIZE_RUN_DONE" $O/c2_sanitizer_$tool.full.log >> $O/c2_sanitizer_$tool.log
  gzip -f $O/c2_sanitizer_$tool.full.log
done
# ncu: launch list of one bench-like run, then full captures of the two blend kernels and a one-sweep pass
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/c2_launches.csv python tools/one_iter.py C3 ours 3 > $O/c2_ncu_launches.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_render_forward|k_render_backward" -s 2 -c 2 -o $O/c2_render_full -f python tools/one_iter.py C3 ours 2 > $O/c2_ncu_render.out 2>&1
ncu --set full --clock-control none -k regex:"k_onesweep|k_scan_emit" -s 6 -c 7 -o $O/c2_binning_full -f python tools/one_iter.py C3 ours 2 > $O/c2_ncu_binning.out 2>&1
ls -la $O > $O/c2_ls.txt
echo CALL2_DONE
