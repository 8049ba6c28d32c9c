#!/usr/bin/env bash
# Round-2 GPU call 3: extraction tests + C5 bench (small, then full, then the reference arm), staging variants with the
# right-sized shared-memory carveout, one-sweep passes with vector / warp-wide look-back.
set -u
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/step_time.jsonl
( time timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 ) > $O/c3_pytest.log 2>&1
for cfg in "GOF_STAGE_FWD=bulk GOF_STAGE_BWD=regs" "GOF_STAGE_FWD=cpasync GOF_STAGE_BWD=cpasync" "GOF_STAGE_FWD=bulk GOF_STAGE_BWD=bulk" "GOF_STAGE_FWD=regs GOF_STAGE_BWD=regs" "GOF_STAGE_FWD=bulk GOF_STAGE_BWD=regs GOF_BINNING=legacy"; do
  env $cfg timeout 300 python tools/step_time.py C3 30 "$cfg" >> $O/c3_ab.log 2>&1
done
python tools/timeline.py > $O/c3_timeline.log 2>&1
timeout 600 python bench.py --config C5 --points 2000000 --views 8 --tets 4000000 --steps 4 --warmup 3 > $O/c3_bench_c5_small.json 2> $O/c3_bench_c5_small.err
timeout 1200 python bench.py --config C5 > $O/c3_bench_c5.json 2> $O/c3_bench_c5.err
timeout 900 python bench.py --config C5 --impl reference --steps 6 --warmup 3 > $O/c3_bench_c5_ref.json 2> $O/c3_bench_c5_ref.err
timeout 600 compute-sanitizer --tool racecheck --print-limit 100 python tools/sanitize_run.py > $O/c3_sanitizer_racecheck.full.log 2>&1
grep -E "^=========" $O/c3_sanitizer_racecheck.full.log | grep -vE "^=========\s*$" | cut -c1-260 | head -100 > $O/c3_sanitizer_racecheck.log
gzip -f $O/c3_sanitizer_racecheck.full.log
ncu --set full --clock-control none -k regex:"k_onesweep|k_scan_emit" -s 6 -c 7 -o $O/c3_binning_full -f python tools/one_iter.py C3 ours 2 > $O/c3_ncu_binning.out 2>&1
nvidia-smi --query-gpu=memory.used,memory.total --format=csv > $O/c3_mem.txt
ls -la $O > $O/c3_ls.txt
echo CALL3_DONE
