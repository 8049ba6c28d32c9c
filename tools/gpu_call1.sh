#!/usr/bin/env bash
# Round-2 GPU call 1: full GPU test suite (nothing gated), baseline bench of both arms, lane-utilisation statistics,
# compute-sanitizer passes, loss / train-step timing.  Everything lands in gpurun_out/c1_*.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/c1_gpu.txt 2>&1
( time python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -80 ) > $O/c1_pytest.log 2>&1
python bench.py --steps 20 --warmup 5 > $O/c1_bench_ours.json 2> $O/c1_bench_ours.err
python bench.py --impl reference --steps 10 --warmup 3 > $O/c1_bench_ref.json 2> $O/c1_bench_ref.err
python tools/mask_stats.py C3 1 > $O/c1_mask_stats.log 2>&1
GOF_STATS=1 python tools/stats_probe.py C3 > $O/c1_stats_probe.log 2>&1
python tools/timeline.py > $O/c1_timeline.log 2>&1
python tools/loss_bench.py > $O/c1_loss_bench.log 2>&1
python tools/train_step_bench.py C3 > $O/c1_train_step.log 2>&1
for tool in memcheck racecheck initcheck synccheck; do
  ( time timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py 2>&1 | grep -v "^$" | tail -60 ) > $O/c1_sanitizer_$tool.log 2>&1
done
ls -la $O > $O/c1_ls.txt
echo CALL1_DONE
