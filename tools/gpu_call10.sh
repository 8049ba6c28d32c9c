#!/usr/bin/env bash
# Round-2 GPU call 10 (1 GPU): the appearance tail with the library's own weight-gradient kernel (A/B against cuDNN), C4 bench, pool A/B.
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_bucket.py tests/test_gpu_binning.py -q -p no:cacheprovider 2>&1 | tail -15 ) > $O/c10_pytest.log 2>&1
for cfg in "GOF_APP_WGRAD=1" "GOF_APP_WGRAD=0" "GOF_APP_NHWC=1" "GOF_APP_CUDNN_BENCH=1" "GOF_APP_WGRAD=0 GOF_APP_CUDNN_BENCH=1"; do
  echo "=== $cfg" >> $O/c10_appearance.log
  env $cfg timeout 300 python tools/appearance_profile.py 2>&1 | grep -v "^$" | cut -c1-260 >> $O/c10_appearance.log
done
timeout 900 python bench.py --config C4 --steps 20 --warmup 5 --no-cpu-baseline > $O/c10_bench_c4.json 2> $O/c10_bench_c4.err
for cfg in "GOF_POOL=1" "GOF_POOL=0"; do
  env $cfg timeout 300 python tools/step_time.py C3 30 "$cfg" >> $O/c10_ab.log 2>&1
  env $cfg timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/c10_bench_c3_${cfg}.json 2>> $O/c10_ab.log
done
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_bucket.py -q -p no:cacheprovider -k conv3x3 > $O/c10_memcheck_conv.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_bucket.py -q -p no:cacheprovider -k conv3x3 > $O/c10_racecheck_conv.log 2>&1
ls -la $O > $O/c10_ls.txt
echo CALL10_DONE
