#!/usr/bin/env bash
# Round-2 final 1-GPU call: the whole GPU suite, smoke(), the default bench line, compute-sanitizer on the kernels added late
# in the round (factored SH gradient, convolution weight gradient).
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 ) > $O/c21_pytest.log 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $O/c21_smoke.log 2>&1
timeout 300 python bench.py > $O/c21_bench_default.json 2> $O/c21_bench_default.err
timeout 170 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_bucket.py -q -p no:cacheprovider -k "factored or conv3x3 or public" > $O/c21_memcheck.log 2>&1
timeout 170 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_bucket.py -q -p no:cacheprovider -k "conv3x3 or 1027" > $O/c21_racecheck.log 2>&1
echo CALL21_DONE
