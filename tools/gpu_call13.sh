#!/usr/bin/env bash
# Round-2 GPU call 13 (1 GPU): diagnostics of the factored SH gradient, the bucket tests with full tracebacks, wgrad kernel under ncu.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/factored_diag.py > $O/c13_diag.log 2>&1
timeout 600 python -m pytest tests/test_gpu_bucket.py -q -p no:cacheprovider > $O/c13_pytest.log 2>&1
timeout 300 python tools/appearance_profile.py 2>&1 | grep -v "^$" | cut -c1-260 > $O/c13_appearance.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_conv3x3_wgrad" -c 3 -o $O/c13_wgrad -f python tools/appearance_profile.py > $O/c13_ncu.out 2>&1
ls -la $O > $O/c13_ls.txt
echo CALL13_DONE
