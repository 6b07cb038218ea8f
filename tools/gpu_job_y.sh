#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/y; mkdir -p $O
timeout 900 python -m pytest tests/test_onset.py tests/test_gpu_parity.py -q -m gpu -k "onset or median or filter or mfilt or npp or tall or box or sse" > $O/tests.log 2>&1
timeout 300 python -m pytest tests/test_reference_tests.py -q -m gpu >> $O/tests.log 2>&1
timeout 200 python tools/box_bench.py > $O/box_bench.json 2>&1
timeout 200 python tools/mfilt_bench.py > $O/mfilt_bench.json 2>&1
grep -E "passed|failed" $O/tests.log; cat $O/box_bench.json | tr -d '\n ' ; echo; grep -A3 "N16384" $O/mfilt_bench.json | tr -d '\n '
