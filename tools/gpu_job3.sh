#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python tools/diag_fast.py > gpurun_out/j3_diag.log 2>&1
DIAG_FLAGS=7 DIAG_HOPS=60 python tools/diag_fast.py >> gpurun_out/j3_diag.log 2>&1
DIAG_HOPS=40 timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/diag_fast.py > gpurun_out/j3_race.log 2>&1
timeout 900 python -m pytest tests/test_reference_tests.py -q -m gpu > gpurun_out/j3_reftests.log 2>&1
cat gpurun_out/j3_diag.log; tail -30 gpurun_out/j3_race.log; tail -15 gpurun_out/j3_reftests.log
