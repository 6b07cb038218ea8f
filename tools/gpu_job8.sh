#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 120 python tools/rt_phases.py 1024 > gpurun_out/j8_rt_phases.txt 2>&1
timeout 300 python tools/rt_latency.py 1024 256 > gpurun_out/j8_rt_latency.log 2>&1
timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "resident or realtime" > gpurun_out/j8_tests.log 2>&1
cat gpurun_out/j8_rt_phases.txt; cat gpurun_out/j8_rt_latency.log; tail -3 gpurun_out/j8_tests.log
