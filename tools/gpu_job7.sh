#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 120 python tools/rt_phases.py 1024 > gpurun_out/j7_rt_phases.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "fast_tile or batch_equals or resident" > gpurun_out/j7_tests.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e-f32 --no-parity --no-cpu > gpurun_out/j7_bench.json 2> gpurun_out/j7_bench.err
cat gpurun_out/j7_rt_phases.txt; tail -3 gpurun_out/j7_tests.log; python -c "
import json; d=json.load(open('gpurun_out/j7_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['latency']['resident_kernel'])"
