#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/j6_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/j6_tests.log
timeout 300 python tools/rt_latency.py 1024 > gpurun_out/j6_rt_latency.log 2>&1
timeout 120 python tools/rt_phases.py 1024 > gpurun_out/j6_rt_phases.txt 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e-f32 --no-parity > gpurun_out/j6_bench.json 2> gpurun_out/j6_bench.err
tail -12 gpurun_out/j6_tests.log; cat gpurun_out/j6_rt_latency.log; cat gpurun_out/j6_rt_phases.txt; python -c "
import json; d=json.load(open('gpurun_out/j6_bench.json')); print(d['value'], d['e2e']['value'], d['latency']['resident_kernel'], d['latency']['resident_two_call'])"
