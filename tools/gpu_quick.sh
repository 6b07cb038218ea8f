#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "fast_tile or batch or tile" > gpurun_out/q_tests.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e-f32 --no-cpu --no-parity --no-latency > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
tail -2 gpurun_out/q_tests.log; python -c "
import json; d=json.load(open('gpurun_out/q_bench.json')); print('VALUE', d['value'], d['ms_per_step'], d['e2e']['value'])"
