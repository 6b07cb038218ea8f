#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/j12_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/j12_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e-f32 --no-cpu > gpurun_out/j12_bench.json 2> gpurun_out/j12_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hpr_tile -s 1 -c 1 -o gpurun_out/r02_fast4 python tools/prof_batch.py 296 30 2 > gpurun_out/j12_ncu.log 2>&1
tail -5 gpurun_out/j12_tests.log; python -c "
import json; d=json.load(open('gpurun_out/j12_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['latency']['resident_kernel'], d['parity']['ok'], d['parity']['batched_equals_checked_path'])"
