#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/j5_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/j5_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e-f32 > gpurun_out/j5_bench.json 2> gpurun_out/j5_bench.err
echo "bench rc=$?" >> gpurun_out/j5_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hpr_tile -s 1 -c 1 -o gpurun_out/r02_fast3 python tools/prof_batch.py 296 30 2 > gpurun_out/j5_ncu.log 2>&1
tail -12 gpurun_out/j5_tests.log; cat gpurun_out/j5_bench.json | cut -c1-300; tail -3 gpurun_out/j5_bench.err
