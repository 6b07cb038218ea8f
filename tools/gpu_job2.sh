#!/bin/bash
# round-2 GPU job 2: the packed FFT + fast tile kernel: whole GPU suite, bench, ncu of the new kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/j2_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/j2_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e-f32 > gpurun_out/j2_bench.json 2> gpurun_out/j2_bench.err
echo "bench rc=$?" >> gpurun_out/j2_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hpr_tile -s 1 -c 1 -o gpurun_out/r02_fast python tools/prof_batch.py 296 30 2 > gpurun_out/j2_ncu.log 2>&1
tail -25 gpurun_out/j2_tests.log; cat gpurun_out/j2_bench.json | cut -c1-600; tail -3 gpurun_out/j2_bench.err
