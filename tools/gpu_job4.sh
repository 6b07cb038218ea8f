#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
ZEN_REF_SO=oracle/_ref/libzen_ref_norace.so timeout 600 python oracle/ref/probe_ref_norace.py gpurun_out/ref_norace > gpurun_out/j4_norace.log 2>&1
cp gpurun_out/ref_norace/*.npz tests/golden/ 2>/dev/null
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/j4_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/j4_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e-f32 > gpurun_out/j4_bench.json 2> gpurun_out/j4_bench.err
echo "bench rc=$?" >> gpurun_out/j4_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hpr_tile -s 1 -c 1 -o gpurun_out/r02_fast2 python tools/prof_batch.py 296 30 2 > gpurun_out/j4_ncu.log 2>&1
cat gpurun_out/j4_norace.log; tail -40 gpurun_out/j4_tests.log; cat gpurun_out/j4_bench.json | cut -c1-300; tail -3 gpurun_out/j4_bench.err
