#!/bin/bash
# round-2 GPU job 1: new tests, baseline bench with PCM16 e2e, link bandwidth, baseline ncu capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/j1_env.txt; lscpu | head -30 >> gpurun_out/j1_env.txt; free -g >> gpurun_out/j1_env.txt
ls /sys/devices/system/node >> gpurun_out/j1_env.txt
timeout 900 python -m pytest tests/test_batch_host.py tests/test_pcm.py tests/test_long_parity.py -x -q -m gpu > gpurun_out/j1_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/j1_tests.log
timeout 300 python tools/link_bw.py > gpurun_out/j1_link_bw.json 2> gpurun_out/j1_link_bw.err
timeout 200 python tools/pcm_bench.py > gpurun_out/j1_pcm_bench.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/j1_bench.json 2> gpurun_out/j1_bench.err
echo "bench rc=$?" >> gpurun_out/j1_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hpr_tile -s 1 -c 1 -o gpurun_out/r02_base python tools/prof_batch.py 296 30 2 > gpurun_out/j1_ncu.log 2>&1
tail -5 gpurun_out/j1_tests.log; cat gpurun_out/j1_link_bw.json | cut -c1-600; cat gpurun_out/j1_pcm_bench.log | tail -2; cat gpurun_out/j1_bench.json | cut -c1-3000
