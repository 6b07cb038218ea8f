#!/bin/bash
# the two ncu --set full captures and the launch list only (the rest of tools/gpu_final.sh already ran on this build)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/final; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hpr_|pcm16|peak_kernel|copy_hop|mask_rows|scale_recip" -c 4000 --csv --log-file $O/launch_list_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-latency --no-e2e-f32 > $O/launch_list_bench.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hpr_tile -s 1 -c 1 -o $O/tile_kernel_296x30 python tools/prof_batch.py 296 30 2 > $O/ncu_296.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:hpr_tile -s 1 -c 1 -o $O/tile_kernel_full python tools/prof_batch.py 4096 60 2 > $O/ncu_full.log 2>&1
ls -la $O | tail -5
