#!/bin/bash
# round-2 closing capture (one GPU): whole GPU suite, the bench line, ncu launch list and --set full captures, latency
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/final; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu > $O/gpu_tests.log 2>&1
echo "tests rc=$?" >> $O/gpu_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err
echo "bench rc=$?" >> $O/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hpr_|pcm16|peak_kernel|copy_hop|mask_rows|scale_recip" -c 4000 --csv --log-file $O/launch_list_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-latency --no-e2e-f32 > $O/launch_list_bench.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hpr_tile -s 1 -c 1 -o $O/tile_kernel_296x30 python tools/prof_batch.py 296 30 2 > $O/ncu_296.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:hpr_tile -s 1 -c 1 -o $O/tile_kernel_full python tools/prof_batch.py 4096 60 2 > $O/ncu_full.log 2>&1
timeout 300 python tools/rt_latency.py 1024 512 256 > $O/rt_latency.log 2>&1; cp gpurun_out/rt_latency.json $O/ 2>/dev/null
timeout 120 python tools/rt_phases.py 1024 > $O/rt_phases_hop1024.txt 2>&1
timeout 400 python tools/batch_sweep.py > $O/batch_sweep.log 2>&1; cp gpurun_out/batch_sweep.json $O/ 2>/dev/null
timeout 300 python tools/offline_bench.py > $O/other_configs.json 2>&1
timeout 200 python tools/mfilt_bench.py > $O/mfilt_bench.json 2>&1
timeout 200 python tools/box_bench.py > $O/box_bench.json 2>&1
timeout 200 python tools/fft_bench.py > $O/fft_bench.json 2>&1
timeout 200 python tools/pcm_bench.py > $O/pcm_bench.log 2>&1
timeout 200 python tools/link_bw.py > $O/link_bw_n1.json 2>&1
timeout 400 python tools/hps_bench.py > $O/hps_bench_sweep.log 2>&1; cp gpurun_out/hps_bench.json $O/hps_bench_sweep.json 2>/dev/null
timeout 120 tools/_build/stream_bench > $O/stream_bench.json 2>&1
timeout 120 tools/_build/tma_poll_microbench 3000 > $O/tma_poll_microbench.json 2>&1
cp gpurun_out/handoff_litmus_c8.json gpurun_out/pcm_bench.json $O/ 2>/dev/null
cp gpurun_out/long_parity.json $O/ 2>/dev/null
rm -f $O/tile_kernel_296x30.ncu-rep.keep
tail -4 $O/gpu_tests.log; cut -c1-400 $O/bench.json; tail -2 $O/bench.err; tail -3 $O/rt_latency.log; ls -la $O
