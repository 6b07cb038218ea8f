#!/bin/bash
# TMA poll microbenchmark, hand-off litmus, fenced-mode latency, launch list of the bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/x; mkdir -p $O
timeout 120 tools/_build/tma_poll_microbench 3000 > $O/tma_poll.json 2> $O/tma_poll.err; echo "rc=$?" >> $O/tma_poll.err
timeout 600 python -m pytest tests/test_handoff_litmus.py -q -m gpu > $O/litmus_tests.log 2>&1
cp gpurun_out/handoff_litmus_c*.json $O/ 2>/dev/null
timeout 200 python tools/rt_latency.py 1024 > $O/rt_latency_plain.log 2>&1
ZEN_B200_RT_FENCED=1 timeout 200 python tools/rt_latency.py 1024 > $O/rt_latency_fenced.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hpr_|pcm16|peak_kernel|copy_hop|mask_rows|scale_recip" -c 4000 --csv --log-file $O/launch_list_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-latency --no-e2e-f32 > $O/launch_list_bench.out 2>&1
cat $O/tma_poll.json; tail -2 $O/tma_poll.err; tail -3 $O/litmus_tests.log; cat $O/handoff_litmus_c8.json; grep cluster8 $O/rt_latency_plain.log $O/rt_latency_fenced.log
