#!/bin/bash
# A/B on one box: long first tiles when a GPU holds fewer streams than resident CTAs (the 8-GPU shard: 512 streams)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_batch_host.py -q -m gpu -k "fast_tile or batch or tile" 2>&1 | tail -2
run() { python bench.py --streams $2 --scaling weak --steps 5 --warmup 3 --no-e2e-f32 --no-cpu --no-parity --no-latency 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', $2, round(d['value']), d['ms_per_step'])"; }
run two_sizes 512
ZEN_B200_NO_STRETCH=1 run one_size 512
run two_sizes 1024
ZEN_B200_NO_STRETCH=1 run one_size 1024
run two_sizes 512
run full 4096
