#!/bin/bash
# A/B on one box: queue order / stretched tiles of the batched kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() { python bench.py --steps 5 --warmup 3 --no-e2e-f32 --no-cpu --no-parity --no-latency 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), d['ms_per_step'])"; }
run tile_major_stretch
ZEN_B200_NO_STRETCH=1 run tile_major_plain
ZEN_B200_STREAM_MAJOR=1 ZEN_B200_NO_STRETCH=1 run stream_major_plain
ZEN_B200_STREAM_MAJOR=1 run stream_major_stretch
run tile_major_stretch_again
