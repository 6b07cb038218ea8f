#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_pcm.py tests/test_batch_host.py -q -m gpu -k "box or median or tall or pcm or filters" > gpurun_out/j14_tests.log 2>&1
timeout 200 python tools/box_bench.py > gpurun_out/j14_box_bench.json 2>&1
timeout 200 python tools/pcm_bench.py > gpurun_out/j14_pcm_bench.log 2>&1
timeout 300 python tools/mfilt_bench.py > gpurun_out/j14_mfilt_bench.json 2>&1
timeout 200 python tools/fft_bench.py > gpurun_out/j14_fft_bench.json 2>&1
tail -15 gpurun_out/j14_tests.log; cat gpurun_out/j14_box_bench.json; tail -1 gpurun_out/j14_pcm_bench.log; python -c "
import json; d=json.load(open('gpurun_out/j14_mfilt_bench.json')); print({k:v for k,v in d.items() if 'N16384' in k or 'N4096' in k}); print(json.load(open('gpurun_out/j14_fft_bench.json')))"
