#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_pcm.py tests/test_batch_host.py -q -m gpu -k "fft or pcm" > gpurun_out/j15_tests.log 2>&1
timeout 200 python tools/pcm_bench.py > gpurun_out/j15_pcm_bench.log 2>&1
timeout 200 python tools/fft_bench.py > gpurun_out/j15_fft_bench.json 2>&1
tail -4 gpurun_out/j15_tests.log; tail -1 gpurun_out/j15_pcm_bench.log; python -c "
import json; print(json.load(open('gpurun_out/j15_fft_bench.json')))"
