#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_batch_host.py tests/test_long_parity.py -q -m gpu -k "fast_tile or batch or tile or long or offline" 2>&1 | tail -2
timeout 300 python tools/batch_sweep.py 2>&1 | cut -c1-200
