#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/y; mkdir -p $O
timeout 900 python -m pytest tests/test_pcm.py tests/test_batch_host.py -q -m gpu > $O/tests.log 2>&1
timeout 200 python tools/pcm_bench.py > $O/pcm_bench.log 2>&1
tail -5 $O/tests.log; cat $O/pcm_bench.log
