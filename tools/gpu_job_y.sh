#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/y; mkdir -p $O
timeout 900 python -m pytest tests/test_onset.py tests/test_pitch.py -q -m gpu > $O/tests.log 2>&1
tail -15 $O/tests.log
