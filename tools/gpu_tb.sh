#!/bin/bash
# whole GPU suite + the bench line (no ncu): the quick confirmation of a build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/tb; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu > $O/gpu_tests.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 300 python tools/batch_sweep.py > $O/batch_sweep.log 2>&1; cp gpurun_out/batch_sweep.json $O/ 2>/dev/null
tail -3 $O/gpu_tests.log; cut -c1-300 $O/bench.json; tail -1 $O/bench.err
