#!/bin/bash
# more end-to-end goldens from the race-free build of the reference (soft mask, SSE), then the parity test on them
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out/ref_norace
ZEN_REF_SO=oracle/_ref/libzen_ref_norace.so timeout 600 python oracle/ref/probe_ref_norace.py gpurun_out/ref_norace rt512_sse_soft,rt1024_soft,rt1024_sse > gpurun_out/ref_norace/probe.log 2>&1
cat gpurun_out/ref_norace/probe.log | tail -5
cp gpurun_out/ref_norace/*.npz tests/golden/ 2>/dev/null
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k norace 2>&1 | tail -15
