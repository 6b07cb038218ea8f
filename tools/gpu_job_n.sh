#!/bin/bash
# usage: gpu_job_n.sh N   (under gpurun --gpus N): bench + raw link bandwidth at N ranks
N=$1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/link_bw.py > gpurun_out/link_bw_n$N.json 2> gpurun_out/link_bw_n$N.err
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e-f32 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench rc=$?" >> gpurun_out/bench_n$N.err
cat gpurun_out/link_bw_n$N.json | cut -c1-400; tail -3 gpurun_out/bench_n$N.err; python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1])
    print({k:d.get(k) for k in ('value','ms_per_step','scaling','n_gpus')}, d['e2e'], d.get('weak'), d.get('parity',{}).get('ok'))
except Exception as e: print('parse', e)
PY
