#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 600 -x -k "forward_host or invariances" 2>&1 | tail -5
for mb in 256 512 1024; do
python bench.py --steps 3 --warmup 3 --global-batch 2048 --no-cpu-baseline --micro-batch $mb 2>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('mb $mb value', round(d['value']), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],1), 'launches', d['gpu_launches'])"
done
tail -3 gpurun_out/bench.err
