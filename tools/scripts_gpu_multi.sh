#!/bin/bash
# N-GPU bench through torchrun exactly as the driver launches it
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${N:-2}
nvidia-smi -L | wc -l
MRD_BENCH_WATCHDOG=240 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; grep -v "^\*\|OMP_NUM\|Warn" gpurun_out/bench_n$N.err | tail -5
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'launches', d['gpu_launches'], d['clocks'])
PY
