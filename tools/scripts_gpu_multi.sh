#!/bin/bash
# N-GPU bench through torchrun exactly as the driver launches it
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${N:-2}
nvidia-smi -L | head -8
export NCCL_DEBUG=WARN
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/nccl_probe.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -20
echo "probe rc=$?"
MRD_BENCH_WATCHDOG=240 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --global-batch ${GB:-1024} > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/bench_n$N.err | tail -30; cat gpurun_out/bench_n$N.json | cut -c1-600
