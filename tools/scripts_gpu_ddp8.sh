#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613"
timeout 170 $TR tools/bench_train.py --steps 20 --warmup 5 > gpurun_out/train_bench_n$N.json 2> gpurun_out/train_bench_n$N.err
echo "bench n$N rc=$?"; tail -2 gpurun_out/train_bench_n$N.err; cut -c1-420 gpurun_out/train_bench_n$N.json
