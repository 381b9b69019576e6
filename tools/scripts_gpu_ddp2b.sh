#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612"
timeout 240 $TR tools/bench_train.py > gpurun_out/train_bench_n2.json 2> gpurun_out/train_bench_n2.err
echo "bench n2 rc=$?"; tail -2 gpurun_out/train_bench_n2.err; cut -c1-420 gpurun_out/train_bench_n2.json
