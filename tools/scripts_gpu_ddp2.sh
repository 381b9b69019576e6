#!/bin/bash
# 2 GPUs: data-parallel training correctness (gradient bucket all-reduce over NCCL) and throughput
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR tools/check_ddp_train.py > gpurun_out/ddp_check.json 2> gpurun_out/ddp_check.err
echo "ddp check rc=$?"; tail -3 gpurun_out/ddp_check.err; cat gpurun_out/ddp_check.json
timeout 600 $TR tools/bench_train.py > gpurun_out/train_bench_n2.json 2> gpurun_out/train_bench_n2.err
echo "bench n2 rc=$?"; tail -3 gpurun_out/train_bench_n2.err; cut -c1-420 gpurun_out/train_bench_n2.json
timeout 600 python tools/bench_train.py --per-gpu-batch 256 --steps 10 > gpurun_out/train_bench_b256.json 2> gpurun_out/train_bench_b256.err
echo "bench b256 rc=$?"; tail -3 gpurun_out/train_bench_b256.err; cut -c1-420 gpurun_out/train_bench_b256.json
