#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/tests_all.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/tests_all.log; grep -n "Error\|FAILED" gpurun_out/tests_all.log | head
MRD_BENCH_WATCHDOG=500 timeout 600 python bench.py --gpus 1 --profile-out gpurun_out/profile_b4096.csv > gpurun_out/bench_b4096.json 2> gpurun_out/bench_b4096.err
echo "bench rc=$?"; tail -2 gpurun_out/bench_b4096.err; cut -c1-250 gpurun_out/bench_b4096.json
timeout 300 python tools/bench_train.py > gpurun_out/train_bench.json 2> gpurun_out/train_bench.err
echo "train bench rc=$?"; cut -c1-330 gpurun_out/train_bench.json
