#!/bin/bash
# training tests + training-step bench with per-label profile
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -s --timeout 300 > gpurun_out/train_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/train_tests.log; grep -n "running-stat\|batch-stat" gpurun_out/train_tests.log | cut -c1-600
timeout 600 python tools/bench_train.py --profile-out gpurun_out/train_profile.csv > gpurun_out/train_bench.json 2> gpurun_out/train_bench.err
echo "bench rc=$?"; tail -5 gpurun_out/train_bench.err; cat gpurun_out/train_bench.json
timeout 600 python tools/bench_train.py --bn-eval --fused-adamw > gpurun_out/train_bench_bneval.json 2> gpurun_out/train_bench2.err
echo "bench2 rc=$?"; tail -3 gpurun_out/train_bench2.err; cut -c1-700 gpurun_out/train_bench_bneval.json
