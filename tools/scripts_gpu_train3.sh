#!/bin/bash
# everything: full gpu suite (the GEMM kernel gained split-K), training bench + profile
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/tests_all.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/tests_all.log
timeout 600 python tools/bench_train.py --profile-out gpurun_out/train_profile.csv > gpurun_out/train_bench.json 2> gpurun_out/train_bench.err
echo "bench rc=$?"; tail -5 gpurun_out/train_bench.err; cat gpurun_out/train_bench.json
timeout 600 python tools/bench_train.py --bn-eval --fused-adamw > gpurun_out/train_bench_bneval.json 2> gpurun_out/train_bench2.err
echo "bench2 rc=$?"; tail -3 gpurun_out/train_bench2.err; cut -c1-500 gpurun_out/train_bench_bneval.json
