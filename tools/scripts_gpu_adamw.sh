#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_train_gpu.py -q -m gpu --timeout 200 -k "fused_adamw or refuses or loop" 2>&1 | tail -4
timeout 300 python tools/bench_train.py --native-adamw > gpurun_out/train_bench_native_adamw.json 2> gpurun_out/train_bench3.err
echo "bench rc=$?"; tail -2 gpurun_out/train_bench3.err; cut -c1-420 gpurun_out/train_bench_native_adamw.json
timeout 300 python tools/bench_train.py --native-adamw --bn-eval > gpurun_out/train_bench_native_adamw_bneval.json 2> gpurun_out/train_bench4.err
echo "bench rc=$?"; cut -c1-330 gpurun_out/train_bench_native_adamw_bneval.json
