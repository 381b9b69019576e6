#!/bin/bash
# training tests (incl. S > 128) + a full ncu capture of the backward's own kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_abi.py -q -m "gpu or not gpu" -s --timeout 300 > gpurun_out/train_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/train_tests.log; grep -n "global rel-L2" gpurun_out/train_tests.log | cut -c1-300
grep -n "Error\|error:" gpurun_out/train_tests.log | head -10
TCMD="python tools/bench_train.py --steps 1 --warmup 2 --bn-eval"
ncu --set full --clock-control none --import-source on -k regex:'attention_bwd_kernel|transpose_pad_kernel|ln_bwd_kernel|colsum_bf16_kernel|gelu_kernel' -s 30 -c 10 -o gpurun_out/prof_train_bwd $TCMD > gpurun_out/ncu_train_full2.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_train_full2.log
