#!/bin/bash
# training-step tests with full output (no -x: one call should show every failure)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -s --timeout 300 > gpurun_out/train_tests.log 2>&1
echo "rc=$?"
grep -E "passed|failed|Error|error|assert|worst|losses|AdamW" gpurun_out/train_tests.log | head -60
