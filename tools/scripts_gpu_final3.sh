#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | grep -v Warn | tail -3
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_train_gpu.py -q -m gpu --timeout 300 -k "determin or batchnorm or amp or fixture or invarian" 2>&1 | tail -3
