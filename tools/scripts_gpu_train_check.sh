#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu --timeout 300 > gpurun_out/train_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/train_tests.log; grep -n "Error" gpurun_out/train_tests.log | head -5
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | grep -v Warn | tail -2
timeout 600 python tools/bench_train.py --profile-out gpurun_out/train_profile.csv > gpurun_out/train_bench.json 2> gpurun_out/train_bench.err
echo "train bench rc=$?"; tail -2 gpurun_out/train_bench.err; cut -c1-1600 gpurun_out/train_bench.json
