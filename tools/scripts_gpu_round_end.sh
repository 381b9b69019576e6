#!/bin/bash
# what the driver runs at round end (gpu tests, smoke, reference arm, default bench) + the training bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/tests_all.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/tests_all.log; grep -n "Error\|FAILED" gpurun_out/tests_all.log | head
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | grep -v Warn | tail -2
timeout 600 python tools/bench_train.py --profile-out gpurun_out/train_profile.csv > gpurun_out/train_bench.json 2> gpurun_out/train_bench.err
echo "train bench rc=$?"; tail -2 gpurun_out/train_bench.err; cut -c1-330 gpurun_out/train_bench.json
timeout 600 python tools/bench_train.py --bn-eval --fused-adamw > gpurun_out/train_bench_bneval.json 2> gpurun_out/train_bench2.err
echo "train bench2 rc=$?"; cut -c1-330 gpurun_out/train_bench_bneval.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 2>/dev/null | cut -c1-300
MRD_BENCH_WATCHDOG=500 timeout 600 python bench.py --gpus 1 --profile-out gpurun_out/profile_b4096.csv > gpurun_out/bench_b4096.json 2> gpurun_out/bench_b4096.err
echo "bench rc=$?"; tail -2 gpurun_out/bench_b4096.err; cut -c1-250 gpurun_out/bench_b4096.json
