#!/bin/bash
# what the driver runs at round end: gpu tests, smoke, reference arm, bench (defaults)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | grep -v Warn | tail -2
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 2>/dev/null | cut -c1-400
MRD_BENCH_WATCHDOG=500 timeout 600 python bench.py --gpus 1 --profile-out gpurun_out/profile_b4096.csv > gpurun_out/bench_b4096.json 2> gpurun_out/bench_b4096.err
echo "rc=$?"; tail -3 gpurun_out/bench_b4096.err; cat gpurun_out/bench_b4096.json
