#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | grep -v Warn | tail -3
python bench.py --steps 3 --warmup 3 --global-batch ${GB:-1024} --profile-out gpurun_out/profile_r01.csv ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "rc=$?"; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json; echo; column -s, -t gpurun_out/profile_r01.csv | head -50
