#!/bin/bash
# session-2 check: gpu tests (incl. all_hidden export) + configs[1]/[2] throughput
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -6 | tee gpurun_out/tests.log
timeout 600 python tools/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err
echo "rc=$?"; tail -3 gpurun_out/configs.err; cat gpurun_out/configs.json
