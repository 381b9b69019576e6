#!/bin/bash
# full ncu capture (with source counters) of the stem + max-pool launch
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --global-batch 1024 --no-cpu-baseline --no-e2e --no-other-configs"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:conv_gemm_kernelILi64ELi1E -s 6 -c 1 -f -o gpurun_out/r2_prof_stem $CMD > gpurun_out/r2_ncu_stem.log 2>&1
echo "stem full rc=$?"; tail -2 gpurun_out/r2_ncu_stem.log
