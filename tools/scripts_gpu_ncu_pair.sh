#!/bin/bash
# full ncu capture of the BERT GEMM launches (cta_group::2 variants) of one bench step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --global-batch 1024 --no-cpu-baseline --no-e2e --no-other-configs"
$CMD > gpurun_out/r2_ncu_plain2.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_ncu_plain2.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 528 -c 14 -f -o gpurun_out/r2_prof_pair $CMD > gpurun_out/r2_ncu_pair.log 2>&1
echo "pair full rc=$?"; tail -3 gpurun_out/r2_ncu_pair.log
