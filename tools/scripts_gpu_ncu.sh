#!/bin/bash
# ncu evidence: launch list of one bench step + full capture of the dominant kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --global-batch 256 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 600 -c 6 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
ls -la gpurun_out | head -30
tail -3 gpurun_out/ncu_list.log gpurun_out/ncu_full.log
