#!/bin/bash
# round-2 evidence with the final build (the determinism stress is tools/stress_*.py, run separately): ncu launch list (time + DRAM bytes) of one bench run,
# full captures of the dominant kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --global-batch 512 --no-cpu-baseline --no-e2e --no-other-configs"
$CMD > gpurun_out/r2_ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_ncu_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_list.log 2>&1
echo "list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 330 -c 16 -f -o gpurun_out/r2_prof_gemm $CMD > gpurun_out/r2_ncu_full.log 2>&1
echo "gemm full rc=$?"
ls -la gpurun_out | grep "r2_prof_gemm\|r2_launches"
