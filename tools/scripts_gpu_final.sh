#!/bin/bash
# evidence run for profiles/: tests, round-end bench (defaults), ncu launch list + dram traffic + full capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | grep -v Warn | tail -1
timeout 600 python bench.py --gpus 1 --profile-out gpurun_out/profile_b4096.csv > gpurun_out/bench_b4096.json 2> gpurun_out/bench_b4096.err
echo "bench rc=$?"; tail -2 gpurun_out/bench_b4096.err
CMD="python bench.py --steps 1 --warmup 3 --global-batch 512 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 420 -c 8 -o gpurun_out/prof_gemm -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
