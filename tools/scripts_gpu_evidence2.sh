#!/bin/bash
# session-2 evidence: default bench (regression check after the split-K change), ncu launch list of one
# training step, full captures of the training step's own kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
MRD_BENCH_WATCHDOG=500 timeout 600 python bench.py --gpus 1 --profile-out gpurun_out/profile_b4096.csv > gpurun_out/bench_b4096.json 2> gpurun_out/bench_b4096.err
echo "bench rc=$?"; tail -2 gpurun_out/bench_b4096.err; cut -c1-600 gpurun_out/bench_b4096.json
TCMD="python tools/bench_train.py --steps 1 --warmup 2"
$TCMD > gpurun_out/ncu_train_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_train.csv $TCMD > gpurun_out/ncu_train_list.log 2>&1
echo "train list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'attention_bwd_kernel|transpose_pad_kernel|ln_bwd_kernel|bn_stats_kernel|embed_ln_bwd_kernel' -s 40 -c 8 -o gpurun_out/prof_train $TCMD > gpurun_out/ncu_train_full.log 2>&1
echo "train full rc=$?"
tail -2 gpurun_out/ncu_train_list.log gpurun_out/ncu_train_full.log
ls -la gpurun_out/prof_train.ncu-rep gpurun_out/launches_train.csv
