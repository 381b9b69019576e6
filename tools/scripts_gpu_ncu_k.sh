#!/bin/bash
# full ncu capture of selected launches: KREGEX=... SKIP=n COUNT=n TAG=name GB=batch
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --global-batch ${GB:-256} --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s ${SKIP:-20} -c ${COUNT:-2} -o gpurun_out/prof_${TAG:-k} -f $CMD > gpurun_out/ncu_k.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_k.log
