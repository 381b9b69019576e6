#!/bin/bash
# branch overlap experiment: correctness test + bench sweep over the SM split
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "overlap" --timeout 200 2>&1 | tail -3
for cfg in "0 0" "1 74" "1 56" "1 92"; do
  set -- $cfg
  timeout 300 python bench.py --gpus 1 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --overlap $1 --overlap-cnn-sms $2 > gpurun_out/bench_ov_$1_$2.json 2> gpurun_out/bench_ov.err
  echo "overlap=$1 cnn_sms=$2 rc=$?"; python -c "
import json,sys
d=json.load(open('gpurun_out/bench_ov_$1_$2.json')); print(round(d['value']), d['ms_per_step'], d['clocks'])"
done
