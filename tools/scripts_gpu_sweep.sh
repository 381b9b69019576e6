#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for cfg in "64 16384" "32 16384" "16 16384" "128 16384" "64 32768" "64 65536" "64 131072" "32 65536"; do
  set -- $cfg
  python bench.py --steps 3 --warmup 3 --global-batch 1024 --no-cpu-baseline --no-e2e --img-chunk $1 --tok-chunk $2 --profile-out gpurun_out/prof_$1_$2.csv 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('img_chunk $1 tok_chunk $2 value', round(d['value']), 'fam', d['kernel_families'])"
done
