#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for cfg in "128 131072" "64 131072" "32 131072" "16 131072" "256 131072"; do
  set -- $cfg
  python bench.py --steps 3 --warmup 3 --global-batch 1024 --no-cpu-baseline --no-e2e --img-chunk $1 --tok-chunk $2 --profile-out gpurun_out/prof_$1_$2.csv 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('img_chunk $1 tok_chunk $2 value', round(d['value']), 'fam', d['kernel_families'])"
  grep -E "layer1.conv3|layer1.conv1|layer2.conv3|layer1.conv2_3x3|layer4.conv2_3x3," gpurun_out/prof_$1_$2.csv | cut -d, -f1,3,4,5,6 | tr '\n' ' '; echo
done
