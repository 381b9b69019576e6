#!/bin/bash
# first GPU contact: kernel unit tests, one process per kernel family so a trap in one family does
# not poison the CUDA context of the others
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for fam in gemm conv stem "maxpool or avgpool or layernorm or bert_embed or mask" attention; do
  echo "=== $fam" | tee -a gpurun_out/kernels.log
  timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 100 -k "$fam" 2>&1 | tail -25 | tee -a gpurun_out/kernels.log
done
