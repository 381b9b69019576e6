#!/bin/bash
# Determinism stress: single launches (stress_kernels / stress_patterns), then the whole forward under every fusion switch
mkdir -p gpurun_out
{
echo "=== stress_kernels (default switches)"; timeout 300 python tools/stress_kernels.py 60
echo "=== stress_patterns"; timeout 300 python tools/stress_patterns.py 60
echo "=== stress_determinism"; timeout 600 python tools/stress_determinism.py 16 1024
} > gpurun_out/r2_stress3.log 2>&1
grep -n "===\|OK  \|FAIL\|bad launches\|deterministic" gpurun_out/r2_stress3.log
