#!/bin/bash
# full ncu capture of one launch each of the non-GEMM-family and special-mode kernels (flat 3x3 on CTA pairs, stem +
# pool, chained bottleneck tail, attention, LayerNorm, repack, fused tail)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --global-batch 1024 --no-cpu-baseline --no-e2e --no-other-configs"
i=0
for rx in "conv_gemm_kernelILi128ELi2E" "conv_gemm_kernelILi64ELi2E" "conv_gemm_kernelILi64ELi1E" "conv_chain_kernelILi64E" "attention_tc_kernel" "layernorm_kernelILi3E" "repack_images_kernel" "tail_fused_kernel" "global_avgpool"; do
  i=$((i+1))
  timeout 120 ncu --set full --clock-control none --kernel-name-base mangled -k regex:$rx -s 3 -c 1 -f -o gpurun_out/r2_prof_misc_$i $CMD > gpurun_out/r2_ncu_misc_$i.log 2>&1
  echo "$rx rc=$?"
done
ls -la gpurun_out | grep r2_prof_misc
