#!/bin/bash
# One `ncu --set full` capture per hot kernel (GPU box, 1 GPU). Reports land in gpurun_out/ncu_*.ncu-rep.
mkdir -p gpurun_out
SPECS=("$@")
if [ ${#SPECS[@]} -eq 0 ]; then SPECS=("gemm:gemm_tcgen05_kernel" "geglu:gemm_tcgen05_kernel" "conv:gemm_tcgen05_kernel" "attn:attn_spatial_pp_kernel" "attn80:attn_spatial_pp_kernel" "norm:gn_apply_kernel" "norm:layernorm_kernel" "temporal:attn_temporal_kernel"); fi
for spec in "${SPECS[@]}"; do
  fn=${spec%%:*}; kn=${spec##*:}
  out=gpurun_out/ncu_${fn}_${kn}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$kn -s 2 -c 1 -f -o $out \
      python scripts/ncu_kernels.py $fn > gpurun_out/ncu_${fn}_${kn}.log 2>&1
  echo "$spec exit=$?"
done
ls -la gpurun_out/*.ncu-rep
