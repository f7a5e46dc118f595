#!/bin/bash
# Round-end evidence on one B200 (everything lands in gpurun_out/; summaries are copied to profiles/ afterwards):
#   1. pytest -m gpu   2. bench.py (default flags)   3. ncu launch list of one clip   4. ncu --set full of the hot kernels
mkdir -p gpurun_out
TAG=${1:-r01}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit=$?"
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit=$?"
cut -c1-300 gpurun_out/${TAG}_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --one-clip > gpurun_out/${TAG}_launches.log 2>&1; echo "ncu list exit=$?"
python scripts/ncu_launchlist.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.txt 2>&1
head -20 gpurun_out/${TAG}_launches_summary.txt
gzip -f gpurun_out/${TAG}_launches.csv
bash scripts/ncu_capture.sh qkv:gemm_tcgen05_kernel gemm:gemm_tcgen05_kernel geglu:gemm_tcgen05_kernel conv:gemm_tcgen05_kernel \
    attn:attn_spatial_pp_kernel attn80:attn_spatial_pp_kernel temporal:attn_temporal_kernel norm:gn_apply_kernel norm:layernorm5_kernel
