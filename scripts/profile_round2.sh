#!/bin/bash
# Round-2 profile evidence on ONE B200 (everything lands in gpurun_out/; summaries are copied to profiles/ afterwards):
#   1. ncu launch list of one clip (time + DRAM bytes per launch)  ->  r02_ncu_launches_summary.txt, ncu_traffic.json
#   2. ncu --set full of the hot kernels at hot-path shapes        ->  r02_ncu_full_summary.txt
mkdir -p gpurun_out
TAG=${1:-r02}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --one-clip > gpurun_out/${TAG}_launches.log 2>&1; echo "ncu list exit=$?"
python scripts/ncu_launchlist.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_ncu_launches_summary.txt 2>&1
head -16 gpurun_out/${TAG}_ncu_launches_summary.txt
python scripts/ncu_traffic.py gpurun_out/${TAG}_ncu_launches_summary.txt 1 gpurun_out/ncu_traffic.json
gzip -f gpurun_out/${TAG}_launches.csv
bash scripts/ncu_capture.sh attn:attn_spatial_pp2_kernel attn80:attn_spatial_pp_kernel geglu:gemm_tcgen05_kernel \
    conv:gemm_tcgen05_kernel ffo:gemm_tcgen05_kernel gemm:gemm_tcgen05_kernel convup:gemm_tcgen05_kernel \
    norm:gn_apply_kernel norm:gn_stats_kernel norm:layernorm5_kernel temporal:attn_temporal_kernel > gpurun_out/${TAG}_ncu_capture.log 2>&1
python scripts/ncu_summary.py gpurun_out/ncu_*.ncu-rep > gpurun_out/${TAG}_ncu_full_summary.txt 2>&1
grep -E "^==|time |tensor %|xu %|dram %|dram rd|dram wr" gpurun_out/${TAG}_ncu_full_summary.txt | head -80
rm -f gpurun_out/ncu_*.ncu-rep
