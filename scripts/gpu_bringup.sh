#!/bin/bash
# Kernel bring-up on a B200 box: every probe in its own process, bounded by timeout; logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
LOG=gpurun_out/bringup.log
: > $LOG
for t in ${@:-gemm_basic gemm_shapes gemm_persistent gemm_epilogue conv_basic conv_shapes norms temporal elementwise spatial_basic spatial_shapes perf}; do
  echo "##### $t" >> $LOG
  timeout 240 python scripts/gpu_probe.py $t >> $LOG 2>&1
  echo "exit=$?" >> $LOG
done
grep -E "^\[|^==|exit=|#####|TFLOP|GB/s" $LOG | tail -150
