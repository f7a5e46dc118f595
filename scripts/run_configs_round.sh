#!/bin/bash
# BASELINE.json configs[3] (768x768 x 48 frames, 30 steps, bf16) and configs[4] (512x512 x 64-frame edit) on the 8-GPU node.
N=${1:-8}
mkdir -p gpurun_out
for C in 4 5; do
  echo "== bench N=$N config $C"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29900+C)) \
      bench.py --gpus $N --config $C --steps 1 --warmup 3 --no-cpu-baseline --no-breakdown \
      > gpurun_out/r2_bench_n${N}_cfg$C.json 2> gpurun_out/r2_bench_n${N}_cfg$C.err
  echo "rc=$?"; head -c 900 gpurun_out/r2_bench_n${N}_cfg$C.json; echo; grep -v "Warning\|warn" gpurun_out/r2_bench_n${N}_cfg$C.err | tail -4 | cut -c1-300
done
