mkdir -p gpurun_out; nvidia-smi -L | wc -l
(timeout 200 python -m pytest tests/test_exchange_gpu.py -q -m gpu -s -k "cfg_branch" > gpurun_out/r2_cfgsplit2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_cfgsplit2.log); tail -3 gpurun_out/r2_cfgsplit2.log | cut -c1-400
scripts/run_mgpu.sh 8 3
for N in 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800+N)) bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
  echo "bench N=$N rc=$?"; head -c 400 gpurun_out/r2_bench_n$N.json; echo; tail -2 gpurun_out/r2_bench_n$N.err | cut -c1-300
done
