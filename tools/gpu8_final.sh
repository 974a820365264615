#!/bin/bash
# Final 8-GPU visit of the round: multi-device parity, AUTO sweeps at W=8/4 vs NCCL, bench.py both arms at N=8 and N=4.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 python -m pytest tests -m gpu -x -q --timeout 200 -k "across_devices or per_gpu or e2e or two_lanes" > gpurun_out/final_pytest_gpu8.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final_pytest_gpu8.log; tail -3 gpurun_out/final_pytest_gpu8.log
for W in 8 4; do
  timeout 300 $TR --nproc-per-node $W --master-port $((29600+W)) tools/sweep_allreduce.py --ctas 0 --algos auto --max-mib 1024 \
     --out gpurun_out/final_sweep_w$W.jsonl > gpurun_out/final_sweep_w$W.log 2>&1
  echo "sweep W=$W rc=$?"
done
for N in 8 4; do
  timeout 300 $TR --nproc-per-node $N --master-port 29701 bench.py --impl reference --gpus $N --steps 30 --warmup 10 > gpurun_out/final_bench_ref_n$N.json 2> gpurun_out/final_bench_ref_n$N.err
  echo "bench ref N=$N rc=$?"; cut -c1-200 gpurun_out/final_bench_ref_n$N.json
  timeout 300 $TR --nproc-per-node $N --master-port 29702 bench.py --gpus $N --steps 30 --warmup 10 > gpurun_out/final_bench_n$N.json 2> gpurun_out/final_bench_n$N.err
  echo "bench b200 N=$N rc=$?"; cut -c1-200 gpurun_out/final_bench_n$N.json; tail -2 gpurun_out/final_bench_n$N.err
done
