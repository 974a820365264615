#!/bin/bash
# One 8-GPU box visit: multi-device parity tests, allreduce sweeps at W=8/4/2 vs NCCL, bench.py at N=8 (both arms).
# Everything lands in gpurun_out/.  Each step has its own timeout so a hang cannot eat the box.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 -k "across_devices or per_gpu" > gpurun_out/pytest_gpu8.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu8.log; tail -3 gpurun_out/pytest_gpu8.log
for W in 8 4; do
  timeout 500 $TR --nproc-per-node $W --master-port $((29600+W)) tools/sweep_allreduce.py --ctas 0,32,128 --algos auto,oneshot,twoshot \
     --max-mib ${SWEEP_MAX_MIB:-1024} --out gpurun_out/sweep_w$W.jsonl > gpurun_out/sweep_w$W.log 2>&1
  echo "sweep W=$W rc=$?"; tail -2 gpurun_out/sweep_w$W.log | cut -c1-300
done
for N in 8; do
  timeout 400 $TR --nproc-per-node $N --master-port 29701 bench.py --impl reference --gpus $N --steps 30 --warmup 10 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
  echo "bench ref N=$N rc=$?"; cat gpurun_out/bench_ref_n$N.json | cut -c1-400
  timeout 400 $TR --nproc-per-node $N --master-port 29702 bench.py --gpus $N --steps 30 --warmup 10 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
  echo "bench b200 N=$N rc=$?"; cat gpurun_out/bench_n$N.json | cut -c1-400; tail -3 gpurun_out/bench_n$N.err
done
