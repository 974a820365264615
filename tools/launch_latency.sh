#!/bin/bash
# Launch latency, both launchers, same worker script: wall time of
#   torchx run -s <sched> dist.ddp -j 1x1 --script examples/train_ddp.py -- --steps 1
# from CLI start to job completion (python start-up, scheduling, rendezvous, CUDA context, one training step).
mkdir -p gpurun_out
: > gpurun_out/launch_latency.txt
for S in local_cuda local_cwd local_cuda local_cwd; do
  t0=$(date +%s.%N)
  python -m torchx_b200.cli.main run -s $S -cfg log_dir=/tmp/ll_$RANDOM dist.ddp -j 1x1 --script examples/train_ddp.py -- --steps 1 > /dev/null 2> gpurun_out/ll_$S.err
  rc=$?
  t1=$(date +%s.%N)
  echo "$S rc=$rc wall_s=$(python -c "print(round($t1-$t0,2))")" | tee -a gpurun_out/launch_latency.txt
done
