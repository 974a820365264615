#!/bin/bash
# Round 2, session J (2 GPUs): the final tree - AUTO thresholds as shipped, the bench line at N=2 with the parity / NCCL comparison.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( timeout 200 $TR --nproc-per-node 2 --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 5 > gpurun_out/j_bench_n2.json 2> gpurun_out/j_bench_n2.err ); echo "bench rc=$?"; tail -c 2200 gpurun_out/j_bench_n2.json; tail -3 gpurun_out/j_bench_n2.err
( time timeout 300 python -m pytest tests/test_allreduce_gpu.py tests/test_ipc_gpu.py -q --timeout 200 -k "across_devices or per_gpu or larger_than or auto" ) > gpurun_out/j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j_pytest.log; tail -4 gpurun_out/j_pytest.log
