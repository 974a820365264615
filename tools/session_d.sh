#!/bin/bash
# Round 2, session D (2 GPUs): batched signallers + sentinel NVLS output buffers + segment table in kernel parameters.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 500 python -m pytest tests/test_allreduce_gpu.py tests/test_ipc_gpu.py tests/test_ddp_gpu.py tests/test_hook_multirank_gpu.py -q --timeout 200 ) > gpurun_out/d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/d_pytest.log; tail -8 gpurun_out/d_pytest.log
timeout 240 $TR --nproc-per-node 2 --master-port 29611 tools/sweep_allreduce.py --sizes-mib 4,7.82,25.04,64,256,1024 \
   --variants "auto;twoshot;oneshot;twoshot_pipe;twoshot_pipe:chunk=512;twoshot_pipe:chunk=8192;nvls;nvls:chunk=512;nvls:chunk=8192" --trace --skip-f32 --check-variants \
   --out gpurun_out/d_sweep_w2.jsonl > gpurun_out/d_sweep_w2.log 2>&1
echo "sweep rc=$?"; tail -1 gpurun_out/d_sweep_w2.log | cut -c1-300
( timeout 200 $TR --nproc-per-node 2 --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 5 > gpurun_out/d_bench_n2.json 2> gpurun_out/d_bench_n2.err ); echo "bench rc=$?"; tail -c 1500 gpurun_out/d_bench_n2.json; tail -3 gpurun_out/d_bench_n2.err
