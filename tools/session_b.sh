#!/bin/bash
# Round 2, session B (2 GPUs): decoupled-signaller pipeline, per-variant parity incl. multi-launch messages, new DDP path.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 600 python -m pytest tests/test_allreduce_gpu.py tests/test_ipc_gpu.py tests/test_ddp_gpu.py tests/test_hook_multirank_gpu.py -q --timeout 200 ) > gpurun_out/b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b_pytest.log; tail -15 gpurun_out/b_pytest.log
timeout 300 $TR --nproc-per-node 2 --master-port 29611 tools/sweep_allreduce.py --sizes-mib 4,7.82,25.04,30.04,64,168.27,256,1024 \
   --variants "auto;twoshot;twoshot_pipe;twoshot_pipe:chunk=512;twoshot_pipe:chunk=8192;nvls;nvls:chunk=512" --trace --skip-f32 --skip-nccl --check-variants \
   --out gpurun_out/b_sweep_w2.jsonl > gpurun_out/b_sweep_w2.log 2>&1
echo "sweep rc=$?"; tail -1 gpurun_out/b_sweep_w2.log | cut -c1-400
B2_STAGE_MB=32 timeout 200 $TR --nproc-per-node 2 --master-port 29612 tools/sweep_allreduce.py --sizes-mib 25.04,64,168.27 \
   --variants "twoshot;twoshot_pipe;nvls" --skip-f32 --skip-nccl --check-variants --out gpurun_out/b_sweep_w2_smallstage.jsonl > gpurun_out/b_sweep_w2_smallstage.log 2>&1
echo "sweep2 rc=$?"
( timeout 200 $TR --nproc-per-node 2 --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 5 > gpurun_out/b_bench_n2.json 2> gpurun_out/b_bench_n2.err ); echo "bench rc=$?"; cut -c1-300 gpurun_out/b_bench_n2.json; tail -3 gpurun_out/b_bench_n2.err
