#!/bin/bash
# Round 2, session A (2 GPUs): bring-up of the VMM/multicast arena, the pipelined kernels and the NVLS path.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/a_gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 150 python __graft_entry__.py smoke ) > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/a_smoke.log
( time timeout 420 python -m pytest tests/test_allreduce_gpu.py tests/test_ipc_gpu.py -q --timeout 150 -x ) > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log; tail -5 gpurun_out/a_pytest.log
B2_VERBOSE=1 timeout 120 python tools/nvls_probe.py --world 2 --out gpurun_out/nvls_probe_w2.npz > gpurun_out/a_probe.log 2>&1; echo "probe rc=$?"; tail -3 gpurun_out/a_probe.log
B2_VERBOSE=1 timeout 300 $TR --nproc-per-node 2 --master-port 29611 tools/sweep_allreduce.py --sizes-mib 1,4,7.82,16,25.04,30.04,64,168.27,256,1024 \
   --variants "auto;twoshot;twoshot_pipe;twoshot_pipe:chunk=512;twoshot_pipe:chunk=8192;twoshot_pipe:ctas=128;nvls;oneshot" --trace --skip-f32 --nvlink-counters \
   --out gpurun_out/a_sweep_w2.jsonl > gpurun_out/a_sweep_w2.log 2>&1
echo "sweep rc=$?"; tail -2 gpurun_out/a_sweep_w2.log | cut -c1-600
