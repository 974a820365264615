#!/bin/bash
# Round 2, session G (2 GPUs): the barrier-free LL two-shot.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 150 python __graft_entry__.py smoke ) > gpurun_out/g_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/g_smoke.log
( time timeout 500 python -m pytest tests/test_allreduce_gpu.py tests/test_ipc_gpu.py -q --timeout 200 ) > gpurun_out/g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/g_pytest.log; tail -8 gpurun_out/g_pytest.log
timeout 200 $TR --nproc-per-node 2 --master-port 29611 tools/sweep_allreduce.py --sizes-mib 0.25,1,4,7.82,25.04,64,256,1024 \
   --variants "auto;twoshot;oneshot;twoshot_ll;twoshot_ll:ctas=128;twoshot_ll:ctas=32" --trace --skip-f32 --check-variants --nvlink-counters \
   --out gpurun_out/g_sweep_w2.jsonl > gpurun_out/g_sweep_w2.log 2>&1
echo "sweep rc=$?"; tail -1 gpurun_out/g_sweep_w2.log | cut -c1-300
