#!/bin/bash
# Round 2, session C (8 GPUs): the kernels at W=8 - parity on real NVSwitch, the NVLS rounding/order probe, the variant
# sweep against NCCL, one counter capture, one in-step bench with the parity key.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | head -8 > gpurun_out/c_gpus.txt 2>&1
( time timeout 400 python -m pytest tests/test_allreduce_gpu.py tests/test_ipc_gpu.py tests/test_hook_multirank_gpu.py tests/test_ddp_gpu.py -q --timeout 180 \
    -k "((across_devices or larger_than_a_stage or one_process_per_gpu or one_gpu_per_rank) and (8 or 4)) or zero_copy or state_dict" ) > gpurun_out/c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c_pytest.log; tail -6 gpurun_out/c_pytest.log
B2_VERBOSE=1 timeout 120 python tools/nvls_probe.py --world 8 --out gpurun_out/nvls_probe_w8.npz > gpurun_out/c_probe8.log 2>&1; echo "probe8 rc=$?"; tail -2 gpurun_out/c_probe8.log
B2_VERBOSE=1 timeout 90 python tools/nvls_probe.py --world 2 --out gpurun_out/nvls_probe_w2b.npz > gpurun_out/c_probe2.log 2>&1; echo "probe2 rc=$?"; tail -2 gpurun_out/c_probe2.log
timeout 90 $TR --nproc-per-node 2 --master-port 29618 tools/nvls_probe.py --out gpurun_out/nvls_probe_mp_w2.npz > gpurun_out/c_probe_mp2.log 2>&1; echo "probe mp2 rc=$?"; grep "multi-process" gpurun_out/c_probe_mp2.log
timeout 90 $TR --nproc-per-node 8 --master-port 29619 tools/nvls_probe.py --out gpurun_out/nvls_probe_mp_w8.npz > gpurun_out/c_probe_mp8.log 2>&1; echo "probe mp8 rc=$?"; grep "multi-process" gpurun_out/c_probe_mp8.log
timeout 420 $TR --nproc-per-node 8 --master-port 29621 tools/sweep_allreduce.py --sizes-mib 0.25,1,4,7.82,9.27,16,25.04,30.04,64,128,168.27,256,512,1024 \
   --variants "auto;twoshot;oneshot;twoshot_pipe;twoshot_pipe:chunk=512;nvls;nvls:chunk=512;nvls:chunk=8192;nvls:ctas=128;nvls:ctas=32:chunk=1024;nvls:ctas=148:chunk=1024" \
   --trace --skip-f32 --check-variants --nvlink-counters --out gpurun_out/c_sweep_w8.jsonl > gpurun_out/c_sweep_w8.log 2>&1
echo "sweep8 rc=$?"; tail -1 gpurun_out/c_sweep_w8.log | cut -c1-300
timeout 240 bash tools/ncu_multirank.sh 8 25.04 auto w8_25 > gpurun_out/c_ncu.log 2>&1; echo "ncu rc=$?"; tail -4 gpurun_out/c_ncu.log
( timeout 240 $TR --nproc-per-node 8 --master-port 29623 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/c_bench_n8.json 2> gpurun_out/c_bench_n8.err ); echo "bench rc=$?"; cut -c1-400 gpurun_out/c_bench_n8.json; tail -3 gpurun_out/c_bench_n8.err
