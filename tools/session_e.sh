#!/bin/bash
# Round 2, session E (8 GPUs): final W=8 sweep vs NCCL, bench both arms (ResNet-50, GPT-2-small), attribution arm, -j 2x4.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 8 --master-port 29621 tools/sweep_allreduce.py --sizes-mib 0.5,2,4,7.82,25.04,30.04,64,168.27,256,1024 \
   --variants "auto;twoshot;twoshot_ll;twoshot_ll:ctas=128;nvls:chunk=8192" --skip-f32 --check-variants --nvlink-counters --trace \
   --out gpurun_out/e_sweep_w8.jsonl > gpurun_out/e_sweep_w8.log 2>&1
echo "sweep8 rc=$?"; tail -1 gpurun_out/e_sweep_w8.log | cut -c1-200
export B2_LL_MIN_BYTES=1048576 B2_NVLS_MIN_BYTES=1099511627776   # ours: LL two-shot from 1 MiB of wire data, NVLS off (provisional AUTO for this session)
for arm in b200 reference; do
  ( timeout 200 $TR --nproc-per-node 8 --master-port 29623 bench.py --gpus 8 --steps 20 --warmup 5 --impl $arm > gpurun_out/e_bench_rn50_${arm}_n8.json 2> gpurun_out/e_bench_rn50_${arm}_n8.err ); echo "rn50 $arm rc=$?"
  grep -o '"value": [0-9.]*, "unit"' gpurun_out/e_bench_rn50_${arm}_n8.json | head -2
done
for arm in b200 reference; do
  ( timeout 200 $TR --nproc-per-node 8 --master-port 29625 bench.py --gpus 8 --steps 12 --warmup 4 --model gpt2 --no-e2e --impl $arm > gpurun_out/e_bench_gpt2_${arm}_n8.json 2> gpurun_out/e_bench_gpt2_${arm}_n8.err ); echo "gpt2 $arm rc=$?"
  grep -o '"value": [0-9.]*, "unit"' gpurun_out/e_bench_gpt2_${arm}_n8.json | head -2; tail -2 gpurun_out/e_bench_gpt2_${arm}_n8.err
done
( timeout 90 python -m torchx_b200.cli.main run -s local_cuda -cfg log_dir=/tmp/j2x4 dist.ddp -j 2x4 --script examples/train_ddp.py -- --steps 20 ) > gpurun_out/e_j2x4.log 2>&1; echo "j2x4 rc=$?"; grep -c "params sha256" gpurun_out/e_j2x4.log; grep -o "sha256 [0-9a-f]*" gpurun_out/e_j2x4.log | sort | uniq -c
