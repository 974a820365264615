#!/bin/bash
# Round 2, session I (1 GPU): re-check the two reworked test files, attribution arms and the GPT-2 W=1 step time.
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_ddp_gpu.py tests/test_hook_multirank_gpu.py tests/test_allreduce_gpu.py -q --timeout 200 -k "ddp or hook or auto" ) > gpurun_out/i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/i_pytest.log; tail -6 gpurun_out/i_pytest.log
for arm in reference reference_tuned; do
  ( timeout 150 python bench.py --gpus 1 --steps 20 --warmup 5 --impl $arm --no-cpu-baseline > gpurun_out/i_bench_rn50_${arm}_n1.json 2> gpurun_out/i_bench_rn50_${arm}_n1.err ); echo "rn50 $arm rc=$?"
  grep -o '"value": [0-9.]*, "unit"' gpurun_out/i_bench_rn50_${arm}_n1.json | head -2
done
for arm in b200 reference; do
  ( timeout 150 python bench.py --gpus 1 --steps 12 --warmup 4 --model gpt2 --no-e2e --no-cpu-baseline --impl $arm > gpurun_out/i_bench_gpt2_${arm}_n1.json 2> gpurun_out/i_bench_gpt2_${arm}_n1.err ); echo "gpt2 $arm rc=$?"
  grep -o '"value": [0-9.]*, "unit"' gpurun_out/i_bench_gpt2_${arm}_n1.json | head -1; tail -1 gpurun_out/i_bench_gpt2_${arm}_n1.err
done
