#!/bin/bash
# Single-GPU profiling visit (B200_PROFILING.md recipe).  Never wrap a multi-rank command in ncu.
#   1. launch list of the bench.py timed region (every kernel with its device time; compare SHARES)
#   2. one `--set full` capture of the W=1 bucket kernel (k_local_pass) at a ResNet-50 bucket size
mkdir -p gpurun_out
export BENCH_CUDA_PROFILER=1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r01_launches_bench_n1.csv python bench.py --steps 2 --warmup 4 --no-e2e --no-cpu-baseline \
    > gpurun_out/r01_bench_under_ncu.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r01_launches_bench_n1.csv)"
unset BENCH_CUDA_PROFILER
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_local_pass -s 6 -c 3 -f -o gpurun_out/r01_local_pass \
    python tools/microbench_local.py --sizes-mib 30.04 --iters 10 > gpurun_out/r01_local_pass_ncu.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/*.ncu-rep 2>/dev/null
