#!/bin/bash
# Round 2, session H (1 GPU): what the driver runs at round end - the whole GPU suite, smoke, bench N=1 - on the 1-GPU topology.
mkdir -p gpurun_out
( time timeout 150 python __graft_entry__.py smoke ) > gpurun_out/h_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/h_smoke.log
( time timeout 700 python -m pytest tests -m gpu -q --timeout 300 ) > gpurun_out/h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/h_pytest.log; tail -8 gpurun_out/h_pytest.log
( timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/h_bench_n1.json 2> gpurun_out/h_bench_n1.err ); echo "bench rc=$?"; tail -c 1200 gpurun_out/h_bench_n1.json; tail -2 gpurun_out/h_bench_n1.err
