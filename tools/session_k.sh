#!/bin/bash
# Round 2, session K (2 GPUs): the final tree - whole GPU suite (1-GPU topologies + real 2-GPU peers), smoke, bench N=1 and N=2.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 100 python __graft_entry__.py smoke ) > gpurun_out/k_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/k_smoke.log | head -1
( time timeout 400 python -m pytest tests -m gpu -q --timeout 200 ) > gpurun_out/k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/k_pytest.log; tail -6 gpurun_out/k_pytest.log
( timeout 150 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/k_bench_n1.json 2> gpurun_out/k_bench_n1.err ); echo "bench n1 rc=$?"; grep -o '"roofline": {[^}]*}' gpurun_out/k_bench_n1.json | cut -c1-700
( timeout 150 $TR --nproc-per-node 2 --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 5 --no-e2e > gpurun_out/k_bench_n2.json 2> gpurun_out/k_bench_n2.err ); echo "bench n2 rc=$?"; grep -o '"bit_exact": [a-z]*' gpurun_out/k_bench_n2.json
