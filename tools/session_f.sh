#!/bin/bash
# Round 2, session F (4 GPUs): BASELINE config #5 - BERT-base dist.ddp -j 1x4, rank 1 killed at step 30.
mkdir -p gpurun_out
( timeout 240 python tools/elastic_recover.py --sched local_cuda --nproc 4 --model bert --steps 45 --fail-at-step 20 --log-dir /tmp/el --out gpurun_out/f_elastic_local_cuda.json ) > gpurun_out/f_elastic_local_cuda.log 2>&1; echo "local_cuda rc=$?"; tail -2 gpurun_out/f_elastic_local_cuda.log | cut -c1-900
( timeout 240 python tools/elastic_recover.py --sched local_cwd --nproc 4 --model bert --steps 45 --fail-at-step 20 --log-dir /tmp/el --out gpurun_out/f_elastic_local_cwd.json ) > gpurun_out/f_elastic_local_cwd.log 2>&1; echo "local_cwd rc=$?"; tail -2 gpurun_out/f_elastic_local_cwd.log | cut -c1-900
