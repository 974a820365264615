#!/bin/bash
# Hardware counters for the multi-GPU collective kernels: rank 0 under ncu, ranks 1..W-1 free (one process per GPU).
# usage: tools/ncu_multirank.sh W MIB ALGO TAG
# ncu cannot serialise co-operating spin kernels, so (a) only rank 0 is profiled, (b) every metric group is chosen to fit
# ONE pass (no kernel replay: a replayed kernel would run without its peers), (c) all ranks idle after the profiled launch.
W=${1:-8}; MIB=${2:-25.04}; ALGO=${3:-auto}; TAG=${4:-w${W}_${MIB}}
mkdir -p gpurun_out
KREGEX='regex:k_pipe|k_twoshot|k_oneshot'
run_group () {  # $1 = group name, $2 = metric list
  local shm="/b2_ncu_$$_$1"
  for r in $(seq 1 $((W-1))); do
    RANK=$r WORLD_SIZE=$W LOCAL_RANK=$r B2_SHM_NAME=$shm timeout 300 python tools/one_allreduce.py --mib $MIB --algo $ALGO > gpurun_out/ncu_${TAG}_$1_r$r.log 2>&1 &
  done
  RANK=0 WORLD_SIZE=$W LOCAL_RANK=0 B2_SHM_NAME=$shm timeout 300 ncu --metrics "$2" --clock-control none --cache-control none \
      -k "$KREGEX" --launch-skip 3 --launch-count 1 --csv --log-file gpurun_out/ncu_${TAG}_$1.csv \
      python tools/one_allreduce.py --mib $MIB --algo $ALGO > gpurun_out/ncu_${TAG}_$1_r0.log 2>&1
  echo "ncu $TAG $1 rc=$?"
  wait
}
run_group dram "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
run_group lts  "lts__t_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum"
NVL=$(ncu --query-metrics 2>/dev/null | grep -i -o '^nvl[a-z_0-9]*bytes[a-z_0-9]*' | sort -u | head -6 | sed 's/$/.sum/' | paste -sd, -)
echo "nvlink metrics on this box: $NVL" > gpurun_out/ncu_${TAG}_nvl_metrics.txt
ncu --query-metrics 2>/dev/null | grep -i 'nvl' | head -40 >> gpurun_out/ncu_${TAG}_nvl_metrics.txt
if [ -n "$NVL" ]; then run_group nvl "$NVL"; fi
