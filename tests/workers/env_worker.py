"""Worker stub for tests/test_local_cuda_scheduler.py: dumps its environment contract, optionally fails."""
import json
import os
import sys
import time

out_dir = sys.argv[1]
mode = sys.argv[2] if len(sys.argv) > 2 else "ok"
keys = ["RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "GROUP_RANK", "GROUP_WORLD_SIZE", "ROLE_RANK", "ROLE_WORLD_SIZE",
        "ROLE_NAME", "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RESTART_COUNT", "TORCHELASTIC_MAX_RESTARTS", "TORCHELASTIC_RUN_ID",
        "TORCHELASTIC_USE_AGENT_STORE", "TORCHELASTIC_ERROR_FILE", "OMP_NUM_THREADS", "B2_SHM_NAME", "B2_EPOCH", "B2_DEVICE",
        "CUDA_VISIBLE_DEVICES", "TORCHX_JOB_ID", "TORCHX_RANK0_HOST", "TORCHX_TRACKING_RUN_NAME", "LOGLEVEL", "PYTHONUNBUFFERED", "MY_ENV"]
env = {k: os.environ.get(k) for k in keys}
env["affinity"] = sorted(os.sched_getaffinity(0))
env["argv"] = sys.argv[1:]
attempt = int(os.environ.get("TORCHELASTIC_RESTART_COUNT", "0"))
rank = int(os.environ.get("RANK", "0"))
with open(os.path.join(out_dir, f"attempt{attempt}_rank{rank}.json"), "w") as f:
    json.dump(env, f)
print(f"hello from rank {rank} attempt {attempt}", flush=True)
print(f"warn from rank {rank}", file=sys.stderr, flush=True)
if mode == "fail_rank1_first_attempt" and rank == 1 and attempt == 0:
    with open(os.environ["TORCHELASTIC_ERROR_FILE"], "w") as f:
        json.dump({"message": {"message": "boom on rank 1", "errorCode": 13, "extraInfo": {"timestamp": int(time.time())}}}, f)
    sys.exit(13)
if mode == "fail_rank1_always" and rank == 1:
    sys.exit(7)
if rank != 1 and (mode == "fail_rank1_always" or (mode == "fail_rank1_first_attempt" and attempt == 0)):
    time.sleep(60)  # survivors hang (as a worker blocked in a collective would) until the scheduler tears them down
if mode == "sleep":
    time.sleep(60)
