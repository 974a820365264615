"""BASELINE.json config #5, measured: `dist.ddp --nproc N` with one rank killed at step K.

  --sched local_cuda : ONE submission with --max_retries 1; the scheduler notices the dead worker, tears the gang down and
                       re-launches it under the next rendezvous epoch (no operator).
  --sched local_cwd  : the reference's path.  Its local scheduler never restarts anything (reference
                       torchx/schedulers/local_scheduler.py:555-563,1057: max_retries is ignored, a dead replica fails the
                       app), so recovery = an operator re-submitting.  Modelled at its best: this driver polls the status
                       every 50 ms and re-submits the moment the app is FAILED.
Reports kill -> first completed step after recovery (seconds, wall clock across processes on one box) and the steady-state
sequences/sec after recovery.  Stamps come from the workers' own ELASTIC lines (examples/train_elastic.py).
"""
import argparse
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from torchx_b200.runner import get_runner  # noqa: E402
from torchx_b200.specs import AppState, is_terminal  # noqa: E402


def collect(runner, handle, role, replicas):
    rows = []
    for k in range(replicas):
        try:
            for ln in runner.log_lines(handle, role, k):
                i = ln.find("ELASTIC ")
                if i >= 0:
                    try:
                        rows.append(json.loads(ln[i + 8:]))
                    except ValueError:
                        pass
        except Exception as e:  # noqa: BLE001
            print(f"[elastic] could not read logs of replica {k}: {e}", file=sys.stderr)
    return rows


def wait_terminal(runner, handle, poll=0.05, timeout=900):
    t0 = time.time()
    while time.time() - t0 < timeout:
        st = runner.status(handle)
        if st is not None and is_terminal(st.state):
            return st, time.time()
        time.sleep(poll)
    raise TimeoutError(handle)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sched", default="local_cuda", choices=["local_cuda", "local_cwd"])
    ap.add_argument("--nproc", type=int, default=4)
    ap.add_argument("--model", default="bert")
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--fail-at-step", type=int, default=30)
    ap.add_argument("--log-dir", default="/tmp/elastic_logs")
    ap.add_argument("--out", default="")
    ap.add_argument("--impl", default="", help="override the worker implementation (gloo = CPU plumbing check)")
    a = ap.parse_args()
    script = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "train_elastic.py")
    impl = a.impl or ("b200" if a.sched == "local_cuda" else "nccl")
    role = "train_elastic"
    base = ["-j", f"1x{a.nproc}", "--script", script]
    sargs = ["--model", a.model, "--impl", impl, "--steps", str(a.steps), "--batch", str(a.batch), "--fail-at-step", str(a.fail_at_step)]
    cfg = {"log_dir": os.path.join(a.log_dir, a.sched)}
    rows, events = [], {}
    with get_runner() as runner:
        t_submit = time.time()
        if a.sched == "local_cuda":
            h = runner.run_component("dist.ddp", base + ["--max_retries", "1", "--"] + sargs, a.sched, cfg)
            st, _ = wait_terminal(runner, h)
            rows = collect(runner, h, role, 1)
            events["final_state"] = str(st.state)
        else:
            h1 = runner.run_component("dist.ddp", base + ["--"] + sargs, a.sched, cfg)
            st1, t_failed = wait_terminal(runner, h1)
            events["first_state"] = str(st1.state)
            events["t_failed_seen"] = t_failed
            h2 = runner.run_component("dist.ddp", base + ["--"] + sargs + ["--no-fail"], a.sched, cfg)
            events["t_resubmitted"] = time.time()
            st2, _ = wait_terminal(runner, h2)
            events["final_state"] = str(st2.state)
            rows = collect(runner, h1, role, 1)
            for r in collect(runner, h2, role, 1):
                r["attempt"] = 1  # the re-submission plays the role of the second attempt
                rows.append(r)
    kills = [r for r in rows if r.get("event") == "kill"]
    steps1 = [r for r in rows if r.get("event") == "step" and r.get("attempt") == 1]
    steps0 = [r for r in rows if r.get("event") == "step" and r.get("attempt") == 0]
    out = {"sched": a.sched, "impl": impl, "model": a.model, "nproc": a.nproc, "batch_per_gpu": a.batch, "events": events,
           "t_submit_to_first_step_s": None, "kill_to_first_step_after_recovery_s": None, "steady_state_seq_per_s": None}
    if steps0:
        out["t_submit_to_first_step_s"] = round(min(r["t"] for r in steps0 if r["step"] == 0) - t_submit, 3)
        per = {}
        for r in steps0:
            per.setdefault(r["rank"], []).append((r["step"], r["t"]))
        dts = []
        for lst in per.values():
            lst.sort()
            dts += [t1 - t0 for (s0, t0), (s1, t1) in zip(lst, lst[1:]) if s0 >= 10]
        if dts:
            out["steady_state_seq_per_s_before_kill"] = round(a.batch * a.nproc / statistics.median(dts), 1)
    if kills and steps1:
        t_kill = kills[0]["t"]
        first = {}
        for r in steps1:
            if r["step"] == 0:
                first[r["rank"]] = r["t"]
        if len(first) == a.nproc:
            out["kill_to_first_step_after_recovery_s"] = round(max(first.values()) - t_kill, 3)
        starts = [r["t"] for r in rows if r.get("event") == "start" and r.get("attempt") == 1]
        comm = [r["t"] for r in rows if r.get("event") == "comm_ready" and r.get("attempt") == 1]
        model = [r["t"] for r in rows if r.get("event") == "model_ready" and r.get("attempt") == 1]
        out["breakdown_s"] = {
            "kill_to_workers_restarted": round(max(starts) - t_kill, 3) if starts else None,
            "to_comm_ready": round(max(comm) - t_kill, 3) if comm else None,
            "to_model_ready": round(max(model) - t_kill, 3) if model else None,
        }
        per = {}
        for r in steps1:
            per.setdefault(r["rank"], []).append((r["step"], r["t"]))
        dts = []
        for lst in per.values():
            lst.sort()
            dts += [t1 - t0 for (s0, t0), (s1, t1) in zip(lst, lst[1:]) if s0 >= 10]
        if dts:
            out["steady_state_seq_per_s"] = round(a.batch * a.nproc / statistics.median(dts), 1)
            out["steady_state_ms_per_step"] = round(statistics.median(dts) * 1e3, 2)
    print(json.dumps(out), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
