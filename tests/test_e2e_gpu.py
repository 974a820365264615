"""End to end on the GPU box: CLI/Runner -> local_cuda scheduler -> one pinned worker per rank -> Communicator.from_env
(shm control block + CUDA IPC) -> mini-DDP training on the fused kernels, including the elastic re-launch of config #5.
With one GPU all ranks share cuda:0 (``devices=0,0``); with more, each rank gets its own device."""
import os
import re

import pytest

from torchx_b200.runner import get_runner
from torchx_b200.specs import AppState

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "examples", "train_ddp.py")


def _devices(world, cuda_count):
    return [str(i) for i in range(world)] if cuda_count >= world else ["0"] * world


def _run(world, tmp_path, cuda_count, extra_component_args=(), script_args=()):
    with get_runner() as runner:
        args = ["-j", f"1x{world}", "--gpu", "1", "--script", SCRIPT, *extra_component_args, "--", "--steps", "8", "--max-ctas", "8", *script_args]
        handle = runner.run_component("dist.ddp", args, "local_cuda", cfg={"log_dir": str(tmp_path / "logs"), "devices": _devices(world, cuda_count)})
        status = runner.wait(handle, wait_interval=0.5)
        lines = list(runner.log_lines(handle, "train_ddp", 0))
        return status, lines


def test_training_job_through_the_scheduler(tmp_path, cuda_count):
    status, lines = _run(2, tmp_path, cuda_count)
    assert status.state == AppState.SUCCEEDED, (status, lines[-20:])
    shas = re.findall(r"params sha256 ([0-9a-f]{16})", "".join(lines))
    assert len(shas) == 2 and len(set(shas)) == 1, lines  # replicas identical after training from DIFFERENT inits
    assert all("attempt 0" in ln for ln in lines if "sha256" in ln)


def test_rank_drop_relaunches_the_gang_under_a_new_epoch(tmp_path, cuda_count):
    status, lines = _run(2, tmp_path, cuda_count, extra_component_args=["--max_retries", "1"], script_args=["--fail-rank", "1", "--fail-at-step", "3"])
    text = "".join(lines)
    assert status.state == AppState.SUCCEEDED and status.num_restarts == 1, (status, lines[-20:])
    assert "injected failure at step 3" in text
    shas = re.findall(r"attempt 1 .*params sha256 ([0-9a-f]{16})", text)
    assert len(shas) == 2 and len(set(shas)) == 1, lines
