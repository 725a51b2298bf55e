"""2-GPU equivalence of the data-parallel layer on real NCCL (SURVEY.md §4 last row, §8(e)): an N-rank sharded iteration ==
the 1-rank full-batch iteration, for the eager exchange and for the CUDA-graph form (tests/_dp_equiv.py).  Needs >= 2 GPUs
(`gpurun --gpus 2`); skipped on a 1-GPU box.  The host-side logic of the same code runs on gloo in tests/test_parallel.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bounded(cmd, env, seconds):
    """Run `cmd` in its own process group with a hard time limit; a hang returns what was printed so far (rc 124)."""
    import signal
    import types

    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, cwd=ROOT, start_new_session=True)
    try:
        out, err = p.communicate(timeout=seconds)
        return types.SimpleNamespace(returncode=p.returncode, stdout=out, stderr=err)
    except subprocess.TimeoutExpired:
        os.killpg(p.pid, signal.SIGKILL)
        out, err = p.communicate()
        return types.SimpleNamespace(returncode=124, stdout=out, stderr=err + f"\n[killed after {seconds} s]")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", ["eval_bn", "sync_bn"])
def test_sharded_step_equals_full_batch_step(mode):
    """eval_bn: BatchNorm on running statistics (samples independent).  sync_bn: train-mode BatchNorm with
    `set_sync_bn(True)` — the N-rank run reproduces the single-process run on the concatenated batch, running statistics
    included ('SyncBN when global-batch parity with a single-process reference run is required')."""
    env = dict(os.environ, FCD_DIST_TIMEOUT_S="120")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "_dp_equiv.py")] +
                         (["--sync-bn"] if mode == "sync_bn" else []),
                         capture_output=True, text=True, timeout=420, env=env, cwd=ROOT) if mode == "eval_bn" else _bounded(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
         "--master-port", "29534", os.path.join(ROOT, "tests", "_dp_equiv.py"), "--sync-bn"], env, 150)
    if res.returncode != 0 or "DP_EQUIV_OK" not in res.stdout:           # keep the whole transcript where gpurun brings it back
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"dp_equiv_failure_{mode}.log"), "w") as fh:
            fh.write(res.stdout + "\n---- stderr ----\n" + res.stderr)
    assert res.returncode == 0 and "DP_EQUIV_OK" in res.stdout, (res.stdout[-1500:], res.stderr[-3000:])
    print(res.stdout.strip().splitlines()[-1])
