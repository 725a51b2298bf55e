"""bench.py on the GPU, small: the default configuration's line carries every key of the contract, its first-iteration losses
agree with the UNMODIFIED reference run on the CPU from the same seeds (`loss_check`), and the cuDNN baseline ran beside it."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_default_config_line_and_loss_check():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3"], capture_output=True,
                         text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout[-2000:]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "tile_pairs_per_sec_gen_disc_fwd_bwd_256x256x13" and d["value"] > 0 and d["gpu_launches"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["value"] > 0 and abs(d["e2e"]["value"] - d["value"]) > 0
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["batch"] == 16
    assert d["loss_check"] is not None and d["loss_check"]["ok"], d["loss_check"]
    gb = d["gpu_baseline"]
    assert "tf32" in gb and "fp32" in gb, gb
    for k in ("tf32", "fp32"):      # the reference on cuDNN starts from the same losses too (TF32: within its own rounding)
        for a, b in zip(gb[k]["first_losses"], d["first_losses"]):
            assert abs(a - b) <= (5e-3 if k == "tf32" else 1e-3) * max(abs(b), 1e-3), (k, gb[k]["first_losses"], d["first_losses"])
