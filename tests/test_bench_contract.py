"""bench.py's output contract on the CPU: the reference arm (`--impl reference`) prints exactly ONE JSON line on stdout with the
keys the driver reads, whatever libraries write elsewhere; non-zero ranks of a torchrun launch print nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    res = _run()
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "tile_pairs_per_sec_gen_disc_fwd_bwd_256x256x13"
    assert d["unit"] == "tile-pairs/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["steps"] == 1
    cb = d["cpu_baseline"]
    # the reference arm runs the UNMODIFIED reference classes (oracle/_ref, staged by __graft_entry__.build()) at the workload's own batch
    assert cb["kind"] == "reference" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and "batch 16" in cb["sample"]
    assert len(cb["first_losses"]) == 2
    assert d["e2e"] == {"value": d["value"], "unit": "tile-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_every_config_is_declared_with_its_algorithmic_work():
    """`--config` covers the north-star configuration and BASELINE.json configs[1..4] with SURVEY.md §8(d)'s GFLOP figures."""
    sys.path.insert(0, ROOT)
    import bench

    assert {k: v["gf"] for k, v in bench.CONFIGS.items()} == {"2": 227.3, "g32": 203.6, "3": 3961.0, "4": 889.0, "5": 1654.0}
    assert bench.CONFIGS["g32"]["B"] == 32 and bench.CONFIGS["4"]["B"] == 16 and bench.CONFIGS["5"]["B"] == 32
    for k in bench.CONFIGS:
        ts = bench.synth_for(k, 1, 0)
        assert len(ts) == {"2": 4, "g32": 2, "3": 2, "4": 3, "5": 4}[k]
        assert ts[0].shape == (1, bench.CONFIGS[k]["C"], bench.CONFIGS[k]["H"], bench.CONFIGS[k]["W"])


def test_reference_arm_other_ranks_stay_silent():
    res = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert res.returncode == 0 and res.stdout.strip() == ""
