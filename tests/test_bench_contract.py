"""CPU test of the bench.py contract on the reference arm (the only arm that runs without a GPU): one JSON line
with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

from oracle import refplumed as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.ref
def test_reference_arm_json_line():
    if not R.available():
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3",
                          "--warmup", "3", "--ref-atoms", "3000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["steps"] == 3 and "workload" in d["config"]
    assert d["config"]["natoms"] == 3000 and "3000 atoms" in d["config"]["workload"]  # the arm states ITS size


def test_non_root_ranks_of_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "3",
                          "--warmup", "3"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_clock_sampler_counts_only_samples_inside_the_timed_regions(tmp_path, monkeypatch):
    """bench.py's nvidia-smi sampler: samples are stamped on receipt, the report covers the timed regions, and a
    region shorter than the sampling period falls back to everything sampled (and says so)"""
    import time
    fake = tmp_path / "nvidia-smi"
    fake.write_text("#!/bin/bash\nwhile true; do echo '1965, 1965, 400.5, Not Active, Not Active, Not Active, Active'; sleep 0.02; done\n")
    fake.chmod(0o755)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0)
    s.start()
    s.wait_first()
    assert s.rows, "the sampler must have delivered before a timed region opens"
    time.sleep(0.1)
    before = len(s.rows)
    s.open_window()
    time.sleep(0.15)
    s.close_window()
    time.sleep(0.1)
    r = s.stop()
    assert r["window"].startswith("timed regions") and 1 <= r["samples"] < len(s.rows) - before + 3
    assert r["samples"] < len(s.rows) and r["sm_mhz"] == 1965.0 and r["reasons"] == ["sw_power_cap"]
    s = bench.ClockSampler(0)
    s.start()
    s.wait_first()
    s.open_window()
    s.close_window()
    r = s.stop()
    assert r["window"].startswith("whole run") and r["samples"] >= 1
