"""CPU suite: the driver-facing contract of bench.py that can be checked without a GPU -- the reference arm
(`--impl reference`: the unmodified reference binary on the host cores) prints ONE JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "CoLoRe_ref")


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/CoLoRe_ref not built (needs /root/reference)")
def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-n-grid", "64"], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("Mcells/s") and d["unit"] == "Mcells/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "n_grid=64" in cb["sample"]
    assert {"fill", "fft", "density", "normalization", "sources"} <= set(cb["stages_ms"])


def test_reference_arm_other_ranks_are_silent():
    """under torchrun only rank 0 runs and prints the reference arm; the other ranks exit 0 without work"""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=60, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_roofline_traffic_comes_from_the_newest_committed_capture():
    """bench.py's roofline.traffic = DRAM bytes per launch of the dominant kernel from the committed `ncu --set full`
    summaries of THIS workload (1024^3, one GPU): the newest capture (round, then _vN tag) wins, captures of other grid
    sizes are ignored."""
    sys.path.insert(0, ROOT)
    import bench
    tab, src = bench.ncu_traffic_table()
    assert "_1024_" in src["fill_fft_z"] and "_1024_" in src["fft_yx"]
    assert src["fill_fft_z"] >= "r2_ncu_fft_1024_v6_full.csv"      # that capture or a later one
    # the fused fill + z pass only WRITES its output: two half-spectra of 1024^2 x 520 complex64 = 8.7 GB
    assert 8.6e9 < tab["fill_fft_z"] < 8.8e9
    # one fused y + x pass reads one half-spectrum and writes one real field
    assert 8.6e9 < tab["fft_yx"] < 8.8e9
