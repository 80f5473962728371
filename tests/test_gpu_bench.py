"""bench.py on the GPU at a reduced frame count: the one JSON line must carry every key of the contract."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_contract():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1", "--steps", "3", "--warmup", "3",
                          "--frames", "4096", "--e2e-frames", "512", "--e2e-steps", "1", "--direct-frames", "256",
                          "--side-frames", "2048", "--cpu-frames", "50"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                         timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] >= 3 and d["scaling"] == "weak"
    assert d["unit"] == "evals/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0 < r["frac"] < 1.2
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["gpu_launches"] == 2 * d["steps"]                       # fused kernel + finalize per step
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["spot_check"]["max_angle_err_deg"] <= 1e-3
    assert d["value"] > 100 * c["value"]
    assert d["reference_native_shape"]["bit_exact"] is True
    assert d["full_size_check"]["equal"] is True and d["full_size_check"]["score_sum"] > 0
    assert d["full_size_check"]["occupancy_equals_flag_counts_per_feature"] is True
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    # staging is inside the number: the timed input is raw frames in plan order, and the caller-order / other-config legs exist
    assert "no staging kernel" in d["config"]["input"] and d["caller_order_input"]["k0_ms"] > 0
    assert set(d["all_configs"]) == {"cfg1", "cfg2", "cfg4_distances_emitted", "cfg4_distance_matrix_only", "cfg4_contact_map", "cfg5"}
    for name, c in d["all_configs"].items():
        assert c["value"] > 0 and 0 < c["roofline_frac_per_gpu"] < 1.2, name
    assert d["config5_sharded"]["frames_total"] == 500000 and d["config5_sharded"]["occupancy_sum"] > 0
