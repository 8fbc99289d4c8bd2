"""bench.py's own arm on the GPU: exactly one JSON line on stdout, carrying every key the measurement contract names."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--periods", "8", "--no-cpu", "--sustain", "0.3", "--shared-carriers", "4"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]                   # diagnostics go to stderr
    d = json.loads(lines[0])
    assert d["metric"].startswith("Msamples/s complex IQ through RECC demod+correlate") and d["unit"] == "Msamples/s"
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["dtype"] == "f32" and "workload" in d["config"]
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and rf["achieved"] > 0 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 8 * d["config"]["samples_per_step_per_gpu"] and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                           # host link in the timed region
    assert d["sc16_input"]["e2e"]["h2d_bytes_per_step"] == 4 * d["config"]["samples_per_step_per_gpu"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    # the parity gate inside the run: the bursts of the host-noise periods were compared with the oracle's, blob for blob
    assert d["parity_checked_bursts"] >= 3 * 3 - 2 and d["bursts_decoded"] >= 3 * 8 - 2
    assert d["config"]["snr_db"] == ["inf", 30, 15]
    assert rf["fp32_tflops"] > 0 and rf["sustained"]["seconds"] >= 0.1 and rf["sustained"]["bursts_equal_to_oracle"] > 0
    assert e["h2d_ceiling"]["value"] >= 0.8 * e["value"]
    sh = d["shared_upload"]
    assert sh["carriers_per_gpu"] == 4 and sh["value"] > 2.0 * d["sc16_input"]["e2e"]["value"]
