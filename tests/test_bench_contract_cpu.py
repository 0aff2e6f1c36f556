"""CPU: the JSON-line contract of bench.py. The reference arm runs here (it is the reference's CPU path on the host cores); the GPU
arm's line is checked on the committed lines of the round under profiles/ (no GPU in the build container)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
             "data", "config", "e2e", "cpu_baseline"}


def _last_json_line(text):
    lines = [l for l in text.strip().splitlines() if l.startswith("{")]
    assert lines, text[-2000:]
    return json.loads(lines[-1])


@pytest.mark.parametrize("extra", [[], ["--workload", "mutag"]])
def test_reference_arm_line(extra):
    """`bench.py --impl reference`: same metric / config keys as the GPU arm, impl = reference, K timed steps, e2e = value with no
    transfers, cpu_baseline describing the run. Under a multi-process launch only rank 0 works."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"] + extra,
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = _last_json_line(out.stdout)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["unit"] == "graphs/s" and d["value"] > 0 and "workload" in d["config"] and "model" not in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"]
    assert abs(d["ms_per_step"] * d["value"] / 1e3 - 50.0) < 1e-6            # a step = a bounded sample of 50 single-graph steps
    if not extra and cb["kind"] == "reference":                              # molecule workload: the reference's own preprocessing next to it
        assert cb["apsp_preprocessing"]["value"] > 0 and 0 < cb["value_incl_apsp"] < cb["value"]
    other = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                           text=True, timeout=120, env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), cwd=ROOT)
    assert other.returncode == 0 and other.stdout.strip() == ""


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_committed_gpu_lines_follow_the_contract(n):
    path = os.path.join(ROOT, "profiles", f"r02_bench_n{n}.json")
    d = _last_json_line(open(path).read())
    assert BASE_KEYS - {"cpu_baseline"} <= set(d) and d["n_gpus"] == n and d["steps"] >= 1 and d["warmup"] >= 3
    assert d["metric"].startswith("GNAN fwd+bwd") and d["unit"] == "graphs/s" and d["data"] == "synthetic" and d["scaling"] == "weak"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["gpu_launches"] > 0
    assert "workload" in d["config"] and "timing" in d["config"]
    assert abs(d["value"] * d["ms_per_step"] / 1e3 / (32768 * n) - 1.0) < 1e-6          # whole-job graphs/s over all ranks
    e = d["e2e"]
    assert e["unit"] == d["unit"] and 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["clocks"]
    assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"] and not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"])
    if n == 1:
        cb = d["cpu_baseline"]
        assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] > 0
        assert {"cora", "mutag", "pubmed", "arxiv"} <= set(d["sub_records"])
        for name, sub in d["sub_records"].items():
            assert sub["value"] > 0 and sub["e2e"]["value"] > 0 and sub["roofline"]["frac"] > 0, name
    else:
        p = d["parity_vs_single_gpu"]
        assert "error" not in p
        assert d["sub_records"]["arxiv"]["scaling"] == "strong" and "collectives_ms_per_step" in d["sub_records"]["arxiv"]
