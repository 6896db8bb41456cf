"""bench.py host side (no GPU): the reference arm's JSON line carries the contract's keys, and the op / traffic
bookkeeping of the GPU arm is consistent with BASELINE.json and the committed profiles."""
import importlib.util
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(REPO, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_arm_line_has_the_contract_keys():
    """`bench.py --impl reference` on a tiny bounded sample (400 points): ONE JSON line, the GPU arm's metric /
    unit / config keys, `impl`, a `cpu_baseline` describing this very run and an `e2e` object without copies."""
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-points", "400"], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(REPO, "BASELINE.json")))
    # BASELINE.json's metric: "LM iters/sec & residual+Jacobian obs/sec, 24 cams x 1M pts, ..." -> the headline half
    assert "LM iters/sec" in base["metric"]
    assert d["impl"] == "reference" and d["metric"] == "LM_iters_per_sec" and d["unit"] == "iter/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["n_gpus"] == 1
    for k in ("value", "steps", "warmup", "ms_per_step", "scaling", "data"):
        assert k in d
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["unit"] == d["unit"] and cb["value"] == d["value"]
    assert cb["cores"] > 0 and "bundleAdjust(1e-4)" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-6 * 1e3


def test_configs_follow_baseline_json_and_traffic_stamps_are_wellformed():
    b = _bench()
    base = json.load(open(os.path.join(REPO, "BASELINE.json")))
    assert sorted(b.CONFIGS) == list(range(1, len(base["configs"]) + 1))
    # algorithmic <= executed int8 ops, both grow linearly in the points
    a1, e1 = b.i8_ops(24, 1_000_000)
    a2, e2 = b.i8_ops(24, 2_000_000)
    assert 0 < a1 <= e1 and abs(a2 / a1 - 2.0) < 1e-3 and abs(e2 / e1 - 2.0) < 1e-3
    tr = json.load(open(os.path.join(REPO, "profiles", "r02_traffic.json")))
    for kernel, (fn, sha) in tr["kernels_sha16"].items():
        assert kernel in tr["dram_bytes_per_launch"] and len(sha) == 16
        assert os.path.exists(os.path.join(REPO, "lasercalib_b200", "csrc", fn))
    assert b._sha16(os.path.join(REPO, "bench.py")) == b._sha16(os.path.join(REPO, "bench.py"))
