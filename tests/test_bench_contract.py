"""bench.py's reference arm (`--impl reference`: the CPU restatement of the reference step on the host cores) prints ONE JSON
line with the keys the driver reads.  CPU only; one warm-up and one timed step of the bounded sample (~20 s)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("train EMG frames/s") and d["unit"] == "frames/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "cfg2" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["unit"] == "frames/s" and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


import pytest  # noqa: E402


@pytest.mark.gpu
def test_gpu_arm_prints_the_contract_line():
    """The measured arm on one B200: device-resident value, end-to-end value from host buffers, launch count, clocks, the
    dominant kernel's roofline entry.  Two timed steps after three warm-ups; the CPU baseline leg is skipped here."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3", "--no-cpu-baseline"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["metric"].startswith("train EMG frames/s") and d["unit"] == "frames/s" and d["n_gpus"] == 1
    assert d["steps"] == 2 and d["warmup"] == 3 and d["scaling"] == "weak" and d["dtype"] == "bf16" and d["vs_baseline"] is None
    assert d["value"] > 5e5 and abs(d["value"] - 64000 / (d["ms_per_step"] / 1e3)) < 1e-3 * d["value"]
    e = d["e2e"]
    assert e["value"] > 5e5 and e["h2d_bytes_per_step"] > 1e7 and e["d2h_bytes_per_step"] == 12
    assert d["gpu_launches"] > 100
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    r = d["roofline"]
    assert r["bound"] in ("tensor", "hbm") and 0.0 < r["frac"] <= 1.0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert "workload" in d["config"] and "cfg2" in d["config"]["workload"]
    # the other BASELINE.json configurations ride along in `also`, measured after the headline
    also = d["also"]
    for k in ("cfg3_hybrid", "cfg4_accumulation", "cfg5_inference", "greedy_search_latency"):
        assert k in also and "error" not in also[k], (k, also.get(k))
    assert also["cfg3_hybrid"]["frames_per_s"] > 3e5 and also["cfg4_accumulation"]["optimizer_steps"] == 3
    assert also["cfg5_inference"]["utterances"] == 1000 and also["cfg5_inference"]["utterances_per_s"] > 100
