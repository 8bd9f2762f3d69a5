"""The bench.py JSON contract (CPU): the reference arm runs here and prints every required key; the last GPU line committed
under profiles/ carries the full set (roofline, cpu_baseline, e2e, clocks, launches)."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline"}


def _last_json_line(text):
    lines = [l for l in text.strip().splitlines() if l.startswith("{")]
    assert lines, text[-400:]
    return json.loads(lines[-1])


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-800:]
    d = _last_json_line(out.stdout)
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["unit"] == "rays/s" and d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_committed_gpu_lines_carry_the_full_contract():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "bench_r01[g-z]_n1*.json")) + glob.glob(os.path.join(ROOT, "profiles", "bench_r02_final_n*.json")) +
                   glob.glob(os.path.join(ROOT, "profiles", "bench_r02[cd]_n1.json")))
    assert files, "no bench line of the final state under profiles/"
    for f in files:
        d = _last_json_line(open(f).read())
        assert BASE_KEYS | {"gpu_launches", "roofline", "clocks"} <= set(d), (f, sorted(BASE_KEYS - set(d)))
        assert d["metric"].startswith("train rays/s") and d["unit"] == "rays/s" and d["scaling"] == "weak" and d["data"] == "synthetic"
        assert d["vs_baseline"] is None and d["gpu_launches"] > 0 and d["warmup"] >= 3
        assert abs(d["value"] - 8192 * d["n_gpus"] / (d["ms_per_step"] * 1e-3)) < 1e-3 * d["value"]
        r = d["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] in ("hbm", "tensor")
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        e = d["e2e"]
        assert e["unit"] == "rays/s" and e["h2d_bytes_per_step"] == 4 * 8192 * 3 * 4 and e["d2h_bytes_per_step"] > 0
        assert e["value"] != d["value"]
        c = d["cpu_baseline"]
        assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["value"] > 0
        k = d["clocks"]
        assert k["sm_mhz"] and k["sm_max_mhz"] and not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        assert "workload" in d["config"] and "l2" in d["config"]
        if "bench_r02" in f:      # round 2: parity inside the bench, the render and S512 sub-benchmarks with their own objects
            assert d["parity"]["ok"] is True and all(d["parity"]["integers_bit_exact"].values())
            assert d["n_gpus"] == 1 or d["parity"]["replicas_identical"] is True
            r = d["render"]
            assert r["frames"] == 200 and "dense" in r["workload"] and {"roofline", "value", "e2e_fps"} <= set(r)
            assert r["roofline"]["march_steps_per_frame"] > 1e8 and abs(r["roofline"]["frac"] - r["roofline"]["achieved"] / r["roofline"]["peak"]) < 1e-9
            assert d["stress_s512"]["train"]["value"] > 0 and d["stress_s512"]["render"]["value"] > 0
            if d["n_gpus"] == 1:
                assert r["parity"]["ok"] is True and r["parity"]["per_pixel_sample_counts_bit_exact"] is True and r["cpu_baseline"]["value"] > 0
                assert c["kind"] == "reference"
