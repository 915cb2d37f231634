"""CPU: the bench lines committed under profiles/ carry every key of the measurement contract, and the
reference arm of the input-pipeline workload runs here end to end (it is CPU-only by definition)."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"}


def final_lines():
    return sorted(glob.glob(os.path.join(ROOT, "profiles", "r0[12]_bench_*_final.json")))


@pytest.mark.parametrize("path", final_lines(), ids=os.path.basename)
def test_committed_bench_lines_follow_the_contract(path):
    line = json.loads(open(path).read().strip().splitlines()[-1])
    assert BASE_KEYS <= set(line), f"missing {BASE_KEYS - set(line)}"
    assert line["vs_baseline"] is None and line["data"] == "synthetic" and "workload" in line["config"]
    assert "model" not in line["config"]
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["value"] != line["value"]
    roof = line["roofline"]
    assert roof["bound"] in ("hbm", "tensor") and roof["unit"] in ("GB/s", "TFLOP/s")
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-6 and 0 < roof["frac"] < 1
    assert line["gpu_launches"] > 0 and line["warmup"] >= 3
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if line["n_gpus"] == 1 and "cpu_baseline" in line:
        assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
        assert line["cpu_baseline"]["kind"] in ("port", "reference")


def test_there_is_a_final_line_for_the_headline_workload():
    names = [os.path.basename(p) for p in final_lines()]
    assert "r01_bench_grounding_final.json" in names


def test_round2_line_carries_every_baseline_config():
    """Round 2: ONE line per default run -- the grounding headline with decode / train / icl under `secondary`, each
    with its own e2e, clocks and roofline; the dominant-kernel roofline entry is the decode kernel (HBM bound)."""
    path = os.path.join(ROOT, "profiles", "r02_bench_all_final.json")
    if not os.path.exists(path):
        pytest.skip("no round-2 final line committed yet")
    line = json.loads(open(path).read().strip().splitlines()[-1])
    assert BASE_KEYS <= set(line) and line["unit"] == "images/s"
    assert line["roofline"]["bound"] == "hbm" and "llama_decode_kernel" in line["roofline"]["kernel"]
    assert line["roofline_gemm"]["bound"] == "tensor"
    assert set(line["secondary"]) == {"decode", "train", "icl"}
    for name, unit in (("decode", "tokens/s"), ("train", "samples/s"), ("icl", "images/s")):
        sec = line["secondary"][name]
        assert "error" not in sec, sec
        assert sec["unit"] == unit and sec["value"] > 0 and sec["e2e"]["h2d_bytes_per_step"] > 0
        assert 0 < sec["roofline"]["frac"] < 1.2 and sec["clocks"]["sm_mhz"]
    assert line["cpu_baseline"]["kind"] == "port" and "FULL depth" in line["cpu_baseline"]["sample"]


def test_reference_arm_of_the_input_pipeline_runs_on_cpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                          "preprocess", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must be exactly one JSON line"
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
