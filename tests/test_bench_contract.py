"""The JSON line of bench.py against the driver's contract (keys, types, consistency), checked on the lines
committed under profiles/ - so a change of bench.py that drops a key shows up on the CPU box - plus the
argument parser and the workload table of bench.py itself (no GPU needed)."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NATIVE = sorted(glob.glob(os.path.join(ROOT, "profiles", "bench_r01[fghi]_*.json")))
REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _line(path):
    with open(path) as fh:
        return json.loads(fh.read().strip().splitlines()[-1])


@pytest.mark.parametrize("path", NATIVE, ids=[os.path.basename(p) for p in NATIVE])
def test_committed_bench_lines_follow_the_contract(path):
    d = _line(path)
    if d.get("impl") == "reference":
        assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
        assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]
        return
    assert REQUIRED <= set(d), REQUIRED - set(d)
    assert d["metric"] == "audio_seconds_per_second" and d["unit"] == "audio-s/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f32"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    # value = audio of all ranks / device time of one step
    audio = d["config"]["audio_seconds_per_gpu_step"] * d["n_gpus"]
    assert d["value"] == pytest.approx(audio / (d["ms_per_step"] * 1e-3), rel=1e-9)
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=2e-3)
    assert r["kernel"] == max(d["kernels"], key=lambda k: k["ms"])["kernel"]
    for k in d["kernels"]:
        assert k["bound"] in ("hbm", "tensor") and k["unit"] in ("GB/s", "TFLOP/s") and k["ms"] > 0
    if d["n_gpus"] == 1 and d.get("cpu_baseline"):
        c = d["cpu_baseline"]
        assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    if d.get("clocks"):
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    p = d["parity"]
    assert p["mfcc_max_abs"] < 1e-3 and p["autocorr_max_abs"] < 2e-5


def test_weak_scaling_lines_are_consistent():
    lines = {n: _line(os.path.join(ROOT, "profiles", f"bench_r01g_n{n}.json")) for n in (1, 2, 4, 8)}
    for n, d in lines.items():
        assert d["n_gpus"] == n and d["config"]["clips_per_gpu"] == 60
        assert d["value"] > 0.95 * n * lines[1]["value"]                 # near-linear by clip, no collective


def test_bench_workload_table_and_algorithmic_counts():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert set(bench.WORKLOADS) == {"c2", "c3", "c4", "c5"}
    assert {n for n, w in bench.WORKLOADS.items() if w["collect"]} == {"c3", "c4"}
    # configs[2] and [4] are split over the ranks (strong), c2 / c4 carry a fixed batch per rank
    assert bench.DEFAULT_SCALING == {"c2": "weak", "c3": "strong", "c4": "weak", "c5": "strong"}
    # the math libraries are pinned before NumPy is imported by the module (children of the CPU pool inherit it)
    assert all(os.environ.get(k) == "1" for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"))
    alg = bench.algorithmic(88200, 1470, 735, 384)
    # per hop-frame at 88.2 kHz
    assert alg["stft_gemm"][1] == 4327680 and alg["autocorr"][1] == 517564
    assert alg["stft_gemm"][0] == alg["autocorr"][0] == "tensor" and alg["fold"][0] == "hbm"
