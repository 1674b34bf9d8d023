"""bench.py's output contract (the driver parses this line): keys, units and the objects the task statement asks for.
The GPU test runs a miniature job through the real code path; the CPU test covers argument handling."""

import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_presets_and_flags_parse():
    sys.path.insert(0, ROOT)
    import bench

    argv = sys.argv
    try:
        sys.argv = ["bench.py", "--preset", "config4"]
        a = bench.parse_args()
        assert (a.games, a.sims, a.width) == (131072, 1600, 64)
        sys.argv = ["bench.py"]
        a = bench.parse_args()
        assert (a.gpus, a.games, a.sims, a.scaling, a.impl) == (1, 16384, 600, "weak", "ours") and a.warmup >= 3
        sys.argv = ["bench.py", "--scaling", "strong", "--total-games", "4096", "--gpus", "8"]
        a = bench.parse_args()
        assert a.scaling == "strong" and a.total_games == 4096
    finally:
        sys.argv = argv
    cfg = bench.workload_config(a, 8)
    assert cfg["global_games"] == a.games * 8 and "lockstep games" in cfg["workload"]


@pytest.mark.gpu
def test_bench_line_has_the_contract_keys():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--games", "512", "--sims", "48", "--steps", "1",
                          "--warmup", "3", "--cpu-seconds", "1", "--no-ablation"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in line, k
    assert line["metric"] == "self-play positions/sec" and line["unit"] == "positions/s" and line["higher_is_better"] is True
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 3 and line["vs_baseline"] is None
    assert line["value"] > 0 and line["gpu_launches"] > 0 and "workload" in line["config"]
    e2e = line["e2e"]
    assert e2e["value"] > 0 and e2e["unit"] == "positions/s" and e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0
    assert e2e["value"] <= line["value"] * 1.001  # the end-to-end number includes the device-timed search
    rf = line["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and rf["peak"] > 0
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["unit"] == "positions/s" and cb["value"] > 0 and cb["sample"]
    assert line["nn_roofline"]["bound"] == "tensor"
