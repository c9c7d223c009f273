"""`bench.py` contract on the CPU: the reference arm (`--impl reference`, the only arm that runs without a GPU) prints ONE JSON
line with the driver's keys, and both arms describe the workload with the same `config` object."""
import json
import os
import subprocess
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

LINE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "cpu_baseline", "e2e")


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--shape", "tiny",
                        "--steps", "1", "--warmup", "1", "--pool", "2", "--pairs", "4"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "exactly one line on stdout"
    d = json.loads(lines[0])
    for k in LINE_KEYS:
        assert k in d, k
    assert d["impl"] == "reference" and d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_both_arms_share_the_config_object():
    sys.path.insert(0, ROOT)
    import bench
    a = types.SimpleNamespace(shape="davis", pairs=32, pool=8)
    assert bench.workload_config(a, 1) == bench.workload_config(a, 1)
    assert bench.workload_config(a, 8)["global_batch"] == 256 and bench.workload_config(a, 8)["parallelism"] == "dp8"


def test_kiba_subrun_parses_the_child_line(monkeypatch):
    sys.path.insert(0, ROOT)
    import bench
    line = open(os.path.join(ROOT, "profiles", "r2_bench_kiba_n1.json")).read().strip()
    fake = types.SimpleNamespace(returncode=0, stdout="NCCL version 2.28.9+cuda12.9\n" + line + "\n", stderr="")
    monkeypatch.setattr(bench.subprocess, "run", lambda *a, **k: fake)
    r = bench.kiba_shape_subrun(types.SimpleNamespace(steps=20, warmup=5, pool=8))
    assert r["unit"] == "pairs/s" and r["value"] > 0 and "kiba-shape, 64 pairs/GPU" in r["config"]["workload"]
    broken = types.SimpleNamespace(returncode=1, stdout="", stderr="boom")
    monkeypatch.setattr(bench.subprocess, "run", lambda *a, **k: broken)
    assert "error" in bench.kiba_shape_subrun(types.SimpleNamespace(steps=20, warmup=5, pool=8))
