"""The reference arm of bench.py runs without a GPU: check that it prints ONE JSON line carrying the keys of the bench
contract (metric / unit / config of the product arm, `impl`, `cpu_baseline`, a zero-copy `e2e`)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "stdout must carry exactly one line (logs go to stderr)"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "height_scan_rays_per_s" and d["unit"] == "rays/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f32"
    assert "4096 envs x 961 rays" in d["config"]["workload"]
    assert d["value"] > 1e5 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["unit"] == "rays/s" and cb["value"] == d["value"] and cb["sample"]
    # cfg-1 (BASELINE.json configs[0]): the whole non-physics step at 256 envs on the host cores, >= 20 + >= 30 steps
    c1 = cb["cfg1"]["cpu"]
    assert c1["envs"] == 256 and c1["env_steps_per_s"] > 0 and c1["warmup_steps"] >= 20 and c1["timed_steps"] >= 30
    # both arms print the SAME config dict (the driver compares them)
    sys.path.insert(0, ROOT)
    import bench

    assert d["config"] == bench.bench_config(1)
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == "rays/s"
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_product_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: the product arm of bench.py must fail loudly on a box without a GPU."""
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("needs a box without a GPU")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode != 0 and "no CPU fallback" in (res.stderr + res.stdout)
