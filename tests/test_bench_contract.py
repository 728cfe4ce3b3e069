"""bench.py's reference arm (the CPU path: runs here without a GPU) against the driver's JSON contract."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env, *args):
    env = dict(os.environ)
    env.update(extra_env)
    env.pop("DUCC0_NUM_THREADS", None)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "128", "--steps", "2",
                           "--warmup", "1"] + list(args), capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = _run({"OMP_NUM_THREADS": "1"})                      # what torch.distributed.run exports to every rank
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ls_operator_applies_per_s_2d" and d["unit"] == "applies/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0
    assert abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-6 * 1e3
    assert d["config"]["grid"] == [128, 128] and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] >= 1 and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "applies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_runs_on_rank_zero_only():
    p = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_cpu_arm_keeps_its_threads_under_torchrun():
    """OMP_NUM_THREADS=1 (exported by torch.distributed.run) must not shrink scipy's FFT pool: bench.py sets DUCC0_NUM_THREADS."""
    code = ("import os, sys; sys.argv=['bench.py']; os.environ['OMP_NUM_THREADS']='1'; os.environ.pop('DUCC0_NUM_THREADS', None); "
            "import importlib.util; s=importlib.util.spec_from_file_location('bench', %r); m=importlib.util.module_from_spec(s); "
            "s.loader.exec_module(m); print(os.environ['DUCC0_NUM_THREADS'])" % os.path.join(ROOT, "bench.py"))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    assert int(p.stdout.strip()) == (os.cpu_count() or 1)
