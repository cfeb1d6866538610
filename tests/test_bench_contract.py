"""CPU: the bench.py JSON-line contract on the arm that runs without a GPU (`--impl reference`, tiny config), and that the
product arm refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=300,
                          cwd=ROOT)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--model", "tiny", "--batch", "2", "--ctx", "16", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-400:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1  # exactly ONE JSON line
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_nonzero_rank_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "tiny", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_times_the_steps_it_prints_and_loads_no_product_code():
    """the reference arm printed `steps: 20` while timing <= 4 and imported the product package (which maps
    libb200spark.so into the reference process).  Now: `steps` timed steps are really run (steps x ms_per_step ~ the wall
    time of the timed region) and no b200spark module / native library is loaded."""
    code = ("import sys, json, io, contextlib, time; sys.argv = ['bench.py', '--impl', 'reference', '--model', 'tiny', '--batch', '2', "
            "'--ctx', '16', '--steps', '7', '--warmup', '2']; import bench; buf = io.StringIO(); t0 = time.perf_counter()\n"
            "with contextlib.redirect_stdout(buf): bench.main()\n"
            "wall = time.perf_counter() - t0; d = json.loads(buf.getvalue().strip().splitlines()[-1])\n"
            "maps = open('/proc/self/maps').read()\n"
            "print(json.dumps({'steps': d['steps'], 'warmup': d['warmup'], 'timed_s': d['steps'] * d['ms_per_step'] * 1e-3, 'wall': wall,\n"
            "  'b2_modules': [m for m in sys.modules if m.startswith('b200spark')], 'so': 'libb200spark' in maps or 'liballspark_b200' in maps}))")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-600:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["steps"] == 7 and d["warmup"] == 2
    assert d["b2_modules"] == [] and d["so"] is False
    assert d["timed_s"] <= d["wall"]  # the claimed timed region fits inside the run


def test_reference_arm_sets_threads_under_torchrun():
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "tiny", "--batch", "2", "--ctx",
                        "16", "--steps", "1", "--warmup", "1", "--gpus", "2"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-400:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    sys.path.insert(0, ROOT)
    import bench
    assert d["cpu_baseline"]["cores"] == bench.physical_cores()
