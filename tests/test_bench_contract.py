"""CPU: the reference arm of bench.py (`--impl reference`: the oracle port of the reference's CPU path on the host cores) runs
without a GPU and prints exactly ONE JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-rows", "400000"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "rows/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("rows/sec windowed group-agg") and d["config"]["workload"].startswith("cfg2")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1


def test_non_zero_ranks_of_the_reference_arm_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == "", (out.stdout, out.stderr[-1000:])


def test_e2e_step_plan_fits_the_device():
    """bench.py runs every e2e step through its own operator (~9 GB each for cfg 2): the number of steps follows the free device
    memory, with the driver's --steps 20 --warmup 5 as the case that used to need 26 operators."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
    assert b.plan_e2e_steps(11, 5, 20, 1) == (3, 7)          # 11 operators fit: 3 warm-ups + 7 timed + the parity pass
    assert b.plan_e2e_steps(40, 5, 20, 1) == (3, 20)
    assert b.plan_e2e_steps(9, 3, 5, 1) == (3, 5)
    assert b.plan_e2e_steps(3, 5, 20, 1) == (1, 1)
    for fit in range(3, 30):
        for par in (0, 1):
            w, s = b.plan_e2e_steps(fit, 5, 20, par)
            assert w >= 1 and s >= 1 and w + s + par <= max(fit, 2 + par)
