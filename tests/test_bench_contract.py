"""bench.py contract on the CPU: the reference arm (`--impl reference`, the oracle port on the host cores) prints exactly
one JSON line with the keys the driver reads, on the same metric / workload strings as the GPU arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    # (--layers 2: the real arm fills the full 13.5 GB model; the contract is the same.)  torchrun exports OMP_NUM_THREADS=1
    # to every rank: the arm must use every core anyway (the N > 1 reference numbers ran on one core)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ctx", "64",
                        "--layers", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "tokens/s" and j["higher_is_better"] is True and j["value"] > 0
    assert j["metric"] == json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
    assert j["config"]["workload"].startswith("LLaMA-7B f16, 1-token decode, ctx=64")
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == j["value"] and "layers" in cb["sample"] and cb["steps"] >= 3
    assert cb["cores"] == len(os.sched_getaffinity(0)), cb
    assert j["invalid"] and set(j["config"]) == {"workload", "n_layer", "arithmetic", "l2", "kv", "parallelism"}
    assert j["e2e"] == {"value": j["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_non_zero_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_keys_are_in_the_source():
    """the GPU arm cannot run here; at least every contract key appears in the line it builds"""
    src = open(os.path.join(ROOT, "bench.py")).read()
    for key in ('"metric"', '"value"', '"unit"', '"n_gpus"', '"steps"', '"warmup"', '"ms_per_step"', '"higher_is_better"', '"scaling"',
                '"vs_baseline"', '"dtype"', '"data"', '"config"', '"clocks"', '"e2e"', '"gpu_launches"', '"roofline"', '"cpu_baseline"',
                '"h2d_bytes_per_step"', '"d2h_bytes_per_step"', '"bound"', '"achieved"', '"peak"', '"frac"', '"traffic"', '"parity_check"',
                '"rel_err"', '"greedy_equal"', '"allreduce"', '"extra"'):
        assert key in src, key
