"""bench.py's JSON contract, checked on the CPU: the reference arm (`--impl reference`, the restated CPU path) at a tiny
degree prints one line with every key the driver reads, and the product arm refuses to run without a CUDA device (there is
no CPU fallback to fall into silently)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_reference_arm_line_has_the_contract_keys():
    r = run("--impl", "reference", "--k", "10", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly ONE JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["unit"] == "s" and d["higher_is_better"] is False and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["value"] > 0 and abs(d["ms_per_step"] - 1000 * d["value"]) < 1e-6 * d["ms_per_step"] + 1e-9
    assert "wall-sec" in d["metric"] and d["config"]["workload"] and d["config"]["k"] == 10
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["unit"] == "s" and cb["value"] == d["value"] and cb["sample"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == "s" and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_product_arm_fails_loudly_without_a_gpu():
    try:
        import torch

        if torch.cuda.is_available():
            return  # on a GPU box the product arm runs; nothing to check here
    except Exception:
        pass
    r = run("--steps", "1", "--warmup", "0", "--k", "10", "--no-cpu-baseline")
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout) and not any(l.startswith("{") for l in r.stdout.splitlines())


def test_step_shape_matches_the_layer1_replay():
    """The unit list of one step is the degree-24 layer replay of DESIGN.md (d): 28 Lagrange commits + 28 transform chains
    + 7 coefficient commits + 1 quotient inverse transform; 35 MSMs of which 19 witness-like; every unit lands on one rank."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    jobs = b.make_jobs()
    kinds = [j[0] for j in jobs]
    assert kinds.count("lmsm") == 28 and kinds.count("ntt") == 28 and kinds.count("msm") == 7 and kinds.count("icoset") == 1
    msms = [j for j in jobs if j[0] in ("lmsm", "msm")]
    assert len(msms) == 35 and sum(1 for j in msms if j[1] == "w") == 19
    assert "2^24" in b.workload_desc(24) and "wall-sec" in b.metric_name(24)
    for world in (1, 2, 4, 8):
        plan = b.assign_jobs(jobs, world)  # per rank: the jobs themselves
        assert sorted(j for r in plan for j in r) == sorted(jobs)
        loads = [sum(j[2] for j in r) for r in plan]
        assert max(loads) <= 1.34 * max(sum(loads) / world, max(j[2] for j in jobs))
