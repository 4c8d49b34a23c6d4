"""The headline numbers quoted in DESIGN.md / README.md are the ones in the committed measurement files under profiles/
(so the prose cannot drift away from the evidence)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _doc(name):
    return open(os.path.join(ROOT, name)).read()


def bench(name):
    lines = [l for l in open(os.path.join(ROOT, "profiles", name)).read().splitlines() if l.startswith("{")]
    return json.loads(lines[-1])


def jsonl(name):
    return [json.loads(l) for l in open(os.path.join(ROOT, "profiles", name)) if l.startswith("{")]


def test_bench_lines_quoted_in_design():
    DESIGN, README = _doc("DESIGN.md"), _doc("README.md")
    for name, digits in (("bench_r01_v12.json", 3), ("bench_r01_v14.json", 3), ("bench_r01_2gpu_v2.json", 3), ("bench_r01_8gpu_v6.json", 3),
                         ("bench_r01_k25.json", 3), ("bench_r01_k26.json", 3)):
        d = bench(name)
        assert name in DESIGN
        assert f"{d['value']:.{digits}f}" in DESIGN, (name, d["value"])
        assert f"{d['e2e']['value']:.{digits}f}" in DESIGN, (name, d["e2e"]["value"])
    v14 = bench("bench_r01_v14.json")
    assert v14["clocks"]["reasons"] == [] and v14["clocks"]["sm_mhz"] == 1965 and "1965 MHz" in DESIGN
    assert v14["gpu_launches"] > 0 and v14["roofline"]["frac"] > 0.85
    assert f"{v14['value']:.2f}" in README


def test_quotient_and_sweep_numbers_quoted_in_design():
    DESIGN = _doc("DESIGN.md")
    k24 = {r["op"]: r for r in jsonl("quotient_time_r01_k24.jsonl")}
    assert f"{k24['graph_evaluate']['ms_best']:.1f} ms" in DESIGN
    assert f"{k24['permutation_product']['ms_best']:.1f} ms" in DESIGN
    assert f"{k24['logup_running_sum']['ms_best']:.1f} ms" in DESIGN
    sweep = jsonl("sweep_r01.jsonl")
    msm24 = next(r for r in sweep if r["op"] == "msm" and r["log_n"] == 24)
    ntt26 = next(r for r in sweep if r["op"] == "ntt" and r["log_n"] == 26)
    assert f"{msm24['c']} / {msm24['W']}, {msm24['ms_best']:.1f}" in DESIGN
    assert f"{ntt26['ms_best']:.1f} ({ntt26['Gbutterflies_s']:.1f} G)" in DESIGN


def test_sanitizer_logs_are_clean():
    mem = open(os.path.join(ROOT, "profiles", "sanitizer_memcheck_r01.log")).read()
    race = open(os.path.join(ROOT, "profiles", "sanitizer_racecheck_r01.log")).read()
    assert "ERROR SUMMARY: 0 errors" in mem and "passed" in mem
    assert "0 hazards displayed (0 errors, 0 warnings)" in race and "passed" in race
