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
    for rnd in ("r01", "r02"):  # r02: the batched MSM, the bit-sliced reduction, the grouped pipeline, the sharded entry points, the session
        mem = open(os.path.join(ROOT, "profiles", f"sanitizer_memcheck_{rnd}.log")).read()
        race = open(os.path.join(ROOT, "profiles", f"sanitizer_racecheck_{rnd}.log")).read()
        assert "ERROR SUMMARY: 0 errors" in mem and "passed" in mem and "failed" not in mem
        assert "0 hazards displayed (0 errors, 0 warnings)" in race and "passed" in race and "failed" not in race


def test_round2_numbers_quoted_in_design_and_readme():
    """the round-2 headline (BASELINE configs[2] step) and the multi-GPU / sharded / drop-in figures are the committed ones"""
    DESIGN, README = _doc("DESIGN.md"), _doc("README.md")
    v = bench("bench_r02_v1.json")
    assert "configs[2]" in v["config"]["workload"] and [l["k"] for l in v["config"]["layers"]] == [20, 24, 25]
    assert f"{v['value']:.3f} s" in DESIGN and f"{v['e2e']['value']:.3f} s" in DESIGN
    v2 = bench("bench_r02_v2.json")  # the final library of the round
    assert f"{v2['value']:.3f} s" in DESIGN and f"{v2['e2e']['value']:.3f} s" in DESIGN
    v3 = bench("bench_r02_v3.json")  # the final library, bounded in-line CPU sample
    assert f"{v3['value']:.3f} s" in DESIGN and f"{v3['e2e']['value']:.3f} s" in DESIGN
    v4 = bench("bench_r02_v4.json")  # the last commit of the round
    assert f"{v4['value']:.3f} s" in DESIGN and f"{v4['e2e']['value']:.3f} s" in DESIGN and f"{v4['value']:.2f} s" in README
    assert v4["clocks"]["reasons"] == [] and v4["value"] <= v3["value"] * 1.01
    assert v3["cpu_baseline"]["scaled_layers"] == ["layer2"] and abs(v3["cpu_baseline"]["value"] / v2["cpu_baseline"]["value"] - 1) < 0.05
    assert v2["clocks"]["reasons"] == [] and v2["gpu_launches"] > 0 and 0.85 < v2["roofline"]["frac"] < 1.0
    ref = bench("bench_r02_reference_arm.json")
    assert ref["impl"] == "reference" and ref["extrapolated_by_op_counts"] and ref["full_size_samples"] == len(ref["sample_values_s"]) == 2
    assert f"{ref['value']:.1f} s" in DESIGN and f"{ref['value']:.0f} s" in README
    total = sum(ref["cpu_baseline"]["detail"][n]["layer_s"] for n in ("inner", "layer1", "layer2"))
    assert abs(total - ref["sample_values_s"][-1]) < 1e-6 * total  # detail x op counts reproduces the value
    for name in ("inner", "layer1", "layer2"):
        assert f"{v['layers_s'][name]:.3f}" in DESIGN, name
    assert v["clocks"]["reasons"] == [] and v["clocks"]["sm_mhz"] == 1965 and v["gpu_launches"] > 0
    assert 0.85 < v["roofline"]["frac"] < 1.0 and f"{v['roofline']['frac']:.2f}" in DESIGN and f"{v['roofline']['frac_of_nominal']:.2f}" in DESIGN
    assert f"{v['cpu_baseline']['value']:.0f} s" in DESIGN and v["cpu_baseline"]["extrapolated_by_op_counts"] is True
    v8 = bench("bench_r02_8gpu_v1.json")
    assert v8["n_gpus"] == 8 and f"{v8['value']:.3f} s" in DESIGN and f"{v8['e2e']['value']:.3f} s" in DESIGN
    for n_gpus in (1, 2, 4, 8):
        rows = {r["log_n"]: r for r in jsonl(f"msm_sharded_r02_n{n_gpus}.jsonl")}
        assert rows[26]["world"] == n_gpus
        assert f"{rows[26]['uniform']['ms']:.2f} / {rows[26]['witness_like']['ms']:.2f}" in DESIGN, n_gpus
    ab = {r["log_n"]: r for r in jsonl("affine_ab_r02.jsonl")}
    assert ab[24]["uniform_equal"] and ab[24]["witness_equal"]
    assert f"{ab[24]['affine_uniform_prof']['msm_accumulate']:.1f} ms" in DESIGN and f"{ab[24]['xyzz_uniform_prof']['msm_accumulate']:.1f} ms" in DESIGN
    drop = {(r["op"], r["log_n"]): r for r in jsonl("dropin_r02.jsonl")}
    assert f"{drop[('best_fft', 24)]['host_pinned_in_out_ms']:.1f} ms pinned" in DESIGN
    g1 = {r["k"]: r for r in jsonl("g1fft_r02.jsonl")}
    assert f"k = 21 {g1[21]['s_best']:.2f} s" in DESIGN
    peak = json.load(open(os.path.join(ROOT, "IMAD_PEAK.json")))
    assert peak["clocks"]["sm_mhz_median_under_load"] == 1965 and 8400 < peak["montgomery_product_gmac32"] < 8600
