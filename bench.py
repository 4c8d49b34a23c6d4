#!/usr/bin/env python
"""bench.py — chunk-proof polynomial-arithmetic wall time on B200 (BASELINE.json metric).

A "step" is ONE pass of the hot path over one proof's worth of synthetic columns: the KZG commits
(MSM) and NTTs that halo2's create_proof issues for the degree-24 layer of a chunk proof
(BASELINE.json configs[1], "inner-prove ... degree-24 SRS"; shape from
/root/reference/integration/configs/layer1.config:3-9 as derived in SURVEY.md §8(d).2):

    28 columns  x [commit_lagrange = MSM 2^k  ->  lagrange_to_coeff = iNTT 2^k  ->  coeff_to_extended = coset NTT 2^(k+2)]
     7 commits  of coefficient-form polynomials (4 quotient pieces + 3 SHPLONK)   = MSM 2^k over g
     1 extended_to_coeff                                                            = inverse coset NTT 2^(k+2)

19 of the 35 MSMs use witness-like scalars (60 % zero / 30 % < 2^16 / 10 % uniform), 16 uniform.  The
real trace cannot be proved here (no Rust toolchain, no SRS files, host witness generation out of scope),
so the polynomial-arithmetic layer is replayed on synthetic data of exactly that shape: "data": "synthetic".

JSON keys: see README/DESIGN.md "Measurement".  `value` = device-resident seconds per step (inputs in HBM),
`e2e` = the same step through the public session API with HOST (pinned) columns, H2D inside the timed region,
commitments read back.  With --gpus N (torchrun) the step's independent jobs are fanned out over the ranks
(strong scaling, no data-path collective; one tiny NCCL all_gather of the commitments).

--impl reference times the CPU restatement of the reference's Rayon path (oracle/, all host threads) on a
bounded sample of the same step; the default arm never touches oracle/ outside its `cpu_baseline` leg.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_WITNESS_COLS, N_UNIFORM_COLS, N_COEFF_MSMS = 19, 9, 7
def metric_name(k: int) -> str:
    return f"chunk-proof wall-sec (degree-{k} layer poly-arith replay)"
# measured on this pool's B200 (profiles/microbench_r01.jsonl): sustained Montgomery products/s and the
# INT32 multiply-pipe rate they imply; nominal = 148 SM x 32 IMAD.WIDE lanes/clk x 1.965 GHz
IMAD_PEAK_MEASURED_GMAC32 = 8500.0
IMAD_PEAK_NOMINAL_GMAC32 = 9307.0
MAC32_PER_MADD = 1280  # 10 modmul x 128 MAC32


def workload_desc(k: int) -> str:
    return (f"configs[1] inner-prove degree-{k} layer shape replay: {N_WITNESS_COLS + N_UNIFORM_COLS} columns x "
            f"(MSM 2^{k} + iNTT 2^{k} + cosetNTT 2^{k + 2}) + {N_COEFF_MSMS} MSM 2^{k} + 1 icosetNTT 2^{k + 2}")


def make_jobs():
    """(kind, scalar distribution, relative cost) for every independent unit of the step."""
    # A column's commitment (MSM) and its transforms (iNTT + coset NTT) are independent, so they are separate jobs.
    # relative costs measured on B200 at k = 24 (ms): witness-like MSM 11, uniform MSM 44, transform chain 18, icoset 16
    jobs = [("lmsm", "w", 11.0)] * N_WITNESS_COLS + [("lmsm", "u", 44.0)] * N_UNIFORM_COLS   # commit_lagrange
    jobs += [("ntt", "w", 18.0)] * N_WITNESS_COLS + [("ntt", "u", 18.0)] * N_UNIFORM_COLS      # lagrange_to_coeff + coeff_to_extended
    jobs += [("msm", "u", 44.0)] * N_COEFF_MSMS + [("icoset", "u", 16.0)]
    return jobs


def assign_jobs(jobs, world: int):
    """LPT greedy fan-out of the independent jobs over ranks."""
    order = sorted(range(len(jobs)), key=lambda i: -jobs[i][2])
    load = [0.0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda x: load[x])
        out[r].append(jobs[i])
        load[r] += jobs[i][2]
    return out


def host_threads() -> int:
    """Threads the reference's Rayon pool would get: min(logical CPUs, affinity mask, cgroup CPU quota)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(int(q) / int(per) + 0.5)))
    except Exception:
        pass
    return n


class ClockSampler:
    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = max([int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        loaded = [x for x in sm if x > 0.5 * (mx or 1)] or sm
        return {"sm_mhz": loaded[len(loaded) // 2] if loaded else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(sm)}


# ====================================================================================================
# reference arm / cpu_baseline: the oracle's restatement of halo2_proofs' Rayon path on the host cores
# ====================================================================================================
_CPU_INPUTS = {}


def _cpu_inputs(k: int, threads: int):
    """bases / scalars / domain for the CPU legs, generated once per size (not part of any timed region)."""
    from oracle import oracle as O  # ORACLE: allowed here only (cpu_baseline / --impl reference)

    if k not in _CPU_INPUTS:
        n = 1 << k
        _CPU_INPUTS[k] = (O.fill_points_chain(n, 7, threads), O.fill_fr(n, 1, True), O.fill_fr(n, 2, False), O.EvaluationDomain(5, k))
    return _CPU_INPUTS[k]


def cpu_sample(k: int, threads: int, calibrate: bool = True):
    """Times a bounded sample of the step on the CPU: 1 witness-like MSM, 1 uniform MSM, 1 iNTT 2^k,
    1 coset NTT 2^(k+2); the step is extrapolated by op counts.  Returns (step_seconds, detail)."""
    from oracle import oracle as O  # ORACLE: allowed here only (cpu_baseline / --impl reference)

    n = 1 << k
    bases, sw, su, dom = _cpu_inputs(k, threads)
    t = {}
    t0 = time.perf_counter(); O.best_multiexp(sw, bases, threads); t["msm_w"] = time.perf_counter() - t0
    t0 = time.perf_counter(); O.best_multiexp(su, bases, threads); t["msm_u"] = time.perf_counter() - t0
    t0 = time.perf_counter(); coeff = dom.lagrange_to_coeff(su, threads); t["intt"] = time.perf_counter() - t0
    t0 = time.perf_counter(); dom.coeff_to_extended(coeff, threads); t["coset"] = time.perf_counter() - t0
    if calibrate:
        # how many of the nominal host threads actually run in parallel (containers are often CPU-quota limited):
        # one 2^14-point chunk alone vs one such chunk per thread, all at once
        m = min(1 << 14, n)
        reps = np.tile(su[:m], (threads, 1)); rb = np.tile(bases[:m], (threads, 1))
        t1 = tT = 1e30
        for _ in range(2):
            t0 = time.perf_counter(); O.best_multiexp(su[:m], bases[:m], 1); t1 = min(t1, time.perf_counter() - t0)
            t0 = time.perf_counter(); O.best_multiexp(reps, rb, threads); tT = min(tT, time.perf_counter() - t0)
        t["effective_parallelism"] = round(threads * t1 / tT, 1) if tT > 0 else None
        t["affinity"] = len(os.sched_getaffinity(0))
        for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
            if os.path.exists(f):
                t["cgroup_cpu"] = open(f).read().strip()
                break
    step = (N_WITNESS_COLS * t["msm_w"] + (N_UNIFORM_COLS + N_COEFF_MSMS) * t["msm_u"]
            + (N_WITNESS_COLS + N_UNIFORM_COLS) * (t["intt"] + t["coset"]) + t["coset"])
    return step, t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = host_threads()
    _cpu_inputs(args.k, threads)
    vals, detail, budget_s, t_start, reduced = [], None, 240.0, time.perf_counter(), 0
    scale = None
    for i in range(args.warmup + args.steps):
        elapsed = time.perf_counter() - t_start
        left = args.warmup + args.steps - i
        if scale is not None and vals_full and elapsed + left * vals_full[-1][1] > budget_s and args.k >= 16:
            # keep the whole run within a few minutes: the remaining iterations time the same ops two sizes down and
            # are scaled by the full/small ratio measured on this box
            st, _ = cpu_sample(args.k - 2, threads, calibrate=False)
            step, reduced = st * scale, reduced + 1
        else:
            t0 = time.perf_counter()
            step, detail = cpu_sample(args.k, threads, calibrate=(i == 0))
            if i == 0:
                vals_full = []
                small, _ = cpu_sample(args.k - 2, threads, calibrate=False) if args.k >= 16 else (None, None)
                scale = (step / small) if small else None
            vals_full.append((step, time.perf_counter() - t0))
        if i >= args.warmup:
            vals.append(step)
    v = sum(vals) / len(vals)
    sample = ("per step: 1 witness-like + 1 uniform MSM 2^%d, 1 iNTT 2^%d, 1 coset NTT 2^%d timed on %d host threads; step "
              "extrapolated by op counts (19/16 MSM, 28 iNTT, 29 coset NTT); %d of %d iterations ran the same ops at 2^%d "
              "scaled by the measured full/small ratio to bound the run" % (args.k, args.k, args.k + 2, threads, reduced,
                                                                              args.warmup + args.steps, args.k - 2))
    line = {"impl": "reference", "metric": metric_name(args.k), "value": v, "unit": "s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": v * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64x4 (254-bit Montgomery integers on the CPU)", "data": "synthetic", "config": {"workload": workload_desc(args.k), "k": args.k},
            "cpu_baseline": {"value": v, "unit": "s", "cores": threads, "kind": "port", "sample": sample, "detail_s": detail},
            "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ====================================================================================================
# product arm
# ====================================================================================================
def run_b200(args):
    import torch

    zk = importlib.import_module("scroll-prover_b200")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    k, n, ek = args.k, 1 << args.k, args.k + 2
    comp = torch.cuda.Stream()  # a non-default stream, so that our CUDA events bracket the library's launches
    torch.cuda.set_stream(comp)
    ctx = zk.Context(local)
    ctx.set_stream(comp.cuda_stream)
    dom = zk.EvaluationDomain(ctx, 5, k)

    # ---- synthetic inputs (device-generated)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)

    def rand_limbs(cnt):
        t = torch.randint(-(2 ** 63), 2 ** 63 - 1, (cnt, 4), dtype=torch.int64, device=dev, generator=gen)
        t[:, 3] &= 0x0FFFFFFFFFFFFFFF  # < 2^252 < r: valid Montgomery limbs of a uniform-ish element
        return t

    def witness_like(cnt):
        sel = torch.rand(cnt, device=dev, generator=gen)
        small = torch.randint(0, 1 << 16, (cnt,), dtype=torch.int64, device=dev, generator=gen)
        raw = torch.zeros((cnt, 4), dtype=torch.int64, device=dev)
        raw[:, 0] = torch.where((sel >= 0.6) & (sel < 0.9), small, torch.zeros_like(small))
        r2 = zk.fr_from_int(1 << 256)  # Montgomery form of R: x * R2 * R^-1 = x*R -> to_mont
        torch.cuda.synchronize()
        mont = ctx.poly_scale(raw, r2)
        uni = rand_limbs(cnt)
        m = (sel >= 0.9).unsqueeze(1)
        torch.cuda.synchronize()
        return torch.where(m, uni, mont).contiguous()

    cols = {"w": [witness_like(n) for _ in range(2)], "u": [rand_limbs(n) for _ in range(2)]}
    hext = rand_limbs(1 << ek)
    srs_scalars = rand_limbs(n)
    g = torch.empty((n, 8), dtype=torch.int64, device=dev)
    gl = torch.empty((n, 8), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    ctx.g1_generator_mul_batch(srs_scalars, out=g)
    ctx.g1_generator_mul_batch(rand_limbs(n), out=gl)
    params = zk.ParamsKZG(ctx, k, g, gl)
    del g, gl, srs_scalars
    my_jobs = assign_jobs(make_jobs(), world)[rank]
    n_msm = sum(1 for j in my_jobs if j[0] in ("lmsm", "msm"))
    commits = np.zeros((len(my_jobs) + 1, 12), np.uint64)

    lmsm_jobs = [(j, job) for j, job in enumerate(my_jobs) if job[0] == "lmsm"]
    ntt_jobs = [(j, job) for j, job in enumerate(my_jobs) if job[0] == "ntt"]
    msm_jobs = [(j, job) for j, job in enumerate(my_jobs) if job[0] == "msm"]
    has_icoset = any(job[0] == "icoset" for job in my_jobs)

    def job_list(src, hsrc):
        """Everything this rank owns as ONE b200zk_run_column_jobs list: a column whose commitment and transforms both
        landed here is one mode-2 job (uploaded once in the e2e arm); the quotient's extended_to_coeff is a mode-4 job."""
        both = min(len(lmsm_jobs), len(ntt_jobs))
        jl = [(src[job[1]][j % 2], params._gl, 2, None, None) for j, job in lmsm_jobs[:both]]
        jl += [(src[job[1]][j % 2], params._gl, 0, None, None) for j, job in lmsm_jobs[both:]]
        jl += [(src[job[1]][j % 2], None, 3, None, None) for j, job in ntt_jobs[both:]]
        jl += [(src[job[1]][j % 2], params._g, 0, None, None) for j, job in msm_jobs]
        if has_icoset:
            jl.append((hsrc, None, 4, None, None))
        return jl

    def run_jobs(jl):
        if jl:
            commits[: len(jl)] = zk.run_column_jobs(ctx, jl, k, omega_inv=dom.omega_inv, extended_omega=dom.extended_omega,
                                                    extended_omega_inv=dom.extended_omega_inv, extended_k=ek)
        torch.cuda.current_stream().synchronize()

    resident_jobs = job_list(cols, hext)

    def step_resident():
        """inputs already resident in HBM; the library runs the commitments and the transforms on two streams"""
        run_jobs(resident_jobs)

    # ---- e2e: host (pinned) columns -> session API; H2D of job j+1 overlaps compute of job j
    host = {"w": [c.cpu().pin_memory() for c in cols["w"]], "u": [c.cpu().pin_memory() for c in cols["u"]],
            "h": hext.cpu().pin_memory() if any(j[0] == "icoset" for j in my_jobs) else None}
    # a column whose commitment and transforms are on the same rank crosses PCIe once (one mode-2 call)
    _nl, _nn = sum(1 for j in my_jobs if j[0] == "lmsm"), sum(1 for j in my_jobs if j[0] == "ntt")
    h2d_bytes = (max(_nl, _nn) + sum(1 for j in my_jobs if j[0] == "msm")) * n * 32 + \
        sum((1 << ek) * 32 for j in my_jobs if j[0] == "icoset")

    host_jobs = job_list(host, host["h"])
    d2h_bytes = len(host_jobs) * 96  # the commitments array of the one run_column_jobs call

    def step_e2e():
        """the same job list from pinned HOST buffers: H2D of job j+1 on the library's copy stream while job j computes"""
        run_jobs(host_jobs)

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_commits():
        if not dist:
            return
        t = torch.zeros((world, 80, 12), dtype=torch.int64, device=dev)
        mine = torch.zeros((80, 12), dtype=torch.int64, device=dev)
        mine[: len(commits)] = torch.from_numpy(commits.view(np.int64)).to(dev)
        dist.all_gather_into_tensor(t.view(-1), mine.view(-1))

    def timed(step_fn, steps, warmup, profile=False):
        for _ in range(warmup):
            step_fn()
            gather_commits()
        barrier()
        if profile:
            ctx.profile_enable(True)
            ctx.profile_reset()
            ctx.msm_total_adds(reset=True)
        l0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(comp)
        for _ in range(steps):
            step_fn()
            gather_commits()
        e1.record(comp)
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        prof = ctx.profile_read() if profile else None
        if profile:
            prof["_actual_adds"] = ctx.msm_total_adds()
        if profile:
            ctx.profile_enable(False)
        launches = ctx.launch_count() - l0
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        if dist:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return tt.item() / 1e3 / steps, wall / steps, prof, launches

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    sec, wall, _, launches = timed(step_resident, args.steps, args.warmup, profile=False)
    clocks = sampler.stop() if sampler else None
    # per-kernel durations (roofline, kernel_ms_per_step) come from a separate profiled pass of the SAME step with the
    # two-stream overlap switched off: while kernels of two streams share the SMs their individual durations say nothing
    ctx.set_overlap(False)
    prof_steps = 1
    _, _, prof, _ = timed(step_resident, prof_steps, 0, profile=True)
    ctx.set_overlap(True)
    if args.skip_e2e:
        e2e_sec = None
    else:
        e2e_sec, e2e_wall, _, _ = timed(step_e2e, max(1, min(args.steps, 3)), 1)

    # ---- units processed (whole job) for the throughput figures
    st = ctx.msm_last_stats()
    c_bits, W = st["window_bits"], st["n_windows"]
    total_msm = N_WITNESS_COLS + N_UNIFORM_COLS + N_COEFF_MSMS
    nw_adds = total_msm * n * W  # N*W upper bound (SURVEY §8(d) "G1-adds/s = N*W / t")
    bf = lambda lg: (1 << lg) // 2 * lg
    total_bf = (N_WITNESS_COLS + N_UNIFORM_COLS) * (bf(k) + bf(ek)) + bf(ek)

    line = None
    if rank == 0:
        my_msm = sum(1 for j in my_jobs if j[0] in ("lmsm", "msm"))
        actual_adds = prof.pop("_actual_adds")
        msm_ms = sum(prof[c]["ms"] for c in prof if c.startswith("msm_"))
        acc_ms, acc_cnt = prof["msm_accumulate"]["ms"], prof["msm_accumulate"]["count"]
        ntt_ms, ntt_cnt = prof["ntt_pass"]["ms"], prof["ntt_pass"]["count"]
        my_bf = sum(bf(k) + bf(ek) for j in my_jobs if j[0] == "ntt") + sum(bf(ek) for j in my_jobs if j[0] == "icoset")
        my_ntt_bytes = 0
        for j in my_jobs:  # algorithmic HBM bytes: 64 B/element/pass, P = ceil(log_n / 8) passes in this build
            if j[0] == "ntt":
                my_ntt_bytes += 64 * n * 3 + 64 * (1 << ek) * 4
            elif j[0] == "icoset":
                my_ntt_bytes += 64 * (1 << ek) * 4
        # algorithmic MAC32 of the accumulate launches = bucket additions actually performed (non-zero signed digits;
        # witness-like columns skip most of the N*W upper bound) x 1280 MAC32 per mixed add
        acc_achieved = (actual_adds * MAC32_PER_MADD) / (acc_ms * 1e-3) / 1e9 if acc_ms else None
        steps_prof = prof_steps
        traffic, traffic_note = None, None
        try:  # DRAM bytes of one msm_accumulate launch from the committed `ncu --set full` capture (uniform 2^24 MSM)
            ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_r01_v6_summary.json")))
            for e in ncu["kernels"]:
                if e["kernel"].startswith("msm_accumulate"):
                    rd, wr = float(e["dram__bytes_read.sum"].split()[0]), float(e["dram__bytes_write.sum"].split()[0])
                    traffic = (rd + wr) * 1e9
                    traffic_note = ("dram__bytes_read+write of ONE msm_accumulate launch (MSM 2^24, uniform scalars, 201 M adds) "
                                    "from profiles/ncu_r01_v6_summary.json; algorithmic bytes of that launch = 201e6 x 68 B = 13.7e9")
                    break
        except Exception:
            pass
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        ntt_gbs = my_ntt_bytes * steps_prof / (ntt_ms * 1e-3) / 1e9 if ntt_ms else None
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            cs, detail = cpu_sample(k, threads)
            cpu = {"value": cs, "unit": "s", "cores": threads, "kind": "port",
                   "sample": "1 witness-like + 1 uniform MSM 2^%d, 1 iNTT 2^%d, 1 coset NTT 2^%d on all host threads; step "
                             "extrapolated by op counts" % (k, k, ek), "detail_s": detail}
        line = {
            "metric": metric_name(args.k), "value": sec, "unit": "s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32x8 (254-bit Montgomery integers)", "data": "synthetic",
            "config": {"workload": workload_desc(k), "k": k, "msm_window_bits": c_bits, "msm_windows": W,
                       "parallelism": f"job fan-out x{world} (no data-path collective)",
                       "l2": "inputs (>= 512 MiB per column) exceed the 126 MB L2; no flush needed"},
            "msm_g1_adds_per_s": (my_msm * n * W * steps_prof) / (msm_ms * 1e-3) if msm_ms else None,
            "ntt_butterflies_per_s": (my_bf * steps_prof) / (ntt_ms * 1e-3) if ntt_ms else None,
            "msm_actual_bucket_adds_per_step_rank0": actual_adds / steps_prof,
            "job_g1_adds_per_s": nw_adds / sec, "job_ntt_butterflies_per_s": total_bf / sec,
            "roofline": {"bound": "int32-imad", "kernel": "msm_accumulate", "achieved": acc_achieved,
                         "peak": IMAD_PEAK_MEASURED_GMAC32, "peak_nominal": IMAD_PEAK_NOMINAL_GMAC32, "unit": "GMAC32/s",
                         "frac": acc_achieved / IMAD_PEAK_MEASURED_GMAC32 if acc_achieved else None,
                         "peak_source": "measured Montgomery-product microbenchmark (profiles/microbench_r01.jsonl), not in MEASURED_PEAKS.json",
                         "traffic": traffic, "traffic_note": traffic_note,
                         "hbm": {"kernel": "ntt_pass", "achieved": ntt_gbs, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": ntt_gbs / hbm_peak if ntt_gbs else None, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)"}},
            "kernel_ms_per_step_rank0": {c: prof[c]["ms"] / steps_prof for c in prof if prof[c]["count"]},
            "kernel_timing_note": "per-kernel CUDA-event durations and the roofline come from one extra profiled pass of the same "
                                  "step with the two-stream overlap off; `value` is measured with the overlap on and profiling off",
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_sec, "unit": "s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "api": "ONE C-ABI call per step with pinned HOST buffers: b200zk_run_column_jobs (mode 2 for the 28 columns, "
                           "mode 0 for the 7 coefficient commits, mode 4 for the quotient; H2D double-buffered on the library's "
                           "copy stream; rank 0 bytes)"},
            "gpu_launches": launches, "clocks": clocks, "wall_s_per_step": wall,
        }
        print(json.dumps(line))
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--k", type=int, default=24, help="log2 rows of the proved layer (24 = configs[1])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only (ncu): device-resident arm only")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
