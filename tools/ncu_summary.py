"""Summaries of ncu captures for profiles/ (run here, on the CPU box, on files brought back in gpurun_out/).

  ncu_summary.py launches gpurun_out/launches_r02.csv  -> per-kernel launch count / total device time / share   (CSV to stdout)
  ncu_summary.py full gpurun_out/prof_r02.ncu-rep       -> selected --set full metrics per captured launch       (JSON to stdout)
"""
import csv
import io
import json
import subprocess
import sys
from collections import defaultdict

FULL_METRICS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
                "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
                "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
                "sm__cycles_elapsed.max"]


def short(name: str) -> str:
    name = name.split("(")[0].strip()
    return name.replace("b200zk::", "").replace("void ", "")


def launches(path):
    rows = [l for l in open(path, errors="replace").read().splitlines() if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("\n".join(rows)))
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else (v if unit in ("ms", "msecond") else v * 1e3))
        k = short(r["Kernel Name"])
        tot[k] += ms
        cnt[k] += 1
    # one-time setup of the bench process (SRS generation and its precomputed tables, twiddle tables, torch's input
    # generators and copies) is listed apart: the shares are taken over the kernels of the timed step
    setup = lambda k: k.startswith(("srs_shift", "g1_generator_mul", "ntt_build_table", "at::", "void at::")) or "at::" in k
    step_total = sum(v for k, v in tot.items() if not setup(k))
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "launches", "total_ms", "share_of_step_pct"])
    for k in sorted((k for k in tot if not setup(k)), key=lambda x: -tot[x]):
        w.writerow([k, cnt[k], round(tot[k], 3), round(100 * tot[k] / step_total, 2)])
    w.writerow(["STEP KERNELS TOTAL", sum(c for k, c in cnt.items() if not setup(k)), round(step_total, 3), 100.0])
    for k in sorted((k for k in tot if setup(k)), key=lambda x: -tot[x]):
        w.writerow(["setup: " + k[:90], cnt[k], round(tot[k], 3), ""])


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = [l for l in out.splitlines() if l.startswith('"')]
    rd = list(csv.reader(io.StringIO("\n".join(rows))))
    header, units, data = rd[0], rd[1], rd[2:]
    idx = {h: i for i, h in enumerate(header)}
    kernels = []
    for r in data:
        e = {"kernel": short(r[idx["Kernel Name"]])}
        for m in FULL_METRICS:
            if m in idx:
                e[m] = f"{r[idx[m]]} {units[idx[m]]}".strip()
        kernels.append(e)
    print(json.dumps({"source": path, "kernels": kernels}, indent=0))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
