"""Writes IMAD_PEAK.json (repo root, tracked): the INT32 multiply-pipe roofline denominators bench.py reports against.

Runs bench_micro/microbench on the GPU while sampling SM clocks / throttle reasons with nvidia-smi, takes the sustained
Montgomery-product rate (`fr_mul_sustained`, ~1 s of back-to-back products) and the best plain-IMAD rate, and records
them with the clocks seen under load.  MEASURED_PEAKS.json (driver-written) has no INT32 figure, hence this file.

    python tools/imad_peak.py            # on a B200 box; also appends the raw lines to profiles/microbench_r02.jsonl
    python tools/imad_peak.py --from profiles/microbench_r01.jsonl profiles/microbench_r01_clocks.csv   # re-derive offline
"""
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def derive(lines, clock_rows, source):
    recs = [json.loads(l) for l in lines if l.startswith("{")]
    dev = next(r for r in recs if "device" in r)
    sustained = next(r for r in recs if r.get("bench") == "fr_mul_sustained")
    imad = max(r["Gops"] for r in recs if r.get("bench") == "imad_lo")
    bfly = max(r["Gops"] for r in recs if r.get("bench") == "fr_bfly")
    sm = sorted(int(r[0]) for r in clock_rows if r and r[0].isdigit())
    mx = max([int(r[1]) for r in clock_rows if len(r) > 1 and r[1].isdigit()] or [dev["clock_khz"] // 1000])
    loaded = [x for x in sm if x > 0.5 * mx] or sm
    sm_med = loaded[len(loaded) // 2] if loaded else None
    reasons = sorted({r[3] for r in clock_rows if len(r) > 3 and r[3] not in ("0x0000000000000000", "0x0000000000000001", "")})
    return {
        "montgomery_product_gmodmul": sustained["Gops"],
        "montgomery_product_gmac32": round(sustained["Gops"] * 128, 1),       # MODMUL = 128 MAC32
        "butterfly_gops": bfly,
        "imad_gops": imad,
        "nominal_imad_wide_gmac32": round(dev["sms"] * 32 * mx / 1e3, 1),     # SMs x 32 IMAD.WIDE lanes/clk x max SM clock (MHz)
        "device": dev["device"], "sms": dev["sms"],
        "clocks": {"sm_mhz_median_under_load": sm_med, "sm_max_mhz": mx, "samples": len(sm), "throttle_reason_bitmasks": reasons},
        "source": source, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
    }


def clock_rows_from_csv(path):
    rows = []
    for l in open(path).read().splitlines()[1:]:
        p = [x.strip() for x in l.split(",")]
        rows.append([p[0].split()[0], p[1].split()[0], p[2].split()[0], p[3] if len(p) > 3 else ""])
    return rows


def main():
    if len(sys.argv) >= 4 and sys.argv[1] == "--from":
        lines = open(sys.argv[2]).read().splitlines()
        rec = derive(lines, clock_rows_from_csv(sys.argv[3]), f"{sys.argv[2]} + {sys.argv[3]}")
    else:
        rows = []
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active"
        mon = subprocess.Popen(["nvidia-smi", "-i", "0", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                               stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        threading.Thread(target=lambda: [rows.append([x.strip() for x in l.split(",")]) for l in mon.stdout], daemon=True).start()
        out = subprocess.run([os.path.join(ROOT, "bench_micro", "microbench")], capture_output=True, text=True, check=True).stdout
        mon.terminate()
        open(os.path.join(ROOT, "profiles", "microbench_r02.jsonl"), "w").write(out)
        rec = derive(out.splitlines(), rows, "bench_micro/microbench run by tools/imad_peak.py (profiles/microbench_r02.jsonl)")
    json.dump(rec, open(os.path.join(ROOT, "IMAD_PEAK.json"), "w"), indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
