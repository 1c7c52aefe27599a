"""Turn an `ncu --set full` capture into the committed evidence under profiles/.

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_extract.py /tmp/raw.csv profiles/r02a_ncu_full_summary.csv [--traffic-regex pio_gemm2_kernel]

Writes (1) a per-launch summary CSV with the handful of metrics the roofline discussion uses (duration, DRAM bytes,
tensor-pipe / DRAM / L2 / SM / XU utilisation, registers, grid, shared memory) and (2) profiles/roofline_traffic.json:
dram__bytes_read.sum + dram__bytes_write.sum per launch averaged over the launches whose kernel name matches
--traffic-regex (the dominant kernel family of bench.py's roofline), with the commit the capture belongs to.  bench.py
reads that file for `roofline.traffic` instead of carrying a literal.
"""
import argparse
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic"]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
              "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw_csv")
    ap.add_argument("summary_csv")
    ap.add_argument("--traffic-regex", default="pio_gemm2_kernel")
    ap.add_argument("--algorithmic-bytes", type=float, default=None,
                    help="algorithmic bytes per launch of the matched kernels (average), recorded next to the traffic")
    ap.add_argument("--no-traffic-json", action="store_true")
    a = ap.parse_args()
    rows = list(csv.reader(l for l in open(a.raw_csv, newline="") if l.startswith('"') or l.startswith("ID")))
    header, units, body = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(header)}
    keep = [k for k in KEEP if k in col]
    missing = [k for k in KEEP if k not in col]
    if missing:
        print("metrics not in the capture:", ", ".join(missing), file=sys.stderr)
    name_i = col["Kernel Name"]
    with open(a.summary_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["ID", "Kernel Name"] + keep)
        w.writerow(["", ""] + [units[col[k]] for k in keep])
        for r in body:
            w.writerow([r[col["ID"]], r[name_i]] + [r[col[k]] for k in keep])
    print(f"{a.summary_csv}: {len(body)} launches")
    if a.no_traffic_json:
        return
    rx = re.compile(a.traffic_regex)
    picked = [r for r in body if rx.search(r[name_i])]
    if not picked:
        print("no launch matches", a.traffic_regex, file=sys.stderr)
        return

    def val(r, k):
        return float(r[col[k]].replace(",", "")) * UNIT_SCALE.get(units[col[k]], 1.0)

    traffic = [val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in picked]
    dur = [val(r, "gpu__time_duration.sum") for r in picked]
    try:
        commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True,
                                text=True).stdout.strip()
    except Exception:
        commit = None
    out = {"traffic_bytes_per_launch": sum(traffic) / len(traffic),
           "algorithmic_bytes_per_launch": a.algorithmic_bytes,
           "launches_averaged": len(picked), "mean_duration_ms_under_ncu": sum(dur) / len(dur),
           "kernels": sorted({re.sub(r"\(.*", "", r[name_i]).replace("void ", "") for r in picked}),
           "source": os.path.relpath(a.summary_csv, ROOT), "commit": commit,
           "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch from one `ncu --set full --clock-control none` "
                   "capture (cold caches, serialised launches)"}
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    json.dump(out, open(p, "w"), indent=1)
    print(p, json.dumps(out))


if __name__ == "__main__":
    main()
