"""Summaries of ncu captures for profiles/ (run where the .ncu-rep / csv files are):
  python tools/ncu_summary.py launches <gpu__time_duration csv> <out csv>
  python tools/ncu_summary.py full <`ncu -i x.ncu-rep --page raw --csv` output> <out txt> <out traffic json> <n_obs>
  python tools/ncu_summary.py stamp <traffic json>      record, per kernel, the hash of the source file it lives in
                                                         (bench.py ignores entries whose source changed since)
"""
import csv
import json
import re
import sys
from collections import OrderedDict

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
           "launch__shared_mem_per_block_dynamic",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def short(name):
    m = re.match(r"(?:void )?(?:lcba::)?(\w+)", name)
    return m.group(1) if m else name


def read_csv(path):
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = list(csv.reader(lines))
    hdr = next(i for i, r in enumerate(rd) if "Kernel Name" in r)
    return rd[hdr], rd[hdr + 1:]


def rows_of(path):          # --page raw: header, units row, one row per launch
    h, rows = read_csv(path)
    return h, rows[0], rows[1:]


KERNEL_FILES = {"k_i8_syrk": "schur_i8.cuh", "k_i8_make": "schur_i8.cuh", "k_i8_rowmax": "schur_i8.cuh",
                "k_jacobian_blocks": "eval.cuh"}


def stamp(path):
    import hashlib
    import os
    with open(path) as f:
        tr = json.load(f)
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lasercalib_b200", "csrc")
    tr["kernels_sha16"] = {}
    for k, fn in KERNEL_FILES.items():
        with open(os.path.join(csrc, fn), "rb") as f:
            tr["kernels_sha16"][k] = [fn, hashlib.sha256(f.read()).hexdigest()[:16]]
    with open(path, "w") as f:
        json.dump(tr, f, indent=1)


def launches(src, dst):     # --metrics gpu__time_duration.sum --csv: one row per (launch, metric)
    h, rows = read_csv(src)
    kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    to_ms = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3}
    agg = OrderedDict()
    for r in rows:
        if len(r) <= mv or not r[mv]:
            continue
        a = agg.setdefault(short(r[kn]), [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", "")) * to_ms.get(r[mu], 1e-6)
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("kernel,launches,total_ms,share\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%s,%d,%.3f,%.4f\n" % (k, a[0], a[1], a[1] / tot))


def full(src, dst_txt, dst_json, n_obs):
    h, units, rows = rows_of(src)
    kn = h.index("Kernel Name")
    seen = OrderedDict()
    for r in rows:
        if len(r) > kn and short(r[kn]) not in seen:
            seen[short(r[kn])] = r
    traffic = {}
    with open(dst_txt, "w") as f:
        for k, r in seen.items():
            f.write("kernel %s\n" % k)
            tot = 0.0
            for m in METRICS:
                if m in h:
                    i = h.index(m)
                    f.write("  %-80s %s %s\n" % (m, r[i], units[i]))
                    if m.startswith("dram__bytes"):
                        tot += float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
            stalls = []
            for i, name in enumerate(h):
                mm = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio|"
                              r"smsp__average_warp_latency_issue_stalled_(\w+)\.ratio", name)
                if mm and i < len(r) and r[i]:
                    try:
                        stalls.append((mm.group(1) or mm.group(2), float(r[i].replace(",", ""))))
                    except ValueError:
                        pass
            stalls.sort(key=lambda kv: -kv[1])
            s = sum(v for _, v in stalls) or 1.0
            f.write("  top_stalls %s\n" % ", ".join("%s %.0f%%" % (a, 100 * v / s) for a, v in stalls[:6]))
            f.write("  dram_bytes_total %.0f\n\n" % tot)
            traffic[k] = tot
    with open(dst_json, "w") as f:
        json.dump({"n_obs": int(n_obs), "dram_bytes_per_launch": traffic}, f, indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "stamp":
        stamp(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5])
