#!/usr/bin/env python
"""Condense `ncu --page raw --csv` exports (gpurun_out/*_raw.csv) and the launch list into small tables under profiles/.

    python tools/ncu_summary.py raw gpurun_out/r01a_tc_raw.csv profiles/r01_ncu_tc_summary.csv
    python tools/ncu_summary.py launches gpurun_out/r01a_launches.csv profiles/r01_launches_summary.csv
"""
import csv
import re
import sys
from collections import defaultdict

KEEP = [
    ("Kernel Name", "kernel"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("gpu__time_duration.sum", "time_us"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "xbar2sm_bytes"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "xbar2sm_TBps"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("sm__cycles_elapsed.avg", "sm_cycles"),
    ("sm__cycles_elapsed.avg.per_second", "sm_clock"),
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6, "usecond": 1.0,
        "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}


def find_col(hdr, name):
    for i, h in enumerate(hdr):
        if h == name:
            return i
    for i, h in enumerate(hdr):
        if name in h:
            return i
    return None


def raw(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r and not r[0].startswith("==")]
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [(find_col(hdr, n), lab) for n, lab in KEEP]
    with open(out, "w", newline="") as fp:
        w = csv.writer(fp)
        w.writerow([lab for i, lab in cols if i is not None])
        for r in data:
            line = []
            for i, lab in cols:
                if i is None:
                    continue
                v = r[i]
                u = units[i]
                if lab == "kernel":
                    v = re.sub(r"\(.*", "", v)[:60]
                elif u in UNIT and lab in ("dram_read", "dram_write", "l2_bytes", "xbar2sm_bytes", "time_us"):
                    v = f"{float(v.replace(',', '')) * UNIT[u]:.6g}"
                line.append(v)
            w.writerow(line)
    print(open(out).read())


def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ik])
        name = re.sub(r"^void |\(anonymous namespace\)::", "", name)
        t = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[iu], 1.0)
        agg[name][0] += 1
        agg[name][1] += t
    total = sum(v[1] for v in agg.values())
    with open(out, "w", newline="") as fp:
        w = csv.writer(fp)
        w.writerow(["kernel", "launches", "total_us", "share_pct", "avg_us"])
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, f"{t:.1f}", f"{100 * t / total:.2f}", f"{t / n:.2f}"])
        w.writerow(["TOTAL", sum(v[0] for v in agg.values()), f"{total:.1f}", "100.00", ""])
    print(open(out).read())


if __name__ == "__main__":
    {"raw": raw, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
