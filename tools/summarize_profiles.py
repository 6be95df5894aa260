#!/usr/bin/env python
"""Turns ncu outputs under gpurun_out/ into the small, tracked summaries under profiles/.

    python tools/summarize_profiles.py launches gpurun_out/launches.csv profiles/r01_launches.md "title"
    python tools/summarize_profiles.py full     gpurun_out/prof_gemm.ncu-rep profiles/r01_gemm_ncu.md "title"
"""
import collections
import csv
import io
import re
import subprocess
import sys

FULL_METRICS = [
    ("gpu__time_duration.sum", "dur_us"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_active_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("lts__t_bytes.sum", "l2_MB"),
    ("sm__cycles_elapsed.max", "cycles"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_%"),
]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(CUtensorMap.*", "", name)
    name = re.sub(r"\((?:const |int|float|long|void|__nv|T\d|at::|std::|unsigned|char).*", "", name)
    return name[:110]


def launches(src, dst, title):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in agg.values())
    mine = sum(v for n, (_, v) in agg.items() if n.startswith("moe::"))
    with open(dst, "w") as f:
        f.write(f"# {title}\n\nsource: `{src}` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, "
                f"serialised launches: compare SHARES, not absolutes)\n\n")
        f.write(f"launches: {sum(c for c, _ in agg.values())}, total {tot / 1e3:.1f} us; `moe::` kernels {mine / 1e3:.1f} us "
                f"= {100 * mine / tot:.1f}% of the window\n\n| us | share | launches | kernel |\n|---:|---:|---:|---|\n")
        for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
            if v / tot < 0.002 and not n.startswith("moe::"):
                continue
            f.write(f"| {v / 1e3:.1f} | {100 * v / tot:.1f}% | {c} | `{n}` |\n")


def full(src, dst, title):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    cols = [(hdr.index(m), lab, units[hdr.index(m)]) for m, lab in FULL_METRICS if m in hdr]
    with open(dst, "w") as f:
        f.write(f"# {title}\n\nsource: `{src}` (ncu --set full --clock-control none --import-source on; per-launch values)\n\n")
        f.write("| kernel | " + " | ".join(f"{lab} ({u})" if u else lab for _, lab, u in cols) + " |\n")
        f.write("|---|" + "---:|" * len(cols) + "\n")
        for r in rows[2:]:
            vals = []
            for i, _, _ in cols:
                try:
                    vals.append(f"{float(r[i].replace(',', '')):.1f}")
                except ValueError:
                    vals.append(r[i])
            f.write(f"| `{short(r[ki])}` | " + " | ".join(vals) + " |\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:5])
