#!/usr/bin/env python
"""DRAM traffic per launch of the six grouped-GEMM ops from one `ncu --set full` capture of tools/gemm_bench.py
(config-2 layer shape), written to profiles/gemm_dram_traffic.json — the file bench.py reads for `roofline.traffic`.

    python tools/gemm_traffic.py gpurun_out/prof_gemm_r1f.ncu-rep profiles/gemm_dram_traffic.json
"""
import csv, io, json, re, subprocess, sys

OPS = {0: "gemm_fc1", 1: "gemm_fc2", 2: "gemm_dgelu", 3: "gemm_dgrad", 4: "gemm_wgrad1", 5: "gemm_wgrad2"}   # EPI template argument
src, dst = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ki, ri, wi, ti = (hdr.index(n) for n in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = {}
for r in rows[2:]:
    m = re.search(r"grouped_gemm_kernel<\(?(?:int\))?(\d+), \(?(?:int\))?(\d+),", r[ki])
    if not m:
        continue
    op = OPS[int(m.group(2))]
    rd = float(r[ri].replace(",", "")) * scale[units[ri]]
    wr = float(r[wi].replace(",", "")) * scale[units[wi]]
    a = acc.setdefault(op, {"bn": int(m.group(1)), "launches": 0, "dram_read": 0.0, "dram_write": 0.0, "us": 0.0})
    a["launches"] += 1; a["dram_read"] += rd; a["dram_write"] += wr
    a["us"] += float(r[ti].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "ms": 1e3}.get(units[ti], 1.0)
res = {"source": src, "shape": "T=50432 d=384 h=1536 E=16 (3152 rows per expert), tools/gemm_bench.py", "per_op": {}}
for op, a in sorted(acc.items()):
    n = a["launches"]
    res["per_op"][op] = {"tile_n": a["bn"], "launches": n, "dram_bytes_read": round(a["dram_read"] / n), "dram_bytes_write": round(a["dram_write"] / n),
                         "dram_bytes": round((a["dram_read"] + a["dram_write"]) / n), "ncu_us": round(a["us"] / n, 1)}
res["mean_dram_bytes_per_launch"] = round(sum(v["dram_bytes"] for v in res["per_op"].values()) / max(1, len(res["per_op"])))
json.dump(res, open(dst, "w"), indent=1)
print(json.dumps(res, indent=1))
