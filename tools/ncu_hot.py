"""Top stall-sample lines of one kernel from an ncu report (SASS view).
usage: python tools/ncu_hot.py <report.ncu-rep> <kernel-id e.g. ::regex:grouped_gemm:4> [topN]"""
import csv, io, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-id", kid] if kid != "x" else [])
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# the page may hold several kernels, each with its own "Kernel Name" + header rows: take section `sec`
sec = int(sys.argv[4]) if len(sys.argv) > 4 else 0
starts = [i for i, r in enumerate(rows) if "# Samples" in r]
h0 = starts[sec]
hdr = rows[h0]
end = starts[sec + 1] - 1 if sec + 1 < len(starts) else len(rows)
data = [r for r in rows[h0 + 1:end] if len(r) == len(hdr)]
print("sections:", len(starts), "kernel:", rows[h0 - 1][1][:100] if h0 > 0 and len(rows[h0 - 1]) > 1 else "")
si, src = hdr.index("# Samples"), hdr.index("Source")
stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si]) for r in data)
print("total samples", tot)
agg = {h: sum(int(r[i]) for r in data) for i, h in stalls}
print("by reason:", {h: f"{100 * v / tot:.1f}%" for h, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
idx = sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:top]
for i in sorted(idx):
    r = data[i]
    why = max(stalls, key=lambda s: int(r[s[0]]))[1]
    print(f"{i:5d} {int(r[si]):6d} {100 * int(r[si]) / tot:5.1f}%  {why:18s} {r[src][:90]}")
