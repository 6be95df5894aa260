#!/bin/bash
# per-kernel device times + DRAM bytes of one layer fwd+bwd (cheap ncu pass: 3 metrics)
cd "$(dirname "$0")/.."
python tools/layer_prof.py "$@" > gpurun_out/plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file gpurun_out/layer_times.csv python tools/layer_prof.py "$@" > gpurun_out/ncu_layer.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/layer_times.csv")) if len(r) > 10]
h = rows[0]; ki, mi, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
idi = h.index("ID")
d = collections.OrderedDict()
for r in rows[1:]:
    k = (r[idi], r[ki].split("(")[0].replace("void ", "").replace("moe::", "")[:60])
    d.setdefault(k, {})[r[mi]] = float(r[vi].replace(",", ""))
tot = 0
for (i, n), m in d.items():
    us = m.get("gpu__time_duration.sum", 0) / 1e3
    tot += us
    by = m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
    print(f"{us:8.1f} us  {by / 1e6:8.1f} MB  {by / max(us, 1e-9) / 1e3:7.0f} GB/s  {n}")
print(f"total {tot:.1f} us")
PY
