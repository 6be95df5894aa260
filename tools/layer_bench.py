"""Isolated-layer breakdown on the GPU (same code path as bench.py's `moe_layer` object).
usage: python tools/layer_bench.py [T d E k cf]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "slim-switch-moe-vit_b200")]
import bench
a = [float(v) for v in sys.argv[1:]]
kw = {}
if len(a) >= 4:
    kw = dict(T=int(a[0]), d=int(a[1]), E=int(a[2]), k=int(a[3]), cf=a[4] if len(a) > 4 else 1.25)
r = bench.layer_bench(bench.load_peaks(), **kw)
print(json.dumps({k: v for k, v in r.items() if k != "kernels"}))
tot = 0.0
for k, v in sorted(r["kernels"].items(), key=lambda kv: -kv[1]["ms"]):
    tot += v["ms"]
    print(f"{k:24s} {v['ms'] * 1e3:8.1f} us  " + "  ".join(f"{a}={b}" for a, b in v.items() if a not in ("ms", "calls_per_iter")))
print(f"sum of kernels {tot * 1e3:.1f} us")
