"""Per-tile timeline of one grouped-GEMM launch (needs a -DMOE_DBG_TIMELINE build of the library):

    tools/build_variant.sh timeline -DMOE_DBG_TIMELINE
    MOE_B200_LIB=tools/variants/libmoe_timeline.so python tools/gemm_timeline.py [--op fc1] [--d 384] [--E 16] [--rows 3152]

Prints, for CTA 0 (leader) and CTA 1 (peer) of cluster 0, clock64 stamps relative to the first one, per tile:
  producer: tile start / first slot free / last TMA issued;   MMA: before tempty wait / after / first operands landed / commit;
  epilogue warp 2 and warp 6: tile start / accumulator full / TMEM released / tile done."""
import argparse, ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "slim-switch-moe-vit_b200"))
import torch  # noqa: E402
from fmoe import _cabi as C  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--op", default="fc1")
ap.add_argument("--d", type=int, default=384)
ap.add_argument("--E", type=int, default=16)
ap.add_argument("--rows", type=int, default=3152)
ap.add_argument("--tiles", type=int, default=17)
a = ap.parse_args()
d, h, E = a.d, 4 * a.d, a.E
dev, bf = "cuda", torch.bfloat16
seg_len = (a.rows + 255) // 256 * 256
rows_cap = seg_len * E + 256
seg = torch.arange(E + 1, dtype=torch.int32, device=dev) * seg_len
tile_e = torch.full((rows_cap // 256,), -1, dtype=torch.int32, device=dev)
tile_e[: seg_len * E // 256] = torch.arange(E, device=dev, dtype=torch.int32).repeat_interleave(seg_len // 256)
nm = torch.tensor([seg_len * E // 256], dtype=torch.int32, device=dev)
rnd = lambda *s: (torch.randn(*s, device=dev) * 0.5).to(bf)
X, Hh, U, dY, dU = rnd(rows_cap, d), rnd(rows_cap, h), rnd(rows_cap, h), rnd(rows_cap, d), rnd(rows_cap, h)
W1, W2 = rnd(E, h, d), rnd(E, d, h)
b1, b2 = torch.randn(E, h, device=dev), torch.randn(E, d, device=dev)
oU, oH, oY, odU, odX = (torch.empty(rows_cap, n, dtype=bf, device=dev) for n in (h, h, d, h, d))
dW1, dW2 = torch.empty(E, h, d, device=dev), torch.empty(E, d, h, device=dev)
st, P = C.stream_ptr(), C.ptr
ops = {
    "fc1": lambda: C.call("moe_grouped_gemm", C.GEMM_FC1, P(X), P(W1), P(oU), P(oH), P(b1), None, P(tile_e), P(nm), None, rows_cap, E, 0, h, d, st),
    "fc2": lambda: C.call("moe_grouped_gemm", C.GEMM_FC2, P(Hh), P(W2), P(oY), None, P(b2), None, P(tile_e), P(nm), None, rows_cap, E, 0, d, h, st),
    "dgelu": lambda: C.call("moe_grouped_gemm", C.GEMM_DGELU, P(dY), P(W2), P(odU), None, None, P(U), P(tile_e), P(nm), None, rows_cap, E, 0, h, d, st),
    "dgrad": lambda: C.call("moe_grouped_gemm", C.GEMM_DGRAD, P(dU), P(W1), P(odX), None, None, None, P(tile_e), P(nm), None, rows_cap, E, 0, d, h, st),
    "wgrad1": lambda: C.call("moe_grouped_gemm", C.GEMM_WGRAD, P(dU), P(X), P(dW1), None, None, P(C.wgrad_flags(E, h, d, dev)), None, None, P(seg), rows_cap, E, h, d, 0, st),
    "wgrad2": lambda: C.call("moe_grouped_gemm", C.GEMM_WGRAD_T, P(Hh), P(dY), P(dW2), None, None, P(C.wgrad_flags(E, h, d, dev)), None, None, P(seg), rows_cap, E, h, d, 0, st),
}
for _ in range(3):
    ops[a.op]()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops[a.op](); e1.record(); torch.cuda.synchronize()
buf = (ctypes.c_longlong * (2 * 4 * 64 * 4))()
fn = C.lib.moe_debug_timeline   # same CDLL handle the calls above went through (the stamps live in its device globals)
assert fn(buf) == 0
t = torch.tensor(list(buf), dtype=torch.int64).view(2, 4, 64, 4)
print(f"{a.op}: {e0.elapsed_time(e1) * 1e3:.1f} us")
for cta in (0, 1):
    base = int(t[cta, 0, 0, 0]) if cta == 0 else int(t[cta, 0, 0, 0])
    print(f"--- CTA {cta} (clocks since the producer's first stamp)")
    print("tile | producer: start slot0free lastTMA | mma: wait_tempty got_tempty first_full commit | epi w2: start tfull released done | epi w6: start tfull released done")
    for ti in range(a.tiles):
        r = lambda role: " ".join(f"{int(t[cta, role, ti, ev]) - base:7d}" if int(t[cta, role, ti, ev]) else "      -" for ev in range(4 if role else 3))
        print(f"{ti:4d} | {r(0)} | {r(1)} | {r(2)} | {r(3)}")
