"""GPU micro-benchmark of the six grouped-GEMM launches of one MoE layer (through the C ABI).

    python tools/gemm_bench.py [--d 384] [--E 16] [--rows 3152] [--iters 20] [--flush] [--ops fc1,fc2,...]

Every expert gets `rows` live rows (segments padded to 256).  Reports ms and TFLOP/s per op (flops
counted on LIVE rows only: 2*R*d*h), and torch.bmm (cuBLAS) on the same per-expert shapes as the
library yardstick.  --flush writes a 512 MB buffer between iterations (cold L2)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "slim-switch-moe-vit_b200"))
import torch  # noqa: E402

from fmoe import _cabi as C  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--d", type=int, default=384)
    ap.add_argument("--E", type=int, default=16)
    ap.add_argument("--rows", type=int, default=3152)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--flush", action="store_true")
    ap.add_argument("--ops", default="fc1,fc2,dgelu,dgrad,wgrad1,wgrad2")
    ap.add_argument("--no-cublas", action="store_true")
    a = ap.parse_args()
    d, h, E = a.d, 4 * a.d, a.E
    dev, bf = "cuda", torch.bfloat16
    seg_len = (a.rows + 255) // 256 * 256
    rows_cap = seg_len * E + 256
    R = a.rows * E
    seg = torch.arange(E + 1, dtype=torch.int32, device=dev) * seg_len
    tile_e = torch.full((rows_cap // 256,), -1, dtype=torch.int32, device=dev)
    tile_e[: seg_len * E // 256] = torch.arange(E, device=dev, dtype=torch.int32).repeat_interleave(seg_len // 256)
    nm = torch.tensor([seg_len * E // 256], dtype=torch.int32, device=dev)
    live = (torch.arange(rows_cap, device=dev) % seg_len < a.rows) & (torch.arange(rows_cap, device=dev) < seg_len * E)
    torch.manual_seed(0)

    def rnd(*s):
        return (torch.randn(*s, device=dev) * 0.5).to(bf)

    X, Hh, U, dY, dU = rnd(rows_cap, d), rnd(rows_cap, h), rnd(rows_cap, h), rnd(rows_cap, d), rnd(rows_cap, h)
    X[~live] = 0
    dY[~live] = 0
    W1, W2 = rnd(E, h, d), rnd(E, d, h)
    b1, b2 = torch.randn(E, h, device=dev), torch.randn(E, d, device=dev)
    oU, oH, oY, odU, odX = (torch.empty(rows_cap, n, dtype=bf, device=dev) for n in (h, h, d, h, d))
    dW1, dW2 = torch.empty(E, h, d, device=dev), torch.empty(E, d, h, device=dev)
    st = C.stream_ptr()
    P = C.ptr
    FL = C.wgrad_flags(E, h, d, dev)
    ops = {
        "fc1": lambda: C.call("moe_grouped_gemm", C.GEMM_FC1, P(X), P(W1), P(oU), P(oH), P(b1), None, P(tile_e), P(nm), None, rows_cap, E, 0, h, d, st),
        "fc2": lambda: C.call("moe_grouped_gemm", C.GEMM_FC2, P(Hh), P(W2), P(oY), None, P(b2), None, P(tile_e), P(nm), None, rows_cap, E, 0, d, h, st),
        "dgelu": lambda: C.call("moe_grouped_gemm", C.GEMM_DGELU, P(dY), P(W2), P(odU), None, None, P(U), P(tile_e), P(nm), None, rows_cap, E, 0, h, d, st),
        "dgrad": lambda: C.call("moe_grouped_gemm", C.GEMM_DGRAD, P(dU), P(W1), P(odX), None, None, None, P(tile_e), P(nm), None, rows_cap, E, 0, d, h, st),
        "wgrad1": lambda: C.call("moe_grouped_gemm", C.GEMM_WGRAD, P(dU), P(X), P(dW1), None, None, P(FL), None, None, P(seg), rows_cap, E, h, d, 0, st),
        "wgrad2": lambda: C.call("moe_grouped_gemm", C.GEMM_WGRAD_T, P(Hh), P(dY), P(dW2), None, None, P(FL), None, None, P(seg), rows_cap, E, h, d, 0, st),
        "wgrad2_mn": lambda: C.call("moe_grouped_gemm", C.GEMM_WGRAD, P(dY), P(Hh), P(dW2), None, None, None, None, None, P(seg), rows_cap, E, d, h, 0, st),
    }
    Xe, He = X[: seg_len * E].view(E, seg_len, d), Hh[: seg_len * E].view(E, seg_len, h)
    cublas = {
        "fc1": lambda: torch.bmm(Xe, W1.transpose(1, 2)),
        "fc2": lambda: torch.bmm(He, W2.transpose(1, 2)),
        "dgelu": lambda: torch.bmm(Xe, W2),
        "dgrad": lambda: torch.bmm(He, W1),
        "wgrad1": lambda: torch.bmm(He.transpose(1, 2), Xe),
        "wgrad2": lambda: torch.bmm(Xe.transpose(1, 2), He),
    }
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev) if a.flush else None
    flops = 2.0 * R * d * h

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(a.iters):
            if flush_buf is not None:
                flush_buf.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / a.iters

    print(f"d={d} h={h} E={E} rows/expert={a.rows} (segment {seg_len}) live rows R={R} flops/op={flops / 1e9:.1f} GF flush={a.flush}")
    total = 0.0
    for name in a.ops.split(","):
        ms = timeit(ops[name])
        total += ms
        line = f"{name:7s} {ms * 1e3:8.1f} us  {flops / ms / 1e9:7.1f} TFLOP/s"
        if not a.no_cublas and name in cublas:
            cms = timeit(cublas[name])
            line += f"   | cuBLAS bmm (padded rows, no epilogue) {cms * 1e3:8.1f} us {2.0 * seg_len * E * d * h / cms / 1e9:7.1f} TFLOP/s"
        print(line, flush=True)
    n = len(a.ops.split(","))
    print(f"all {n} ops: {total * 1e3:.1f} us -> {n * flops / total / 1e9:.1f} TFLOP/s aggregate")


if __name__ == "__main__":
    main()
