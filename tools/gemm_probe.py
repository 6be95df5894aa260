"""GPU probe: run ONE grouped-GEMM case through the C ABI and compare with torch.matmul.
usage: python tools/gemm_probe.py <op> <E> <rows_per_expert_csv|N> <M> <N> <K>
Each case runs in its own process (a device trap must not poison the next case)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "slim-switch-moe-vit_b200"))
import torch  # noqa: E402

from fmoe import _cabi as C  # noqa: E402


def segs(counts):
    seg = [0]
    for c in counts:
        seg.append(seg[-1] + (c + 255) // 256 * 256)
    return seg


def main():
    op = int(sys.argv[1]); E = int(sys.argv[2])
    counts = [int(v) for v in sys.argv[3].split(",")]
    if len(counts) == 1:
        counts = counts * E
    M, N, K = int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
    dev = "cuda"
    torch.manual_seed(0)
    seg = segs(counts)
    rows = seg[-1]
    rows_cap = rows + 256
    seg_t = torch.tensor(seg, dtype=torch.int32, device=dev)
    tile_e = torch.full((rows_cap // 256,), -1, dtype=torch.int32)
    for e in range(E):
        tile_e[seg[e] // 256: seg[e + 1] // 256] = e
    tile_e = tile_e.to(dev)
    nm = torch.tensor([rows // 256], dtype=torch.int32, device=dev)
    st = C.stream_ptr()
    bf = torch.bfloat16

    def rnd(*s):
        return (torch.randn(*s, device=dev) * 0.5).to(bf)

    def live_mask():
        m = torch.zeros(rows_cap, dtype=torch.bool, device=dev)
        for e in range(E):
            m[seg[e]: seg[e] + counts[e]] = True
        return m

    if op in (C.GEMM_FC1, C.GEMM_FC2):
        A = rnd(rows_cap, K); B = rnd(E, N, K); bias = torch.randn(E, N, device=dev)
        o0 = torch.zeros(rows_cap, N, dtype=bf, device=dev); o1 = torch.zeros_like(o0)
        C.call("moe_grouped_gemm", op, C.ptr(A), C.ptr(B), C.ptr(o0), C.ptr(o1) if op == C.GEMM_FC1 else None,
               C.ptr(bias), None, C.ptr(tile_e), C.ptr(nm), None, rows_cap, E, 0, N, K, st)
        torch.cuda.synchronize()
        ref = torch.zeros(rows_cap, N, device=dev)
        for e in range(E):
            ref[seg[e]:seg[e + 1]] = A[seg[e]:seg[e + 1]].float() @ B[e].float().t() + bias[e]
        want0 = ref
        if op == C.GEMM_FC1:   # out0 = gelu'(U), out1 = gelu(U)
            want0 = 0.5 * (1 + torch.erf(ref / 2 ** 0.5)) + ref * torch.exp(-0.5 * ref * ref) / (2 * 3.141592653589793) ** 0.5
        err0 = (o0[:rows].float() - want0[:rows]).abs().max().item()
        print(f"op={op} out0 max_abs_err={err0:.4g} ref_max={want0[:rows].abs().max().item():.4g}")
        ok = err0 < 0.02 * max(1.0, want0.abs().max().item())
        if op == C.GEMM_FC1:
            g = torch.nn.functional.gelu(ref)
            err1 = (o1[:rows].float() - g[:rows]).abs().max().item()
            print(f"      out1 (gelu) max_abs_err={err1:.4g}")
            ok = ok and err1 < 0.02 * max(1.0, g.abs().max().item())
        tail = o0[rows:].float().abs().max().item()
        print(f"      rows beyond live range untouched: {tail == 0.0}")
        ok = ok and tail == 0.0
    elif op in (C.GEMM_DGELU, C.GEMM_DGRAD):
        A = rnd(rows_cap, K); B = rnd(E, K, N); aux = rnd(rows_cap, N)   # B = the forward weight [E, K, N], read MN-major
        o0 = torch.zeros(rows_cap, N, dtype=bf, device=dev)
        C.call("moe_grouped_gemm", op, C.ptr(A), C.ptr(B), C.ptr(o0), None, None,
               C.ptr(aux) if op == C.GEMM_DGELU else None, C.ptr(tile_e), C.ptr(nm), None, rows_cap, E, 0, N, K, st)
        torch.cuda.synchronize()
        ref = torch.zeros(rows_cap, N, device=dev)
        for e in range(E):
            ref[seg[e]:seg[e + 1]] = A[seg[e]:seg[e + 1]].float() @ B[e].float()
        if op == C.GEMM_DGELU:   # out0 = (A B) * aux
            ref = ref * aux.float()
        err0 = (o0[:rows].float() - ref[:rows]).abs().max().item()
        print(f"op={op} out0 max_abs_err={err0:.4g} ref_max={ref[:rows].abs().max().item():.4g}")
        ok = err0 < 0.02 * max(1.0, ref.abs().max().item())
    else:
        A = rnd(rows_cap, M); B = rnd(rows_cap, N)
        lm = live_mask()
        A[~lm] = 0  # contract: pad rows are zero in at least one operand
        tr = op == C.GEMM_WGRAD_T    # out [E, N, M] = (A_e^T B_e)^T
        o0 = torch.full((E, N, M) if tr else (E, M, N), 7.0, device=dev)
        fl = C.wgrad_flags(E, M, N, dev)
        for _ in range(2):   # twice: the second launch finds the flags the first one left behind
            C.call("moe_grouped_gemm", op, C.ptr(A), C.ptr(B), C.ptr(o0), None, None, C.ptr(fl), None, None, C.ptr(seg_t),
                   rows_cap, E, M, N, 0, st)
        torch.cuda.synchronize()
        print(f"      split-K flags left clear: {int(fl.abs().sum()) == 0}")
        ref = torch.zeros(E, M, N, device=dev)
        for e in range(E):
            ref[e] = A[seg[e]:seg[e + 1]].float().t() @ B[seg[e]:seg[e + 1]].float()
        if tr:
            o0 = o0.transpose(1, 2)
        err0 = (o0 - ref).abs().max().item()
        print(f"op={op} wgrad max_abs_err={err0:.4g} ref_max={ref.abs().max().item():.4g}")
        ok = err0 < 1e-3 * max(1.0, ref.abs().max().item())
    print("PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
