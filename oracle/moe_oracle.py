"""oracle/moe_oracle.py — CPU restatement of the Switch-style MoE layer (FastMoE
`FMoETransformerMLP` as the reference wraps it in /root/reference/models/resMoE.py:15-29).

TEST INFRASTRUCTURE ONLY — never imported by the product path.  PARITY UNPINNED: FastMoE is
absent from /root/reference and from this image, and the reference ships no golden vectors;
the semantics restated here are FastMoE's published ones, with the two
non-deterministic upstream choices made canonical (see oracle/gate_ref.c header).

Three independent restatements live here, each checked against the others in tests/:

* `forward_model` / `backward_model`  — the *arithmetic model* of the CUDA path: same routing
  (bit-exact, via oracle/gate_ref.c), and bf16 rounding at exactly the points where the
  kernels round (dispatch buffer, weights, activation derivative G = gelu'(U), hidden H, expert output Y and
  the backward buffers).  CUDA results are compared with this under a tight tolerance.
* `ideal_forward` — plain differentiable PyTorch (fp32/fp64) of the same function with the
  routing held fixed: `F.linear -> F.gelu(approximate='none') -> F.linear`, index_add combine.
  fp64 autograd through it is the gradient reference.
* `bruteforce_forward` — a per-token Python loop that evaluates the selected experts densely;
  it shares no indexing code with the other two and cross-checks them on small cases.
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

from . import build as _build

ALIGN = 256  # rows; every expert segment of the packed buffers starts on a multiple of this (one CTA-pair tile)

SCORE_TOPK_SOFTMAX = 0  # NaiveGate / GShardGate: softmax over the k selected logits
SCORE_FULL_SOFTMAX = 1  # SwitchGate: probability of the selected expert under softmax over all E


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest-even to bf16, returned widened to fp32 (what cvt.rn.bf16.f32 does)."""
    return t.to(torch.float32).to(torch.bfloat16).to(torch.float32)


def capacity_from_factor(cf: float, T: int, k: int, E: int) -> int:
    """C = ceil(cf * T * k / E)  (Switch/GShard convention).  cf <= 0 => no limit."""
    if cf is None or cf <= 0:
        return T * k
    return min(T * k, int(math.ceil(cf * T * k / E)))


# --------------------------------------------------------------------------------------
# gate + routing (bit-exact restatement; C)
# --------------------------------------------------------------------------------------
def gate_logits(x: torch.Tensor, Wg: torch.Tensor, bg: torch.Tensor | None,
                noise: torch.Tensor | None = None) -> torch.Tensor:
    """logits[T,E] in LOGIT ORDER v1 (oracle/gate_ref.c).  x may be fp32 or bf16."""
    lib = _build.load()
    xf = np.ascontiguousarray(x.detach().to(torch.float32).cpu().numpy())
    w = np.ascontiguousarray(Wg.detach().to(torch.float32).cpu().numpy())
    T, d = xf.shape
    E = w.shape[0]
    b = None if bg is None else np.ascontiguousarray(bg.detach().to(torch.float32).cpu().numpy())
    nz = None if noise is None else np.ascontiguousarray(noise.detach().to(torch.float32).cpu().numpy())
    out = np.empty((T, E), dtype=np.float32)
    lib.moe_oracle_gate_logits(_ptr(xf), T, d, _ptr(w), None if b is None else _ptr(b),
                               None if nz is None else _ptr(nz), E, _ptr(out))
    return torch.from_numpy(out)


@dataclass
class Routing:
    idx: torch.Tensor        # int32 [T,k]
    score: torch.Tensor      # fp32  [T,k]
    count: torch.Tensor      # int32 [E]   pairs routed to each expert (before capacity)
    kept: torch.Tensor       # int32 [E]   min(count, C)
    seg_start: torch.Tensor  # int32 [E+1] first row of each expert's segment (ALIGN-aligned)
    pos: torch.Tensor        # int32 [T,k] row in the packed buffer, -1 = dropped
    psum: torch.Tensor       # fp64  [E]   sum_t softmax(logits[t])[e]
    capacity: int

    @property
    def rows(self) -> int:
        return int(self.seg_start[-1])


def route(logits: torch.Tensor, k: int, score_mode: int, capacity: int, align: int = ALIGN, token_mask=None) -> Routing:
    """token_mask [T] (optional, non-zero = keep): skipped tokens get idx = pos = -1, score = 0 and are not counted."""
    lib = _build.load()
    tm = None if token_mask is None else np.ascontiguousarray((token_mask.detach().reshape(-1).cpu() != 0).numpy().astype(np.uint8))
    lg = np.ascontiguousarray(logits.detach().to(torch.float32).cpu().numpy())
    T, E = lg.shape
    idx = np.empty((T, k), np.int32)
    score = np.empty((T, k), np.float32)
    count = np.empty(E, np.int32)
    kept = np.empty(E, np.int32)
    seg = np.empty(E + 1, np.int32)
    pos = np.empty((T, k), np.int32)
    psum = np.empty(E, np.float64)
    lib.moe_oracle_route(_ptr(lg), T, E, k, score_mode, int(capacity), align, _ptr(idx), _ptr(score),
                         _ptr(count), _ptr(kept), _ptr(seg), _ptr(pos), _ptr(psum), None if tm is None else _ptr(tm))
    tn = torch.from_numpy
    return Routing(tn(idx), tn(score), tn(count), tn(kept), tn(seg), tn(pos), tn(psum), int(capacity))


def route_python(logits: torch.Tensor, k: int, score_mode: int, capacity: int, align: int = ALIGN, token_mask=None) -> Routing:
    """Independent pure-Python restatement of `route` (small cases only) — pins the C code."""
    lg = logits.detach().to(torch.float32)
    T, E = lg.shape
    idx = torch.zeros(T, k, dtype=torch.int32)
    score = torch.zeros(T, k, dtype=torch.float32)
    live = [True] * T if token_mask is None else [bool(v) for v in (token_mask.reshape(-1) != 0).tolist()]
    for t in range(T):
        if not live[t]:
            idx[t] = -1
            continue
        row = lg[t].tolist()
        order = sorted(range(E), key=lambda e: (-row[e], e))[:k]  # desc value, ties -> low index
        idx[t] = torch.tensor(order, dtype=torch.int32)
        if score_mode == SCORE_TOPK_SOFTMAX:
            score[t] = torch.softmax(lg[t, order], dim=-1)
        else:
            score[t] = torch.softmax(lg[t], dim=-1)[order]
    flat = idx.reshape(-1).tolist()
    seen = [0] * E
    rank = []
    for e in flat:
        if e < 0:
            rank.append(capacity)   # skipped token: never kept
            continue
        rank.append(seen[e])
        seen[e] += 1
    count = torch.tensor(seen, dtype=torch.int32)
    kept = torch.clamp(count, max=capacity).to(torch.int32)
    seg = [0]
    for e in range(E):
        seg.append(seg[-1] + (int(kept[e]) + align - 1) // align * align)
    pos = torch.tensor([seg[e] + r if (e >= 0 and r < capacity) else -1 for e, r in zip(flat, rank)], dtype=torch.int32)
    psum = (torch.softmax(lg.double(), dim=-1) * torch.tensor(live, dtype=torch.float64).unsqueeze(1)).sum(0)
    return Routing(idx, score, count, kept, torch.tensor(seg, dtype=torch.int32), pos.reshape(T, k), psum, capacity)


# --------------------------------------------------------------------------------------
# arithmetic model of the CUDA path
# --------------------------------------------------------------------------------------
def gelu_erf(u: torch.Tensor) -> torch.Tensor:
    return F.gelu(u, approximate="none")


def gelu_erf_grad(u: torch.Tensor) -> torch.Tensor:
    """d/du [u * Phi(u)] = Phi(u) + u * phi(u)."""
    cdf = 0.5 * (1.0 + torch.erf(u * (1.0 / math.sqrt(2.0))))
    pdf = torch.exp(-0.5 * u * u) * (1.0 / math.sqrt(2.0 * math.pi))
    return cdf + u * pdf


@dataclass
class Saved:
    x: torch.Tensor
    logits: torch.Tensor
    r: Routing
    row_src: torch.Tensor   # int64 [rows] flattened pair index t*k+j of each buffer row, -1 = pad
    row_exp: torch.Tensor   # int64 [rows] expert owning the row (pads included)
    Xb: torch.Tensor
    Gb: torch.Tensor        # bf16(gelu'(U)) of the fp32 pre-activation U (U itself is never stored)
    Hb: torch.Tensor
    Yb: torch.Tensor
    W1b: torch.Tensor
    W2b: torch.Tensor
    score_mode: int
    k: int


def _row_tables(r: Routing, E: int):
    rows = r.rows
    row_src = torch.full((rows,), -1, dtype=torch.int64)
    flat_pos = r.pos.reshape(-1).to(torch.int64)
    valid = flat_pos >= 0
    row_src[flat_pos[valid]] = torch.nonzero(valid, as_tuple=False).reshape(-1)
    row_exp = torch.zeros(rows, dtype=torch.int64)
    for e in range(E):
        row_exp[int(r.seg_start[e]):int(r.seg_start[e + 1])] = e
    return row_src, row_exp


def _grouped_linear(X, W, b, row_exp, seg_start):
    """Y[r] = X[r] @ W[e(r)]^T + b[e(r)], fp32 accumulation, per expert segment (pads included)."""
    E = W.shape[0]
    out = torch.empty(X.shape[0], W.shape[1], dtype=torch.float32)
    for e in range(E):
        s, t = int(seg_start[e]), int(seg_start[e + 1])
        if t > s:
            out[s:t] = X[s:t] @ W[e].t() + (b[e] if b is not None else 0.0)
    return out


def forward_model(x, Wg, bg, W1, b1, W2, b2, k: int, score_mode: int, capacity: int,
                  routing: Routing | None = None, logits: torch.Tensor | None = None):
    """Arithmetic model of the CUDA forward.  Returns (y in x.dtype, Saved)."""
    T, d = x.shape
    E = W1.shape[0]
    xf = x.detach().to(torch.float32)
    if logits is None:
        logits = gate_logits(x, Wg, bg)
    r = routing if routing is not None else route(logits, k, score_mode, capacity)
    row_src, row_exp = _row_tables(r, E)
    rows = r.rows
    Xb = torch.zeros(rows, d, dtype=torch.float32)
    valid = row_src >= 0
    Xb[valid] = bf16_round(xf[row_src[valid] // k])
    W1b, W2b = bf16_round(W1.detach()), bf16_round(W2.detach())
    U = _grouped_linear(Xb, W1b, b1.detach().float(), row_exp, r.seg_start)
    Gb = bf16_round(gelu_erf_grad(U))
    Hb = bf16_round(gelu_erf(U))
    Y = _grouped_linear(Hb, W2b, b2.detach().float(), row_exp, r.seg_start)
    Yb = bf16_round(Y)
    y = torch.zeros(T, d, dtype=torch.float32)
    for j in range(k):
        p = r.pos[:, j].to(torch.int64)
        m = p >= 0
        y[m] += r.score[m, j].unsqueeze(1) * Yb[p[m]]
    saved = Saved(xf, logits, r, row_src, row_exp, Xb, Gb, Hb, Yb, W1b, W2b, score_mode, k)
    return y.to(x.dtype), saved


def gate_backward(logits, r: Routing, dscore, dpsum, score_mode: int):
    """dlogits[T,E] (fp32 math in fp64 here; compared under tolerance)."""
    lg = logits.double()
    T, E = lg.shape
    k = r.idx.shape[1]
    idx = r.idx.to(torch.int64)
    dl = torch.zeros(T, E, dtype=torch.float64)
    g = dscore.double()
    s = r.score.double()
    if score_mode == SCORE_TOPK_SOFTMAX:
        inner = (s * g).sum(1, keepdim=True)
        dl.scatter_add_(1, idx, s * (g - inner))
    else:
        p = torch.softmax(lg, dim=-1)
        for j in range(k):
            pj = p.gather(1, idx[:, j:j + 1])            # [T,1]
            onehot = torch.zeros(T, E, dtype=torch.float64).scatter_(1, idx[:, j:j + 1], 1.0)
            dl += g[:, j:j + 1] * pj * (onehot - p)
    if dpsum is not None:
        p = torch.softmax(lg, dim=-1)
        dp = dpsum.double().unsqueeze(0)
        dl += p * (dp - (p * dp).sum(1, keepdim=True))
    return dl.float()


def backward_model(sv: Saved, dy, Wg, dpsum=None):
    """Arithmetic model of the CUDA backward.  Returns dict of grads (fp32; dx in fp32)."""
    r, k = sv.r, sv.k
    T, d = sv.x.shape
    E = sv.W1b.shape[0]
    dyf = dy.detach().to(torch.float32)
    rows = r.rows
    dYb = torch.zeros(rows, d, dtype=torch.float32)
    dscore = torch.zeros(T, k, dtype=torch.float32)
    for j in range(k):
        p = r.pos[:, j].to(torch.int64)
        m = p >= 0
        dYb[p[m]] = bf16_round(r.score[m, j].unsqueeze(1) * dyf[m])
        dscore[m, j] = (dyf[m] * sv.Yb[p[m]]).sum(1)
    dH = torch.empty(rows, sv.W2b.shape[2], dtype=torch.float32)
    dW1 = torch.zeros_like(sv.W1b)
    dW2 = torch.zeros_like(sv.W2b)
    db1 = torch.zeros(E, sv.W1b.shape[1])
    db2 = torch.zeros(E, d)
    for e in range(E):
        s, t = int(r.seg_start[e]), int(r.seg_start[e + 1])
        dH[s:t] = dYb[s:t] @ sv.W2b[e]
    dUb = bf16_round(dH * sv.Gb)
    dXb = torch.empty(rows, d, dtype=torch.float32)
    for e in range(E):
        s, t = int(r.seg_start[e]), int(r.seg_start[e + 1])
        dW2[e] = dYb[s:t].t() @ sv.Hb[s:t]
        db2[e] = dYb[s:t].sum(0)
        dW1[e] = dUb[s:t].t() @ sv.Xb[s:t]
        db1[e] = dUb[s:t].sum(0)
        dXb[s:t] = dUb[s:t] @ sv.W1b[e]
    dXb = bf16_round(dXb)
    dlogits = gate_backward(sv.logits, r, dscore, dpsum, sv.score_mode)
    dx = dlogits @ Wg.detach().float()
    for j in range(k):
        p = r.pos[:, j].to(torch.int64)
        m = p >= 0
        dx[m] += dXb[p[m]]
    dWg = dlogits.t() @ sv.x
    dbg = dlogits.sum(0)
    return dict(dx=dx, dWg=dWg, dbg=dbg, dW1=dW1, db1=db1, dW2=dW2, db2=db2,
                dscore=dscore, dlogits=dlogits, dYb=dYb, dUb=dUb, dXb=dXb)


# --------------------------------------------------------------------------------------
# ideal (un-rounded, differentiable) restatement and brute force
# --------------------------------------------------------------------------------------
def ideal_forward(x, Wg, bg, W1, b1, W2, b2, r: Routing, score_mode: int, dtype=torch.float64):
    """Differentiable plain-PyTorch layer with the routing integers held fixed.
    Returns (y, psum) so aux losses built from psum are differentiable too."""
    c = lambda t: t.to(dtype)
    x, Wg, W1, b1, W2, b2 = map(c, (x, Wg, W1, b1, W2, b2))
    T, d = x.shape
    E = W1.shape[0]
    k = r.idx.shape[1]
    logits = F.linear(x, Wg, None if bg is None else c(bg))
    idx = r.idx.to(torch.int64)
    if score_mode == SCORE_TOPK_SOFTMAX:
        score = torch.softmax(logits.gather(1, idx), dim=-1)
    else:
        score = torch.softmax(logits, dim=-1).gather(1, idx)
    psum = torch.softmax(logits, dim=-1).sum(0)
    y = torch.zeros(T, d, dtype=dtype)
    for j in range(k):
        keep = r.pos[:, j] >= 0
        for e in range(E):
            m = keep & (idx[:, j] == e)
            if m.any():
                h = F.gelu(F.linear(x[m], W1[e], b1[e]), approximate="none")
                y = y.index_add(0, torch.nonzero(m).reshape(-1), score[m, j].unsqueeze(1) * F.linear(h, W2[e], b2[e]))
    return y, psum


def bruteforce_forward(x, Wg, bg, W1, b1, W2, b2, k: int, score_mode: int, capacity: int):
    """Per-token loop, fp64, independent of every indexing helper above (small cases only)."""
    x64 = x.double()
    T, d = x.shape
    E = W1.shape[0]
    logits = gate_logits(x, Wg, bg).double()   # same logits => same picks; the rest is independent
    y = torch.zeros(T, d, dtype=torch.float64)
    used = [0] * E
    for t in range(T):
        row = logits[t].tolist()
        order = sorted(range(E), key=lambda e: (-row[e], e))[:k]
        if score_mode == SCORE_TOPK_SOFTMAX:
            sc = torch.softmax(logits[t, order], dim=-1)
        else:
            sc = torch.softmax(logits[t], dim=-1)[order]
        for j, e in enumerate(order):
            keep = used[e] < capacity
            used[e] += 1
            if not keep:
                continue
            u = W1[e].double() @ x64[t] + b1[e].double()
            h = 0.5 * u * (1.0 + torch.erf(u / math.sqrt(2.0)))
            y[t] += sc[j] * (W2[e].double() @ h + b2[e].double())
    return y


AUX_NONE, AUX_SWITCH, AUX_GSHARD = 0, 1, 2


def aux_coef(r: Routing, T: int, aux_mode: int) -> torch.Tensor:
    """d aux_loss / d psum [E] (fp64): aux_loss = sum_e coef_e * psum_e, coef_e = E / T * share_e with
    share_e = kept_e / sum(kept) (Switch) or count_e / (T k) (GShard) — SURVEY.md §8a."""
    E = r.count.shape[0]
    k = r.idx.shape[1]
    if aux_mode == AUX_SWITCH:
        kept = r.kept.double()
        share = kept / kept.sum().clamp(min=1)
    elif aux_mode == AUX_GSHARD:
        share = r.count.double() / float(T * k)
    else:
        return torch.zeros(E, dtype=torch.float64)
    return share * E / T


def switch_aux_loss(r: Routing, psum: torch.Tensor, T: int) -> torch.Tensor:
    """E * sum_e f_e * P_e with f_e = kept_e / sum(kept), P_e = psum_e / T."""
    E = psum.shape[0]
    kept = r.kept.to(psum.dtype)
    f = kept / kept.sum().clamp(min=1)
    return E * (f * (psum / T)).sum()
