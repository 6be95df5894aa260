"""Fits the polynomials used by the GEMM epilogues for exact-erf GELU and its derivative.

  gelu(u)  = relu(u) - |u| * Q(|u|),          Q(a) = Phi(-a) = 0.5 erfc(a / sqrt 2) = exp2(P(a))   [rounds 1b-1d: deg 6]
                                                                                   = g(a) * W(a)   [shipped: W deg 7, shared with gelu']
  gelu'(u) = u < 0 ? m(|u|) : 1 - m(|u|),      m(a) = Phi(-a) - a phi(a)         = g(a) * (W(a) - a / sqrt(2 pi))   [shipped]
  g(a) = exp(-a^2 / 2)  (one ex2.approx),  P, W and w polynomials in a on [0, A_MAX] (a is clamped).

Weighted least squares on Chebyshev nodes (weight = the factor that multiplies the polynomial in the
final result), then the fp32 Horner evaluation is checked against float64 erfc on a dense grid.
"""
import numpy as np
from numpy.polynomial import chebyshev as Ch, polynomial as Pl
from scipy.special import erfc, erfcx

A_MAX = 6.5   # beyond this g(a) <= 6.7e-10 and both corrections are < 1e-9 in absolute value


def fit(fn, weight, deg, n=4001):
    x = np.cos(np.pi * (np.arange(n) + 0.5) / n)          # Chebyshev nodes on [-1, 1]
    a = (x + 1) * 0.5 * A_MAX
    w = weight(a)
    c = Ch.chebfit(x, fn(a), deg, w=w)
    # convert to monomials in a
    px = Ch.cheb2poly(c)                                    # in x
    # x = 2a/A_MAX - 1
    lin = np.array([-1.0, 2.0 / A_MAX])
    out = np.zeros(1)
    for k, ck in enumerate(px):
        out = Pl.polyadd(out, ck * Pl.polypow(lin, k))
    return out


def horner32(coef, a):
    a = a.astype(np.float32)
    r = np.full_like(a, np.float32(coef[-1]))
    for c in coef[-2::-1]:
        r = (r * a + np.float32(c)).astype(np.float32)     # not fused, slightly pessimistic
    return r


def main():
    g = lambda a: np.exp(-0.5 * a * a)
    W = lambda a: 0.5 * erfcx(a / np.sqrt(2.0))
    w = lambda a: 0.5 * erfcx(a / np.sqrt(2.0)) - a / np.sqrt(2 * np.pi)
    a = np.linspace(0, A_MAX, 2_000_001)
    for deg in (8, 9, 10, 11, 12):
        cW = fit(W, lambda t: np.maximum(t, 0.05) * g(t), deg)
        cw = fit(w, lambda t: g(t), deg)
        g32 = np.exp2((-(a.astype(np.float32) ** 2) * np.float32(0.5 * np.log2(np.e))).astype(np.float32)).astype(np.float32)
        Q = g32 * horner32(cW, a)
        m = g32 * horner32(cw, a)
        errQ = np.abs(a * (Q - 0.5 * erfc(a / np.sqrt(2))))
        errm = np.abs(m - (0.5 * erfc(a / np.sqrt(2)) - a * g(a) / np.sqrt(2 * np.pi)))
        print(f"deg {deg}: max |gelu err| = {errQ.max():.3e}   max |gelu' err| = {errm.max():.3e}")
        if deg in (9, 10, 11):
            print("  W:", ", ".join(f"{c:.9e}f" for c in cW))
            print("  w:", ", ".join(f"{c:.9e}f" for c in cw))


def main_shipped_v2():
    """Round 1e: gelu and gelu' share g = exp(-a^2/2) and one degree-7 polynomial W(a) = 0.5 erfcx(a / sqrt 2)."""
    a = np.linspace(0, A_MAX, 2_000_001)
    g = lambda t: np.exp(-0.5 * t * t)
    W = lambda t: 0.5 * erfcx(t / np.sqrt(2.0))
    c0 = 1 / np.sqrt(2 * np.pi)
    cW = fit(W, lambda t: np.maximum(t, 0.5) * g(t), 7)
    g32 = np.exp2((-(a.astype(np.float32) ** 2) * np.float32(0.5 * np.log2(np.e))).astype(np.float32)).astype(np.float32)
    Wv = horner32(cW, a)
    errQ = np.abs(a * (g32 * Wv - 0.5 * erfc(a / np.sqrt(2))))
    errm = np.abs(g32 * (Wv - np.float32(c0) * a.astype(np.float32)) - (0.5 * erfc(a / np.sqrt(2)) - a * g(a) * c0))
    print(f"W7 : max |gelu err| = {errQ.max():.3e}   max |gelu' err| = {errm.max():.3e}")
    print("  kGeluW:", ", ".join(f"{c:.9e}f" for c in cW))


def main_shipped():
    from scipy.special import log_ndtr
    a = np.linspace(0, A_MAX, 2_000_001)
    Q = lambda t: 0.5 * erfc(t / np.sqrt(2.0))
    P = lambda t: log_ndtr(-t) / np.log(2.0)
    g = lambda t: np.exp(-0.5 * t * t)
    w = lambda t: 0.5 * erfcx(t / np.sqrt(2.0)) - t / np.sqrt(2 * np.pi)
    cP = fit(P, lambda t: np.maximum(t, 0.02) * Q(t), 6)
    q = np.exp2(horner32(cP, a).astype(np.float64))
    print(f"P6 : max |gelu err| = {np.abs(a * (q - Q(a))).max():.3e}")
    print("  kGeluLogQ:", ", ".join(f"{c:.9e}f" for c in cP))
    cw = fit(w, lambda t: g(t), 8)
    g32 = np.exp2((-(a.astype(np.float32) ** 2) * np.float32(0.5 * np.log2(np.e))).astype(np.float32)).astype(np.float64)
    m = g32 * horner32(cw, a)
    print(f"w8 : max |gelu' err| = {np.abs(m - (Q(a) - a * g(a) / np.sqrt(2 * np.pi))).max():.3e}")
    print("  kGeluGradW:", ", ".join(f"{c:.9e}f" for c in cw))


if __name__ == "__main__":
    main_shipped_v2()
