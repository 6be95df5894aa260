"""Build recipe for the C part of the oracle (test infrastructure, never shipped).

`ensure_built()` compiles oracle/gate_ref.c with gcc into oracle/_build/libmoe_oracle.so.
There is no `oracle/_ref` for this repository: the reference's arithmetic for the hot path
lives in FastMoE (a CUDA-only, un-vendored third-party package), so there are no reference
sources under /root/reference that could be compiled on a CPU — see DESIGN.md "Oracle".
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "gate_ref.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_OUT = os.path.join(_OUT_DIR, "libmoe_oracle.so")

# -mfma: fmaf() must be a single fused instruction (fast, and identical to the GPU's FFMA);
# -ffp-contract=off: nothing else may be fused behind our back.
_CFLAGS = ["-O2", "-fPIC", "-shared", "-std=c99", "-mfma", "-ffp-contract=off"]


def ensure_built(force: bool = False) -> str:
    os.makedirs(_OUT_DIR, exist_ok=True)
    stale = (not os.path.exists(_OUT)) or os.path.getmtime(_OUT) < os.path.getmtime(_SRC)
    if force or stale:
        cmd = ["gcc", *_CFLAGS, "-o", _OUT, _SRC, "-lm"]
        subprocess.run(cmd, check=True)
    return _OUT


_lib = None


def load():
    global _lib
    if _lib is None:
        path = ensure_built()
        lib = ctypes.CDLL(path)
        i64, i32, p = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p
        lib.moe_oracle_gate_logits.argtypes = [p, i64, i32, p, p, p, i32, p]
        lib.moe_oracle_gate_logits.restype = None
        lib.moe_oracle_route.argtypes = [p, i64, i32, i32, i32, i64, i32, p, p, p, p, p, p, p, p]
        lib.moe_oracle_route.restype = None
        _lib = lib
    return _lib


if __name__ == "__main__":
    print(ensure_built(force=True))
