"""Minimal stand-ins for the few timm symbols /root/reference/models imports (timm is not installed
in this image).  TEST INFRASTRUCTURE: used only by tests/golden/make_golden.py and the CPU test
that imports the unmodified reference models against our `fmoe` drop-in.  Semantics follow
timm 0.5 (the version the reference's vision_transformer.py was forked from)."""
from .models.registry import create_model  # noqa: F401
