from .registry import create_model, register_model  # noqa: F401
