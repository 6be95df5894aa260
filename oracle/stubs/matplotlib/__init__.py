"""Empty stand-in: /root/reference/models/base_model.py:7 imports matplotlib.pyplot at module scope."""
