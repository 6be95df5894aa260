"""`fmoe` — B200-native drop-in for the subset of FastMoE that d0-rb/slim-switch-moe-vit uses
(`from fmoe import FMoETransformerMLP`, /root/reference/models/resMoE.py:6).  Hot path:
hand-written sm_100a kernels behind the C ABI in include/moe_b200.h (libmoe_b200.so)."""
from . import _cabi  # noqa: F401  (fails loudly if libmoe_b200.so is not built)
from .dense import DenseFFN  # noqa: F401
from .distributed import DistributedGroupedDataParallel  # noqa: F401
from .fused import AddLayerNorm, Linear, add_layer_norm, fast_linear  # noqa: F401
from .gates import BaseGate, GShardGate, NaiveGate, SwitchGate  # noqa: F401
from .integration import MoEAuxCriterion, build_moe_mlp, install_switch_moe, load_balance_stats, log_load_balance, make_gate  # noqa: F401
from .layers import FMoE  # noqa: F401
from .linear import FMoELinear  # noqa: F401
from .transformer import FMoETransformerMLP  # noqa: F401

__all__ = ["DistributedGroupedDataParallel", "AddLayerNorm", "add_layer_norm", "DenseFFN", "FMoE", "FMoELinear", "FMoETransformerMLP", "BaseGate", "NaiveGate", "SwitchGate", "GShardGate"]
__version__ = "0.1.0+b200"
