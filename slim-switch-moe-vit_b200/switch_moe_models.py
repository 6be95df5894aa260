"""`@register_model` factories for the north-star configs (BASELINE.json configs[0..3]), written the way the
reference writes its own (/root/reference/models/resMoE.py:190-209: build the dense DeiT / ViT, then swap
`module.mlp` of its `Block`s) so that the UNCHANGED /root/reference/main.py selects them with `--model <name>`
(`create_model(args.model, ...)`, main.py:520-530).

Use: copy (or symlink) this file into the reference's `models/` directory and add `from .switch_moe_models import *`
to `models/__init__.py` — or simply `import switch_moe_models` before `create_model` with the reference root and this
directory on `sys.path`.  It needs the reference's `models` package (for the dense backbones and `Block`), timm's
registry, and the B200 `fmoe` drop-in; none of them is imported until a factory is called or `register()` runs.

| factory                                   | BASELINE config | backbone (file:line in the reference)              | experts | gate          | MoE blocks |
|-------------------------------------------|-----------------|----------------------------------------------------|---------|---------------|------------|
| switch_moe_tiny_patch16_224_e8_top1       | configs[0]      | deit_tiny_patch16_224  (models/model.py:80-100)      | 8       | Switch, cf 1.25 | 6 of 12  |
| switch_moe_small_patch16_224_e16_top1     | configs[1]      | deit_small_patch16_224 (models/model.py:140-160)     | 16      | Switch, cf 1.25 | 6 of 12  |
| gshard_moe_base_patch16_224_e32_top2      | configs[2]      | deit_base_patch16_224  (models/model.py:163-183)     | 32      | GShard, cf 1.25 | 6 of 12  |
| switch_moe_large_patch16_224_e64_top1     | configs[3]      | VisionTransformer 1024/24/16 (vision_transformer.py:1216-1224) | 64 | Switch, cf 1.25 | 12 of 24 |

Every factory takes `expert_parallel=None` (None: shard the experts over the default process group when
torch.distributed is initialised with more than one rank — main.py:460 initialises it before the model is built —
else keep all experts local), `capacity_factor`, `moe_stride`, and swallows the reference's
`starting_threshold` / `target_threshold` kwargs (main.py:528-529 passes them to every model).
"""
from __future__ import annotations

from functools import partial

__all__ = ["switch_moe_tiny_patch16_224_e8_top1", "switch_moe_small_patch16_224_e16_top1",
           "gshard_moe_base_patch16_224_e32_top2", "switch_moe_large_patch16_224_e64_top1", "CONFIGS"]

# name -> (embed_dim, num_experts, top_k, gate kind)
CONFIGS = {
    "switch_moe_tiny_patch16_224_e8_top1": (192, 8, 1, "switch"),
    "switch_moe_small_patch16_224_e16_top1": (384, 16, 1, "switch"),
    "gshard_moe_base_patch16_224_e32_top2": (768, 32, 2, "gshard"),
    "switch_moe_large_patch16_224_e64_top1": (1024, 64, 1, "switch"),
}


def _reference():
    """The reference's own modules (`models` package of d0-rb/slim-switch-moe-vit), whichever way this file was installed."""
    try:
        from . import model as ref_model, vision_transformer as ref_vit      # copied into the reference's models/
    except ImportError:
        from models import model as ref_model, vision_transformer as ref_vit  # imported from outside, reference root on sys.path
    return ref_model, ref_vit


def _backbone(name: str, pretrained: bool, kwargs: dict):
    import torch.nn as nn
    ref_model, ref_vit = _reference()
    if name == "large":   # the DeiT file stops at Base; same constructor pattern as models/model.py:163-183 at ViT-L/16 size
        model = ref_vit.VisionTransformer(patch_size=16, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4, qkv_bias=True,
                                          norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
        model.default_cfg = ref_model._cfg()
        return model, ref_vit.Block
    fn = {"tiny": ref_model.deit_tiny_patch16_224, "small": ref_model.deit_small_patch16_224,
          "base": ref_model.deit_base_patch16_224}[name]
    return fn(pretrained=pretrained, **kwargs), ref_vit.Block


def _build(size: str, factory_name: str, pretrained: bool, expert_parallel, capacity_factor: float, moe_stride: int, kwargs: dict):
    import fmoe
    kwargs.pop("starting_threshold", None)   # token-skip Gate arguments main.py passes to every model (main.py:528-529)
    kwargs.pop("target_threshold", None)
    dim, E, k, gate = CONFIGS[factory_name]
    model, Block = _backbone(size, pretrained, kwargs)
    if expert_parallel is None:
        import torch.distributed as dist
        expert_parallel = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    fmoe.install_switch_moe(model, Block, dim, num_experts=E, top_k=k, gate=gate, capacity_factor=capacity_factor,
                            moe_stride=moe_stride, expert_parallel=bool(expert_parallel))
    return model


def switch_moe_tiny_patch16_224_e8_top1(pretrained=False, expert_parallel=None, capacity_factor=1.25, moe_stride=2, **kwargs):
    return _build("tiny", "switch_moe_tiny_patch16_224_e8_top1", pretrained, expert_parallel, capacity_factor, moe_stride, kwargs)


def switch_moe_small_patch16_224_e16_top1(pretrained=False, expert_parallel=None, capacity_factor=1.25, moe_stride=2, **kwargs):
    return _build("small", "switch_moe_small_patch16_224_e16_top1", pretrained, expert_parallel, capacity_factor, moe_stride, kwargs)


def gshard_moe_base_patch16_224_e32_top2(pretrained=False, expert_parallel=None, capacity_factor=1.25, moe_stride=2, **kwargs):
    return _build("base", "gshard_moe_base_patch16_224_e32_top2", pretrained, expert_parallel, capacity_factor, moe_stride, kwargs)


def switch_moe_large_patch16_224_e64_top1(pretrained=False, expert_parallel=None, capacity_factor=1.25, moe_stride=2, **kwargs):
    return _build("large", "switch_moe_large_patch16_224_e64_top1", pretrained, expert_parallel, capacity_factor, moe_stride, kwargs)


def register():
    """Put the four factories into timm's model registry (what `@register_model` does at import time in the reference)."""
    from timm.models import register_model
    for name in CONFIGS:
        register_model(globals()[name])


try:   # same effect as the reference's decorators when timm is importable; silent otherwise (GPU box without timm)
    register()
except ImportError:
    pass
