"""Host model for the MoE layer: a DeiT-style ViT whose MLPs are (some of them) Switch-MoE layers.

This is the *caller* of the hot path, not the hot path.  It restates, in stock PyTorch, what the
reference builds around the layer so that BASELINE.json's named configs can be run where
/root/reference and timm do not exist (the GPU box):

  * block structure      /root/reference/models/vision_transformer.py:283-322  (pre-norm attention +
                          pre-norm MLP, both residual; `self.mlp` is the attribute the factories swap)
  * model skeleton       /root/reference/models/vision_transformer.py:642-848  (conv patch embed, cls
                          token, learned pos-embed, blocks, final norm, linear head on the cls token)
  * the swap             /root/reference/models/resMoE.py:190-209  (`module.mlp = CustomizedMoEMLP(d, 4d,
                          moe_num_experts=E, moe_top_k=k, drop=0.0)` for every Block)
  * model sizes          /root/reference/models/vision_transformer.py:1084-1090 (Ti), :1124-1132 (S),
                          :1170-1178 (B), :1216-1224 (L)

The reference hard-codes E=8, k=2, every block (resMoE.py:194-207); `moe_stride`, `num_experts`,
`top_k`, `gate` and `capacity_factor` are the north-star extensions.  The MoE module
is injected through `moe_mlp` (default: the B200 `fmoe.FMoETransformerMLP`) so bench.py's CPU
baseline can build the same model around the CPU restatement; nothing here imports `oracle/`.
Attention and the dense MLP stay stock PyTorch and data-parallel; with the B200 layer the residual + LayerNorm
pairs run on fmoe.AddLayerNorm (and, opt-in, the dense MLP on fmoe.DenseFFN).
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

SIZES = {  # name -> (embed dim, depth, heads)
    "tiny": (192, 12, 3),
    "small": (384, 12, 6),
    "base": (768, 12, 12),
    "large": (1024, 24, 16),
}


@dataclass(frozen=True)
class MoEViTConfig:
    size: str = "small"
    num_experts: int = 16          # GLOBAL number of experts
    top_k: int = 1
    capacity_factor: float = 1.25  # <= 0: no capacity limit (what the reference's NaiveGate does)
    moe_stride: int = 2            # every `moe_stride`-th block is MoE (1 = every block, as in the reference)
    gate: str = "switch"           # "switch" (top-1, full softmax, capacity, aux loss) | "naive" | "gshard"
    num_classes: int = 1000
    img_size: int = 224
    patch: int = 16
    world_size: int = 1            # expert-parallel group size (experts sharded contiguously)

    @property
    def dims(self):
        return SIZES[self.size]

    def is_moe_block(self, i: int) -> bool:
        return self.moe_stride > 0 and i % self.moe_stride == self.moe_stride - 1

    def describe(self) -> str:
        d, depth, _ = self.dims
        n_moe = sum(self.is_moe_block(i) for i in range(depth))
        return (f"ViT-{self.size.capitalize()}/{self.patch} Switch-MoE E{self.num_experts} top-{self.top_k} "
                f"cf{self.capacity_factor:g} gate={self.gate} moe_blocks={n_moe}/{depth}")


class _SplitQKV(torch.autograd.Function):
    """q, k, v = qkv.unbind(2) as views, with ONE stacked gradient in backward.  (Plain `unbind` makes autograd
    build the gradient of the projection output from three zero-filled tensors and two adds per block —
    3.7 ms of a 20 ms training step at batch 256, profiles/r01b_step_launches.md.)"""

    @staticmethod
    def forward(ctx, qkv):                     # [B, N, 3, H, D]
        q, k, v = qkv.unbind(2)
        return q, k, v

    @staticmethod
    def backward(ctx, dq, dk, dv):
        return torch.stack((dq, dk, dv), dim=2)


class Attention(nn.Module):
    """Dense multi-head self-attention (reference models/vision_transformer.py:260-280); stock SDPA."""

    def __init__(self, dim, heads, linear=nn.Linear):
        super().__init__()
        self.heads = heads
        self.qkv = linear(dim, dim * 3)
        self.proj = linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        # q, k, v as strided [B, H, N, D] views of the projection output: no permute copies in either direction
        q, k, v = _SplitQKV.apply(self.qkv(x).view(B, N, 3, self.heads, C // self.heads))
        out = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
        return self.proj(out.transpose(1, 2).reshape(B, N, C))


class Mlp(nn.Module):
    def __init__(self, dim, hidden, linear=nn.Linear):
        super().__init__()
        self.fc1 = linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Block(nn.Module):
    """Pre-norm block of the reference (vision_transformer.py:319-322): x + attn(norm1(x)), then x + mlp(norm2(x))."""

    def __init__(self, dim, heads, mlp: nn.Module, norm=nn.LayerNorm, linear=nn.Linear):
        super().__init__()
        self.norm1 = norm(dim, eps=1e-6)
        self.attn = Attention(dim, heads, linear)
        self.norm2 = norm(dim, eps=1e-6)
        self.mlp = mlp

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))

    def forward_fused(self, x, delta):
        """Same arithmetic with every residual add folded into the LayerNorm that follows it
        (fmoe.AddLayerNorm).  `delta` is the previous sub-layer's output that has not been added to the
        fp32 residual stream `x` yet; returns (x, pending delta)."""
        if delta is None:
            n1 = self.norm1(x)
        else:
            x, n1 = self.norm1(x, delta)
        x, n2 = self.norm2(x, self.attn(n1))
        return x, self.mlp(n2)


def _b200_moe_mlp(cfg: MoEViTConfig, dim: int, hidden: int) -> nn.Module:
    from fmoe.integration import build_moe_mlp  # the B200 drop-in (fails loudly without libmoe_b200.so)
    return build_moe_mlp(dim, hidden, num_experts=cfg.num_experts, top_k=cfg.top_k, gate=cfg.gate,
                         capacity_factor=cfg.capacity_factor, world_size=cfg.world_size)


class MoEViT(nn.Module):
    def __init__(self, cfg: MoEViTConfig, moe_mlp=None, fused_norm: bool | None = None, dense_ffn: bool = False):
        """`moe_mlp(dim, hidden)` builds the MoE module (default: the B200 layer).  `fused_norm` selects the
        fused residual-add + LayerNorm kernels around the sub-layers (default: on with the B200 layer, off
        otherwise — the CPU baseline runs the stock block)."""
        super().__init__()
        self.cfg = cfg
        dim, depth, heads = cfg.dims
        self.fused_norm = (moe_mlp is None) if fused_norm is None else fused_norm
        norm, linear = nn.LayerNorm, nn.Linear
        if self.fused_norm:
            from fmoe import AddLayerNorm, DenseFFN, Linear
            norm, linear = AddLayerNorm, Linear     # same parameters; fused residual+LN and the fast bias-gradient backward
        moe_mlp = moe_mlp or partial(_b200_moe_mlp, cfg)
        # dense blocks: stock MLP by default.  `dense_ffn=True` runs them on the layer's grouped GEMM as a single expert
        # (fmoe.DenseFFN, same fc1 / fc2 parameters): measured at config 2 it ties with cuBLAS + the two GELU passes
        # (18.08 vs 17.90 ms per step, profiles/r01f_*), so it is off unless asked for.
        dense_mlp = DenseFFN if (dense_ffn and self.fused_norm and dim % 64 == 0) else partial(Mlp, linear=linear)
        n_patches = (cfg.img_size // cfg.patch) ** 2
        self.patch_embed = nn.Conv2d(3, dim, kernel_size=cfg.patch, stride=cfg.patch)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n_patches + 1, dim))
        self.blocks = nn.ModuleList(
            Block(dim, heads, moe_mlp(dim, 4 * dim) if cfg.is_moe_block(i) else dense_mlp(dim, 4 * dim), norm, linear)
            for i in range(depth))
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.head = nn.Linear(dim, cfg.num_classes)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)

    @property
    def moe_layers(self):
        return [b.mlp for i, b in enumerate(self.blocks) if self.cfg.is_moe_block(i)]

    def aux_loss(self):
        """Sum of the gates' load-balancing losses set during the last forward."""
        losses = [m.gate.get_loss() for m in self.moe_layers if m.gate.has_loss]
        return sum(l.sum() for l in losses) if losses else None

    def _patchify(self, img):
        """Patch embedding (reference: timm PatchEmbed = Conv2d(3, d, 16, stride 16)) evaluated as one GEMM over
        flattened patches — identical arithmetic and parameters, without the slow strided-convolution kernels."""
        pe, p = self.patch_embed, self.cfg.patch
        B, Cin, Hh, Ww = img.shape
        patches = img.view(B, Cin, Hh // p, p, Ww // p, p).permute(0, 2, 4, 1, 3, 5).reshape(B, (Hh // p) * (Ww // p), Cin * p * p)
        w = pe.weight.view(pe.weight.shape[0], -1)
        if self.fused_norm:
            from fmoe import fast_linear
            return fast_linear(patches, w, pe.bias)
        return F.linear(patches, w, pe.bias)

    def forward(self, img):
        x = self._patchify(img) if img.is_cuda else self.patch_embed(img).flatten(2).transpose(1, 2)
        x = torch.cat((self.cls_token.expand(x.shape[0], -1, -1), x), dim=1) + self.pos_embed
        if not self.fused_norm:
            for blk in self.blocks:
                x = blk(x)
            return self.head(self.norm(x)[:, 0])
        x, delta = x.float(), None
        for blk in self.blocks:
            x, delta = blk.forward_fused(x, delta)
        cls = x[:, 0] + delta[:, 0].float()       # only the class token reaches the head
        return self.head(F.layer_norm(cls, self.norm.normalized_shape, self.norm.weight, self.norm.bias, self.norm.eps))

    def train_flops_per_image(self, kept_fraction: float = 1.0) -> float:
        """fwd+bwd (3x forward) matmul flops per image, SURVEY.md §8d formula."""
        cfg = self.cfg
        d, depth, _ = cfg.dims
        n = (cfg.img_size // cfg.patch) ** 2 + 1
        per_tok = 0.0
        for i in range(depth):
            per_tok += 8 * d * d + 4 * n * d
            if cfg.is_moe_block(i):
                per_tok += 16 * d * d * cfg.top_k * kept_fraction + 2 * d * cfg.num_experts
            else:
                per_tok += 16 * d * d
        once = 2 * (3 * cfg.patch ** 2) * d * (n - 1) + 2 * d * cfg.num_classes
        return 3.0 * (n * per_tok + once)
