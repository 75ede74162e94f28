"""Minimal stand-in for the `timm` package (absent in this image, not installable offline) so that the REFERENCE's
own `prismatic/extern/hf/modeling_prismatic.py` can be imported and executed to produce golden vectors.

This is test tooling written for this repo, not reference code: an nn.Module restatement of the timm 0.9.10
`VisionTransformer` surface that modeling_prismatic.py touches (`create_model`, `.blocks`, `.embed_dim`,
`.get_intermediate_layers(x, n={...})`, `LayerScale` with `.gamma` / `.inplace`).  Tower sizes are taken from
`TOWERS`, which the golden script sets before building the reference model (tiny sizes for fixtures).
"""
from __future__ import annotations

import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

__version__ = "0.9.10"

# timm id -> dict(dim, depth, heads, mlp, n_prefix, layerscale, patch); filled in by make_golden.py
TOWERS: dict = {}


class LayerScale(nn.Module):
    def __init__(self, dim, init_values=1e-5, inplace=False):
        super().__init__()
        self.inplace = inplace
        self.gamma = nn.Parameter(init_values * torch.ones(dim))

    def forward(self, x):
        return x.mul_(self.gamma) if self.inplace else x * self.gamma


class Attention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads, self.head_dim = heads, dim // heads
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        x = F.scaled_dot_product_attention(q, k, v)
        return self.proj(x.transpose(1, 2).reshape(B, N, C))


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1, self.act, self.fc2 = nn.Linear(dim, hidden), nn.GELU(), nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim, heads, mlp, layerscale):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, heads)
        self.ls1 = LayerScale(dim) if layerscale else nn.Identity()
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, mlp)
        self.ls2 = LayerScale(dim) if layerscale else nn.Identity()

    def forward(self, x):
        x = x + self.ls1(self.attn(self.norm1(x)))
        return x + self.ls2(self.mlp(self.norm2(x)))


class PatchEmbed(nn.Module):
    def __init__(self, patch, dim):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch, bias=True)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class VisionTransformer(nn.Module):
    def __init__(self, img_size, patch, dim, depth, heads, mlp, n_prefix, layerscale):
        super().__init__()
        self.embed_dim = dim
        self.num_prefix_tokens = n_prefix
        self.patch_embed = PatchEmbed(patch, dim)
        n = (img_size // patch) ** 2
        self.pos_embed = nn.Parameter(torch.randn(1, n, dim) * 0.02)
        if n_prefix:
            self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
            self.reg_token = nn.Parameter(torch.zeros(1, n_prefix - 1, dim))
        self.blocks = nn.Sequential(*[Block(dim, heads, mlp, layerscale) for _ in range(depth)])

    def _pos_embed(self, x):
        x = x + self.pos_embed                      # no_embed_class=True (DINOv2 reg4) / no prefix (SigLIP)
        if self.num_prefix_tokens:
            B = x.shape[0]
            x = torch.cat([self.cls_token.expand(B, -1, -1), self.reg_token.expand(B, -1, -1), x], dim=1)
        return x

    def get_intermediate_layers(self, x, n=1, reshape=False, return_prefix_tokens=False, norm=False):
        take = set(n) if not isinstance(n, int) else set(range(len(self.blocks) - n, len(self.blocks)))
        x = self._pos_embed(self.patch_embed(x))
        outs = []
        for i, blk in enumerate(self.blocks):       # timm runs every block, keeps the requested ones
            x = blk(x)
            if i in take:
                outs.append(x)
        return tuple(o[:, self.num_prefix_tokens:] for o in outs)


def create_model(name, pretrained=False, num_classes=0, img_size=224, act_layer=None, **kw):
    t = TOWERS[name]
    return VisionTransformer(img_size, t["patch"], t["dim"], t["depth"], t["heads"], t["mlp"], t["n_prefix"],
                             t["layerscale"])


def install() -> None:
    """Register this module as `timm` (+ `timm.models.vision_transformer`) in sys.modules."""
    me = sys.modules[__name__]
    models = types.ModuleType("timm.models")
    vt = types.ModuleType("timm.models.vision_transformer")
    vt.LayerScale = LayerScale
    vt.VisionTransformer = VisionTransformer
    models.vision_transformer = vt
    me.models = models
    sys.modules["timm"] = me
    sys.modules["timm.models"] = models
    sys.modules["timm.models.vision_transformer"] = vt
