"""HF state-dict schema of the path and random initialisation on the device.

Names follow `vla-scripts/extern/convert_openvla_weights_to_hf.py:73-115` of the reference (`projector.fc{1,2,3}`,
`language_model.*`, `vision_backbone.featurizer.*` = DINOv2 with LayerScale stored as `.scale_factor`,
`vision_backbone.fused_featurizer.*` = SigLIP).  There is no network for checkpoints, so benchmarks use the
random-init architecture (`_init_weights`, modeling_prismatic.py:185-205: N(0, 0.02)).
"""
from __future__ import annotations

from typing import Dict, Iterator, Tuple

import torch

from .config import OpenVLAConfig

TOWER_PREFIX = ("vision_backbone.featurizer", "vision_backbone.fused_featurizer")


def state_dict_shapes(c: OpenVLAConfig) -> Dict[str, Tuple[int, ...]]:
    s: Dict[str, Tuple[int, ...]] = {}
    for ti, t in enumerate(c.towers):
        p = TOWER_PREFIX[ti]
        s[f"{p}.patch_embed.proj.weight"] = (t.dim, 3, c.patch, c.patch)
        s[f"{p}.patch_embed.proj.bias"] = (t.dim,)
        s[f"{p}.pos_embed"] = (1, c.n_patches, t.dim)
        if t.n_prefix:
            s[f"{p}.cls_token"] = (1, 1, t.dim)
            s[f"{p}.reg_token"] = (1, t.n_prefix - 1, t.dim)
        for i in range(t.depth):
            b = f"{p}.blocks.{i}"
            s[f"{b}.norm1.weight"] = (t.dim,)
            s[f"{b}.norm1.bias"] = (t.dim,)
            s[f"{b}.attn.qkv.weight"] = (3 * t.dim, t.dim)
            s[f"{b}.attn.qkv.bias"] = (3 * t.dim,)
            s[f"{b}.attn.proj.weight"] = (t.dim, t.dim)
            s[f"{b}.attn.proj.bias"] = (t.dim,)
            s[f"{b}.norm2.weight"] = (t.dim,)
            s[f"{b}.norm2.bias"] = (t.dim,)
            s[f"{b}.mlp.fc1.weight"] = (t.mlp, t.dim)
            s[f"{b}.mlp.fc1.bias"] = (t.mlp,)
            s[f"{b}.mlp.fc2.weight"] = (t.dim, t.mlp)
            s[f"{b}.mlp.fc2.bias"] = (t.dim,)
            if t.layerscale:
                s[f"{b}.ls1.scale_factor"] = (t.dim,)
                s[f"{b}.ls2.scale_factor"] = (t.dim,)
    vd, ld = c.vision_dim, c.text_config.hidden_size
    if c.use_fused_vision_backbone:
        s["projector.fc1.weight"] = (4 * vd, vd); s["projector.fc1.bias"] = (4 * vd,)
        s["projector.fc2.weight"] = (ld, 4 * vd); s["projector.fc2.bias"] = (ld,)
        s["projector.fc3.weight"] = (ld, ld); s["projector.fc3.bias"] = (ld,)
    else:
        s["projector.fc1.weight"] = (ld, vd); s["projector.fc1.bias"] = (ld,)
        s["projector.fc2.weight"] = (ld, ld); s["projector.fc2.bias"] = (ld,)
    tc = c.text_config
    lm = "language_model.model"
    s[f"{lm}.embed_tokens.weight"] = (tc.vocab_size, ld)
    for i in range(tc.num_hidden_layers):
        b = f"{lm}.layers.{i}"
        s[f"{b}.input_layernorm.weight"] = (ld,)
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            s[f"{b}.self_attn.{n}.weight"] = (ld, ld)
        s[f"{b}.post_attention_layernorm.weight"] = (ld,)
        s[f"{b}.mlp.gate_proj.weight"] = (tc.intermediate_size, ld)
        s[f"{b}.mlp.up_proj.weight"] = (tc.intermediate_size, ld)
        s[f"{b}.mlp.down_proj.weight"] = (ld, tc.intermediate_size)
    s[f"{lm}.norm.weight"] = (ld,)
    s["language_model.lm_head.weight"] = (tc.vocab_size, ld)
    return s


def _is_scale_like(name: str) -> bool:
    return name.endswith(("norm1.weight", "norm2.weight", "layernorm.weight", "model.norm.weight", "scale_factor"))


def random_tensors(c: OpenVLAConfig, device, seed: int = 0) -> Iterator[Tuple[str, torch.Tensor]]:
    """Yield (name, bf16 tensor on `device`) one at a time (the 7B model never exists twice in memory)."""
    g = torch.Generator(device=device).manual_seed(seed)
    for name, shape in state_dict_shapes(c).items():
        w = torch.randn(shape, generator=g, device=device, dtype=torch.float32)
        w = (1.0 + 0.1 * w) if _is_scale_like(name) else 0.02 * w
        if name.endswith("embed_tokens.weight"):
            w[c.pad_token_id] = 0.0
        yield name, w.to(torch.bfloat16)


def bind_random(model, seed: int = 0) -> None:
    """Random-init a model's engine in place on its device."""
    lm_head = None
    for name, w in random_tensors(model.config, model.device, seed):
        model.engine.bind(name, w)
        if name == "language_model.lm_head.weight":
            lm_head = w
    model.engine.finalize()
    model._lm_head_dev = lm_head
