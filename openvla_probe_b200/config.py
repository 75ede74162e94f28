"""Architecture constants of the path (what "openvla-7b" means numerically).

Mirrors the fields of the reference's `OpenVLAConfig` / `PrismaticConfig`
(prismatic/extern/hf/configuration_prismatic.py:72-140) that the hot path reads, plus the timm tower sizes that the
reference obtains from `timm.create_model` (configuration_prismatic.py:22-36) and the LlamaConfig defaults used as
`text_config` (configuration_prismatic.py:119-123; vocab patched to 32064 by convert_openvla_weights_to_hf.py:174-176).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple


@dataclass(frozen=True)
class TowerConfig:
    timm_id: str
    dim: int
    depth: int
    heads: int
    mlp: int
    n_prefix: int        # cls + register tokens that are prepended and later stripped
    layerscale: bool

    @property
    def head_dim(self) -> int:
        return self.dim // self.heads


DINOV2_L14_REG4 = TowerConfig("vit_large_patch14_reg4_dinov2.lvd142m", 1024, 24, 16, 4096, 5, True)
SIGLIP_SO400M_14 = TowerConfig("vit_so400m_patch14_siglip_224", 1152, 27, 16, 4304, 0, False)


@dataclass(frozen=True)
class TextConfig:
    """Subset of transformers.LlamaConfig read by the path."""
    hidden_size: int = 4096
    intermediate_size: int = 11008
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    vocab_size: int = 32064
    rms_norm_eps: float = 1e-6
    rope_theta: float = 10000.0
    max_position_embeddings: int = 2048
    pad_token_id: int = 32000
    bos_token_id: int = 1
    eos_token_id: int = 2


@dataclass(frozen=True)
class OpenVLAConfig:
    vision_backbone_id: str = "dinosiglip-vit-so-224px"
    llm_backbone_id: str = "llama2-7b-pure"
    arch_specifier: str = "no-align+fused-gelu-mlp"
    towers: Tuple[TowerConfig, ...] = (DINOV2_L14_REG4, SIGLIP_SO400M_14)
    image_size: int = 224
    patch: int = 14
    image_resize_strategy: str = "resize-naive"      # configuration_prismatic.py / openvla-7b's config.json
    text_config: TextConfig = TextConfig()
    pad_token_id: int = 32000
    pad_to_multiple_of: int = 64
    n_action_bins: int = 256
    norm_stats: Optional[Dict] = None
    output_hidden_states: bool = False
    use_return_dict: bool = True

    @property
    def use_fused_vision_backbone(self) -> bool:
        return len(self.towers) == 2

    @property
    def n_patches(self) -> int:
        return (self.image_size // self.patch) ** 2

    @property
    def vision_dim(self) -> int:
        return sum(t.dim for t in self.towers)

    @property
    def timm_model_ids(self):
        return [t.timm_id for t in self.towers]

    @property
    def image_sizes(self):
        return [self.image_size] * len(self.towers)


def openvla_7b(norm_stats: Optional[Dict] = None) -> OpenVLAConfig:
    return OpenVLAConfig(norm_stats=norm_stats)


def siglip_7b(norm_stats: Optional[Dict] = None) -> OpenVLAConfig:
    """prism-siglip-224px + Llama-2-7B single-backbone variant (prismatic/conf/models.py:174-176)."""
    return OpenVLAConfig(vision_backbone_id="siglip-vit-so400m", arch_specifier="no-align+gelu-mlp",
                         towers=(SIGLIP_SO400M_14,), norm_stats=norm_stats)


def tiny(fused: bool = True, llm_layers: int = 2, depth=(3, 3), norm_stats: Optional[Dict] = None) -> OpenVLAConfig:
    """Structure-preserving small config for tests (same head dims 64 / 72 / 128, prefix tokens, LayerScale)."""
    dino = TowerConfig("tiny_dinov2_reg4", 128, depth[0], 2, 512, 5, True)
    sig = TowerConfig("tiny_siglip", 144, depth[1], 2, 536, 0, False)
    return OpenVLAConfig(
        towers=(dino, sig) if fused else (sig,), image_size=56,
        text_config=TextConfig(hidden_size=256, intermediate_size=704, num_hidden_layers=llm_layers,
                               num_attention_heads=2),
        norm_stats=norm_stats,
        vision_backbone_id="dinosiglip-vit-so-224px" if fused else "siglip-vit-so400m",
        arch_specifier="no-align+fused-gelu-mlp" if fused else "no-align+gelu-mlp",
    )
