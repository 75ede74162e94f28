"""Shared test helpers: build matching (oracle dims, product config) pairs and the error-envelope comparison."""
import numpy as np
import torch

from oracle import openvla_oracle as O
from openvla_probe_b200 import config as C


def pair(kind="tiny", fused=True, llm_layers=2, depth=(3, 3)):
    """(oracle VLADims, product OpenVLAConfig) describing the same architecture."""
    if kind == "tiny":
        od = O.tiny_dims(fused=fused, llm_layers=llm_layers, depth=depth)
        pc = C.tiny(fused=fused, llm_layers=llm_layers, depth=depth)
    elif kind == "full-width":
        # real layer widths (1024 / 1152 / 4304 / 4096 / 11008, 224 px, T = 288) with shallow depth: exercises the
        # full-size kernel shapes while the CPU oracle still finishes in seconds
        import dataclasses
        dino = O.TowerDims(1024, 3, 16, 4096, 5, True)
        sig = O.TowerDims(1152, 3, 16, 4304, 0, False)
        od = dataclasses.replace(O.OPENVLA_7B, towers=(dino, sig), llm_layers=llm_layers)
        base = C.openvla_7b()
        pc = dataclasses.replace(
            base, towers=(dataclasses.replace(C.DINOV2_L14_REG4, depth=3), dataclasses.replace(C.SIGLIP_SO400M_14, depth=3)),
            text_config=dataclasses.replace(base.text_config, num_hidden_layers=llm_layers))
    elif kind == "openvla-7b":
        od, pc = O.OPENVLA_7B, C.openvla_7b()
    elif kind == "siglip-7b":
        od, pc = O.SIGLIP_7B, C.siglip_7b()
    else:
        raise ValueError(kind)
    assert od.n_patches == pc.n_patches and od.vision_dim == pc.vision_dim
    assert [(t.dim, t.depth, t.heads, t.mlp, t.n_prefix, t.layerscale) for t in od.towers] == \
           [(t.dim, t.depth, t.heads, t.mlp, t.n_prefix, t.layerscale) for t in pc.towers]
    tc = pc.text_config
    assert (od.llm_dim, od.llm_inter, od.llm_layers, od.llm_heads, od.vocab) == \
           (tc.hidden_size, tc.intermediate_size, tc.num_hidden_layers, tc.num_attention_heads, tc.vocab_size)
    return od, pc


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_abs(a, b):
    return float((torch.as_tensor(a).double() - torch.as_tensor(b).double()).abs().max())


def envelope_ok(ours, ref32, ref_bf16, c=2.0, floor=1e-3):
    """Error-envelope policy: our bf16 result may deviate from the fp32 oracle at most `c`
    times as much as the oracle's own bf16 evaluation (the reference's numerics) does, in rel-L2 and max-abs."""
    e_ours, e_ref = rel_l2(ours, ref32), rel_l2(ref_bf16, ref32)
    m_ours, m_ref = max_abs(ours, ref32), max_abs(ref_bf16, ref32)
    scale = float(torch.as_tensor(ref32).double().abs().max()) + 1e-30
    ok = e_ours <= c * e_ref + floor and m_ours <= c * m_ref + floor * scale
    return ok, dict(rel_ours=e_ours, rel_ref=e_ref, max_ours=m_ours, max_ref=m_ref)


def to_f32(W):
    return {k: v.float() for k, v in W.items()}
