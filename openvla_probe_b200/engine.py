"""Thin ctypes wrapper over the native OvlaEngine (include/ovla_b200.h).  PyTorch is used only for device memory,
streams and pinned host buffers; all arithmetic happens inside libovla_b200.so."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib
from .config import OpenVLAConfig


class _Tower(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("dim", "depth", "heads", "mlp", "n_prefix", "layerscale")]


class _Dims(C.Structure):
    _fields_ = [
        ("image_size", C.c_int), ("patch", C.c_int), ("n_towers", C.c_int), ("towers", _Tower * 2),
        ("llm_dim", C.c_int), ("llm_inter", C.c_int), ("llm_layers", C.c_int), ("llm_heads", C.c_int),
        ("vocab", C.c_int), ("rms_eps", C.c_float), ("max_batch", C.c_int), ("max_seq", C.c_int),
    ]


class _RunArgs(C.Structure):
    _fields_ = [
        ("input_ids_dev", C.c_void_p), ("pixel_values_dev", C.c_void_p),
        ("B", C.c_int), ("P", C.c_int), ("pool_len", C.c_int), ("pool_mode", C.c_int), ("n_new_tokens", C.c_int),
        ("pooled_out_dev", C.c_void_p), ("tokens_out_dev", C.c_void_p), ("step_logits_out_dev", C.c_void_p),
        ("hidden_out_dev", C.c_void_p), ("projector_out_dev", C.c_void_p), ("patches_out_dev", C.c_void_p),
        ("prompt_lens_dev", C.c_void_p),
    ]


def rope_tables(head_dim: int, theta: float, max_seq: int):
    """LlamaRotaryEmbedding (transformers modeling_llama.py): inv_freq and angles in fp32, cos/sin cast to bf16.
    Only the first half of `emb = cat(freqs, freqs)` is stored (both halves are identical)."""
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    fr = torch.arange(max_seq, dtype=torch.float32).view(-1, 1) * inv.view(1, -1)
    return fr.cos().to(torch.bfloat16).contiguous(), fr.sin().to(torch.bfloat16).contiguous()


class Engine:
    """One engine per GPU: packed weights + workspace for `max_batch` observations of up to `max_seq` positions."""

    def __init__(self, config: OpenVLAConfig, max_batch: int, max_prompt_len: int = 40, max_new_tokens: int = 8,
                 device: int = 0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.OvlaError("libovla_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.config = config
        self.device = torch.device("cuda", device)
        self.max_batch = max_batch
        tc = config.text_config
        self.max_seq = config.n_patches + max_prompt_len + max_new_tokens
        d = _Dims()
        d.image_size, d.patch, d.n_towers = config.image_size, config.patch, len(config.towers)
        for i, t in enumerate(config.towers):
            d.towers[i] = _Tower(t.dim, t.depth, t.heads, t.mlp, t.n_prefix, int(t.layerscale))
        d.llm_dim, d.llm_inter, d.llm_layers = tc.hidden_size, tc.intermediate_size, tc.num_hidden_layers
        d.llm_heads, d.vocab, d.rms_eps = tc.num_attention_heads, tc.vocab_size, tc.rms_norm_eps
        d.max_batch, d.max_seq = max_batch, self.max_seq
        self._h = C.c_void_p()
        _lib.check(self.lib.ovla_create(C.byref(d), device, C.byref(self._h)))
        self.lib.ovla_workspace_bytes.restype = C.c_longlong
        self.lib.ovla_weight_bytes.restype = C.c_longlong
        hd = tc.hidden_size // tc.num_attention_heads
        cos, sin = rope_tables(hd, tc.rope_theta, self.max_seq)
        self.bind("rope.cos", cos.to(self.device))
        self.bind("rope.sin", sin.to(self.device))
        self._finalized = False

    # ------------------------------------------------------------------ weights
    def bind(self, name: str, tensor: torch.Tensor) -> None:
        t = tensor.detach()
        if t.dtype != torch.bfloat16:
            t = t.to(torch.bfloat16)
        t = t.to(self.device).contiguous()
        shape = (C.c_longlong * t.dim())(*t.shape)
        _lib.check(self.lib.ovla_bind_weight(self._h, name.encode(), C.c_void_p(t.data_ptr()), shape, t.dim()))

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor]) -> None:
        for k, v in state_dict.items():
            self.bind(k, v)
        self.finalize()

    def finalize(self) -> None:
        _lib.check(self.lib.ovla_finalize(self._h))
        self._finalized = True

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.ovla_workspace_bytes(self._h))

    @property
    def weight_bytes(self) -> int:
        return int(self.lib.ovla_weight_bytes(self._h))

    # ------------------------------------------------------------------ run
    def run(self, input_ids: torch.Tensor, pixel_values: torch.Tensor, pool_len: int = 0, pool_mode: int = 0,
            n_new_tokens: int = 0, want_hidden: bool = False, want_logits: bool = False, want_projector: bool = False,
            want_patches: bool = False, prompt_lens: Optional[torch.Tensor] = None):
        """Device-resident entry (ovla_run).  Returns a dict of device tensors.  `prompt_lens` (int32 [B], optional):
        true lengths of right-padded rows of `input_ids` (OvlaRunArgs.prompt_lens_dev)."""
        if not self._finalized:
            raise _lib.OvlaError("engine weights are not bound (call load_state_dict / finalize)")
        cfg, tc = self.config, self.config.text_config
        B, P = input_ids.shape
        T = cfg.n_patches + P
        ids = input_ids.to(self.device, torch.int64).contiguous()
        px = pixel_values.to(self.device, torch.bfloat16).contiguous()
        L, D, V = tc.num_hidden_layers, tc.hidden_size, tc.vocab_size
        out = {}
        a = _RunArgs()
        a.input_ids_dev, a.pixel_values_dev = ids.data_ptr(), px.data_ptr()
        a.B, a.P, a.pool_len, a.pool_mode, a.n_new_tokens = B, P, pool_len, pool_mode, n_new_tokens
        if prompt_lens is not None:
            if prompt_lens.numel() != B:
                raise ValueError("prompt_lens must hold one length per row")
            lens = prompt_lens.to(self.device, torch.int32).contiguous()
            if B and (int(lens.min()) < 1 or int(lens.max()) > P):
                raise ValueError(f"prompt lengths must lie in [1, {P}]")
            a.prompt_lens_dev = lens.data_ptr()
        if pool_len > 0:
            out["pooled"] = torch.empty(L + 1, B, D, dtype=torch.float32, device=self.device)
            a.pooled_out_dev = out["pooled"].data_ptr()
        if n_new_tokens > 0:
            out["tokens"] = torch.empty(B, n_new_tokens, dtype=torch.int64, device=self.device)
            a.tokens_out_dev = out["tokens"].data_ptr()
            if want_logits:
                out["step_logits"] = torch.empty(n_new_tokens, B, V, dtype=torch.float32, device=self.device)
                a.step_logits_out_dev = out["step_logits"].data_ptr()
        if want_hidden:
            out["hidden"] = torch.empty(L + 1, B, T, D, dtype=torch.bfloat16, device=self.device)
            a.hidden_out_dev = out["hidden"].data_ptr()
        if want_projector:
            out["projector"] = torch.empty(B, cfg.n_patches, D, dtype=torch.bfloat16, device=self.device)
            a.projector_out_dev = out["projector"].data_ptr()
        if want_patches:
            out["patches"] = torch.empty(B, cfg.n_patches, cfg.vision_dim, dtype=torch.bfloat16, device=self.device)
            a.patches_out_dev = out["patches"].data_ptr()
        if B > 0:
            _lib.check(self.lib.ovla_run(self._h, C.byref(a), _lib.stream_ptr()))
        return out

    def run_host(self, input_ids_host: torch.Tensor, pixel_values_host: torch.Tensor, pool_len: int, pool_mode: int,
                 n_new_tokens: int, pooled_out_host: Optional[torch.Tensor], tokens_out_host: Optional[torch.Tensor],
                 prompt_lens_host: Optional[torch.Tensor] = None):
        """Host-buffer entry (ovla_run_host): H2D copies, fused pass, D2H copies, stream sync -- all inside the call."""
        if not self._finalized:
            raise _lib.OvlaError("engine weights are not bound (call load_state_dict / finalize)")
        B, P = input_ids_host.shape
        assert input_ids_host.dtype == torch.int64 and input_ids_host.is_contiguous() and not input_ids_host.is_cuda
        assert pixel_values_host.dtype == torch.bfloat16 and pixel_values_host.is_contiguous()
        if prompt_lens_host is not None:
            assert prompt_lens_host.dtype == torch.int32 and prompt_lens_host.is_contiguous() and prompt_lens_host.numel() == B
        _lib.check(self.lib.ovla_run_host(
            self._h, C.c_void_p(input_ids_host.data_ptr()), C.c_void_p(pixel_values_host.data_ptr()), B, P,
            pool_len, pool_mode, n_new_tokens,
            C.c_void_p(pooled_out_host.data_ptr()) if pooled_out_host is not None else None,
            C.c_void_p(tokens_out_host.data_ptr()) if tokens_out_host is not None else None,
            C.c_void_p(prompt_lens_host.data_ptr()) if prompt_lens_host is not None else None,
            _lib.stream_ptr()))

    def set_option(self, name: str, value: int) -> None:
        """ovla_set_option: "decode_mega", "attn_tc", "fuse_rope", "two_streams", "graph_max_batch"."""
        _lib.check(self.lib.ovla_set_option(self._h, name.encode(), int(value)))

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.ovla_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
