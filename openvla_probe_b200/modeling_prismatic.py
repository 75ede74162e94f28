"""Drop-in Python surface of the reference's HF-style model for the predict_action / capture path.

Mirrors `OpenVLAForActionPrediction(PrismaticForConditionalGeneration)` from the reference
(prismatic/extern/hf/modeling_prismatic.py:208-562): same method names, argument meaning, return types and error
behaviour, but every tensor op runs in libovla_b200.so (hand-written sm_100a kernels) through `Engine`.

Differences that are deliberate and documented in DESIGN.md:
  * `predict_action` accepts B >= 1 (the reference's generation path is batch-1 only, modeling_prismatic.py:326,460-463);
    row b of the result equals the reference's B == 1 result for observation b; shape (7,) for B == 1, (B, 7) otherwise.
  * Ragged batches: prompts of different lengths are passed RIGHT-padded with an `attention_mask` of ones followed by
    zeros (the tokenizer's `padding_side="right"`, processing_prismatic.py; the reference splices that mask at
    modeling_prismatic.py:388-390).  Row b then behaves as its own batch-1 call: 29871 goes right after its last real
    token, the capture pools over its true length and it decodes at its own positions.
  * `predict_action_and_capture` does the reference's two passes (`vla(**inputs, output_hidden_states=True)` then
    `vla.predict_action(**inputs)`, experiments/robot/openvla_utils.py:188-203) in ONE prefill: LLM attention is
    causal, so hidden states at positions [0, T-1) of the predict pass (which appends token 29871) are those of the
    capture pass.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .config import OpenVLAConfig
from .engine import Engine

EMPTY_TOKEN_ID = 29871   # '' token the reference appends after "Out:" (modeling_prismatic.py:512-515)


@dataclass
class PrismaticCausalLMOutputWithPast:
    """modeling_prismatic.py:162-173.  `past_key_values` stays inside the engine (KV cache is engine-owned)."""
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    past_key_values: Optional[Any] = None
    hidden_states: Optional[Tuple[torch.Tensor, ...]] = None
    attentions: Optional[Any] = None
    projector_features: Optional[torch.Tensor] = None


class OpenVLAForActionPrediction:
    config_class = OpenVLAConfig

    def __init__(self, config: OpenVLAConfig, max_batch: int = 1, max_prompt_len: int = 40, device: int = 0) -> None:
        self.config = config
        self.norm_stats = config.norm_stats                       # modeling_prismatic.py:497
        self.bins = np.linspace(-1, 1, config.n_action_bins)      # :499-501
        self.bin_centers = (self.bins[:-1] + self.bins[1:]) / 2.0
        # vocab size for de-tokenisation -- revert the added "multiple of" padding (:503-504)
        self.vocab_size = config.text_config.vocab_size - config.pad_to_multiple_of
        self.pad_token_id = config.pad_token_id
        self.engine = Engine(config, max_batch=max_batch, max_prompt_len=max_prompt_len, max_new_tokens=8,
                             device=device)
        self.device = self.engine.device
        self._centers_dev = torch.from_numpy(self.bin_centers).to(self.device)
        self._pin: Dict[str, torch.Tensor] = {}

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, state_dict: Dict[str, torch.Tensor]) -> None:
        """HF names fixed by vla-scripts/extern/convert_openvla_weights_to_hf.py:73-115."""
        self.engine.load_state_dict(state_dict)

    def to(self, *a, **k):          # the reference calls vla.to(DEVICE) (openvla_utils.py:57); weights already live there
        return self

    def eval(self):
        return self

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _check_unnorm_key(norm_stats: Dict[str, Dict[str, Any]], unnorm_key: Optional[str]) -> str:
        """modeling_prismatic.py:538-552 (same AssertionErrors)."""
        if unnorm_key is None:
            assert len(norm_stats) == 1, (
                f"Your model was trained on more than one dataset, "
                f"please pass a `unnorm_key` from the following options to choose the statistics "
                f"used for un-normalizing actions: {norm_stats.keys()}"
            )
            unnorm_key = next(iter(norm_stats.keys()))
        assert unnorm_key in norm_stats, (
            f"The `unnorm_key` you chose is not in the set of available dataset statistics, "
            f"please choose from: {norm_stats.keys()}"
        )
        return unnorm_key

    def get_action_dim(self, unnorm_key: Optional[str] = None) -> int:
        unnorm_key = self._check_unnorm_key(self.norm_stats, unnorm_key)
        return len(self.norm_stats[unnorm_key]["action"]["q01"])

    def get_action_stats(self, unnorm_key: Optional[str] = None) -> Dict[str, Any]:
        unnorm_key = self._check_unnorm_key(self.norm_stats, unnorm_key)
        return self.norm_stats[unnorm_key]["action"]

    def _append_empty(self, input_ids: torch.Tensor, lens: Optional[torch.Tensor] = None):
        """modeling_prismatic.py:512-515: batch-wide test, append 29871 unless every row already ends with it.
        Ragged rows (`lens` = true lengths, right-padded): the test looks at each row's last REAL token and 29871 goes
        right behind it; returns (ids, lens) with the lengths counting the appended token."""
        if lens is None:
            if not torch.all(input_ids[:, -1] == EMPTY_TOKEN_ID):
                pad = torch.full((input_ids.shape[0], 1), EMPTY_TOKEN_ID, dtype=input_ids.dtype, device=input_ids.device)
                input_ids = torch.cat((input_ids, pad), dim=1)
            return input_ids, None
        lens = lens.to("cpu", torch.int64)                     # the lengths stay on the host
        at = lens.to(input_ids.device)
        rows = torch.arange(input_ids.shape[0], device=input_ids.device)
        if torch.all(input_ids[rows, at - 1] == EMPTY_TOKEN_ID):
            return input_ids, lens
        pad = torch.full((input_ids.shape[0], 1), self.pad_token_id, dtype=input_ids.dtype, device=input_ids.device)
        input_ids = torch.cat((input_ids, pad), dim=1)
        input_ids[rows, at] = EMPTY_TOKEN_ID
        return input_ids, lens + 1

    def _check_inputs(self, input_ids, pixel_values, attention_mask):
        if input_ids is None or pixel_values is None:
            raise ValueError("predict_action needs `input_ids` and `pixel_values`")
        if input_ids.dim() != 2:
            raise ValueError("`input_ids` must be [batch, prompt_len]")
        if input_ids.shape[0] != pixel_values.shape[0]:
            # modeling_prismatic.py:418-419
            raise ValueError("Non-homogenous batch of (text, image) input -- forward() does not support mixed batches!")
        lens = None
        if attention_mask is not None and not bool(torch.all(attention_mask != 0)):
            if tuple(attention_mask.shape) != tuple(input_ids.shape):
                raise ValueError("`attention_mask` must have the shape of `input_ids`")
            m = (attention_mask != 0).to(torch.int64).cpu()
            lens = m.sum(1)
            right_padded = bool(torch.all(m == (torch.arange(m.shape[1]).view(1, -1) < lens.view(-1, 1)).to(torch.int64)))
            if not right_padded or int(lens.min()) < 1:
                raise ValueError("ragged prompts must be RIGHT-padded (mask = ones then zeros, at least the BOS token per "
                                 "row); left padding would shift the BOS / patch splice of modeling_prismatic.py:380-390")
        c = self.config
        want = (3 * len(c.towers), c.image_size, c.image_size)
        if tuple(pixel_values.shape[1:]) != want:
            raise ValueError(f"`pixel_values` must be [B, {want[0]}, {want[1]}, {want[2]}], got {tuple(pixel_values.shape)}")
        return lens

    def _finish_sequences(self, input_ids: torch.Tensor, new_tokens: np.ndarray, n: int, lens=None) -> np.ndarray:
        """HF greedy `generate` stops a row at EOS; `generated_ids[0, -n:]` (modeling_prismatic.py:521) then reaches
        back into the prompt.  The engine always produces n tokens (greedy is deterministic, so the prefix up to the
        first EOS is what HF would have produced); the truncation is replayed here per row."""
        eos = self.config.text_config.eos_token_id
        ids_host = input_ids.cpu().numpy()
        out = np.empty((new_tokens.shape[0], n), dtype=np.int64)
        for b in range(new_tokens.shape[0]):
            row = new_tokens[b]
            hit = np.nonzero(row == eos)[0]
            gen = row[: hit[0] + 1] if hit.size else row
            prompt = ids_host[b] if lens is None else ids_host[b, : int(lens[b])]
            seq = np.concatenate([prompt, gen])
            out[b] = seq[-n:]
        return out

    def _detokenize(self, token_ids: np.ndarray, unnorm_key: Optional[str]) -> np.ndarray:
        """modeling_prismatic.py:521-534, float64, on the device (bit-identical to the numpy formula)."""
        import ctypes as C

        stats = self.get_action_stats(unnorm_key)
        q01 = np.asarray(stats["q01"], dtype=np.float64)
        q99 = np.asarray(stats["q99"], dtype=np.float64)
        mask = np.asarray(stats.get("mask", np.ones_like(q01, dtype=bool)), dtype=bool)
        n = token_ids.size
        ids = torch.from_numpy(np.ascontiguousarray(token_ids, dtype=np.int64)).to(self.device)
        q01_d, q99_d = torch.from_numpy(q01).to(self.device), torch.from_numpy(q99).to(self.device)
        mask_d = torch.from_numpy(mask.astype(np.uint8)).to(self.device)
        out = torch.empty(n, dtype=torch.float64, device=self.device)
        lib = _lib.load()
        _lib.check(lib.ovla_detokenize(C.c_void_p(ids.data_ptr()), n, len(q01), self.vocab_size,
                                       C.c_void_p(self._centers_dev.data_ptr()), int(self.bin_centers.shape[0]),
                                       C.c_void_p(q01_d.data_ptr()), C.c_void_p(q99_d.data_ptr()),
                                       C.c_void_p(mask_d.data_ptr()), C.c_void_p(out.data_ptr()), _lib.stream_ptr()))
        return out.cpu().numpy().reshape(token_ids.shape)

    def _pinned(self, key: str, shape, dtype) -> torch.Tensor:
        t = self._pin.get(key)
        n = int(np.prod(shape))
        if t is None or t.dtype != dtype or t.numel() < n:
            t = torch.empty(max(n, 1), dtype=dtype, pin_memory=True)
            self._pin[key] = t
        return t[:n].view(*shape)

    # ------------------------------------------------------------------ frames -> pixel_values on the device
    _TOWER_STATS = {  # timm data configs consumed at convert_openvla_weights_to_hf.py:193-197
        "dino": ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)),
        "siglip": ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)),
    }

    @torch.no_grad()
    def preprocess_frames(self, frames_u8: torch.Tensor) -> torch.Tensor:
        """uint8 HWC frames [B, S, S, 3] already at model resolution -> bf16 `pixel_values` [B, 3*towers, S, S] on the
        device: `PrismaticImageProcessor.apply_transform` (to_tensor + per-tower normalize, processing_prismatic.py:128-145)
        plus the bf16 cast of openvla_utils.py:186, bit-identical to the host path.  Only the uint8 bytes cross PCIe."""
        import ctypes as C

        c = self.config
        if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or tuple(frames_u8.shape[1:]) != (c.image_size, c.image_size, 3):
            raise ValueError(f"frames must be uint8 [B, {c.image_size}, {c.image_size}, 3] (see center_crop_frames; the "
                             "lanczos resize of the simulator frame is host-side input prep)")
        names = ["dino", "siglip"] if c.use_fused_vision_backbone else ["siglip"]
        if not hasattr(self, "_norm_dev"):
            mean = torch.tensor([v for n in names for v in self._TOWER_STATS[n][0]], dtype=torch.float32)
            std = torch.tensor([v for n in names for v in self._TOWER_STATS[n][1]], dtype=torch.float32)
            self._norm_dev = (mean.to(self.device), std.to(self.device))
        fr = frames_u8.to(self.device, non_blocking=True).contiguous()
        B = fr.shape[0]
        out = torch.empty(B, 3 * len(names), c.image_size, c.image_size, dtype=torch.bfloat16, device=self.device)
        if B:
            _lib.check(self.engine.lib.ovla_preprocess_frames(
                C.c_void_p(fr.data_ptr()), B, c.image_size, len(names), C.c_void_p(self._norm_dev[0].data_ptr()),
                C.c_void_p(self._norm_dev[1].data_ptr()), C.c_void_p(out.data_ptr()), _lib.stream_ptr()))
        return out

    @torch.no_grad()
    def center_crop_frames(self, frames_u8: torch.Tensor, crop_scale: float = 0.9) -> torch.Tensor:
        """The reference's `center_crop=True` branch (openvla_utils.py:155-175) on the device: uint8 HWC frames
        [B, H, W, 3] -> centred square crop of area `crop_scale`, bilinear `tf.image.crop_and_resize` back to the model
        resolution, uint8 [B, S, S, 3] (stays on the device: feed it to `preprocess_frames`)."""
        import ctypes as C

        if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[-1] != 3:
            raise ValueError("frames must be uint8 [B, H, W, 3]")
        fr = frames_u8.to(self.device, non_blocking=True).contiguous()
        B, H, W, _ = fr.shape
        S = self.config.image_size
        out = torch.empty(B, S, S, 3, dtype=torch.uint8, device=self.device)
        if B:
            _lib.check(self.engine.lib.ovla_center_crop_frames(
                C.c_void_p(fr.data_ptr()), B, H, W, C.c_float(crop_scale), C.c_void_p(out.data_ptr()), S, _lib.stream_ptr()))
        return out

    # ------------------------------------------------------------------ public surface
    @torch.no_grad()
    def predict_action(self, input_ids: Optional[torch.Tensor] = None, unnorm_key: Optional[str] = None,
                       **kwargs: Any) -> np.ndarray:
        """modeling_prismatic.py:506-536.  kwargs: `pixel_values` (bf16/fp32 [B, 6, 224, 224]), `attention_mask`
        (must be all ones), `do_sample=False` (greedy only)."""
        actions, _ = self._predict(input_ids, unnorm_key, capture=False, **kwargs)
        return actions

    @torch.no_grad()
    def predict_action_and_capture(self, input_ids: torch.Tensor, unnorm_key: Optional[str] = None,
                                   layer_indices: Optional[Sequence[int]] = None, pooling_method: str = "mean",
                                   **kwargs: Any):
        """Fused equivalent of get_vla_action's two passes.  Returns (embeds, actions): embeds maps each requested
        layer index (negative allowed, default (-1,), openvla_utils.py:193-199) to fp32 [B, D] pooled states.

        ``pooled_out=`` (keyword, optional): a caller-owned CPU float32 tensor [n_layers + 1, B, D] (ideally pinned)
        that receives the pooled states directly from the device -- the returned arrays are then views of it and no
        fresh 138 MB result is allocated per call (numpy's ``out=`` convention)."""
        actions, pooled = self._predict(input_ids, unnorm_key, capture=True, pooling_method=pooling_method, **kwargs)
        n = pooled.shape[0]
        embeds = {idx: pooled[idx if idx >= 0 else n + idx] for idx in (layer_indices or (-1,))}
        return embeds, actions

    def _predict(self, input_ids, unnorm_key, capture: bool, pooling_method: str = "mean", pixel_values=None,
                 attention_mask=None, do_sample: bool = False, return_tokens: bool = False, pooled_out=None, **unused):
        if do_sample:
            raise ValueError("only greedy decoding (do_sample=False) is implemented, as used by the reference path")
        lens0 = self._check_inputs(input_ids, pixel_values, attention_mask)
        n_act = self.get_action_dim(unnorm_key)
        B = input_ids.shape[0]
        if B == 0:
            empty = np.zeros((0, n_act), dtype=np.float64)
            return empty, np.zeros((self.config.text_config.num_hidden_layers + 1, 0,
                                    self.config.text_config.hidden_size), dtype=np.float32)
        P0 = input_ids.shape[1]
        ids, lens = self._append_empty(input_ids, lens0)
        P = ids.shape[1]
        lens32 = lens.to(torch.int32).contiguous() if lens is not None else None
        # capture pass of the reference sees the prompt as given (P0 tokens): pool over [0, n_patches + P0); a ragged row
        # pools over its own n_patches + len0_b rows (the engine subtracts P - len_b = P0 - len0_b from pool_len per row)
        pool_len = self.config.n_patches + P0 if capture else 0
        pool_mode = 0 if pooling_method == "mean" else 1
        tc = self.config.text_config
        if ids.is_cuda or pixel_values.is_cuda:
            r = self.engine.run(ids, pixel_values, pool_len, pool_mode, n_act, prompt_lens=lens32)
            tokens = r["tokens"].cpu().numpy()
            pooled = r["pooled"].cpu().numpy() if capture else None
        else:
            # reference-facing host path: pinned staging, H2D + compute + D2H inside one native call
            ids_p = self._pinned("ids", (B, P), torch.int64)
            ids_p.copy_(ids)
            if pixel_values.dtype == torch.bfloat16 and pixel_values.is_contiguous() and pixel_values.is_pinned():
                px_p = pixel_values                   # caller already holds the frame batch in pinned memory
            else:
                px_p = self._pinned("px", tuple(pixel_values.shape), torch.bfloat16)
                px_p.copy_(pixel_values)
            tok_p = self._pinned("tok", (B, n_act), torch.int64)
            pool_shape = (tc.num_hidden_layers + 1, B, tc.hidden_size)
            own_out = capture and pooled_out is not None
            if own_out:
                if (not isinstance(pooled_out, torch.Tensor) or pooled_out.is_cuda or pooled_out.dtype != torch.float32
                        or tuple(pooled_out.shape) != pool_shape or not pooled_out.is_contiguous()):
                    raise ValueError(f"pooled_out must be a contiguous CPU float32 tensor of shape {pool_shape}")
                pool_p = pooled_out
            else:
                pool_p = self._pinned("pool", pool_shape, torch.float32) if capture else None
            self.engine.run_host(ids_p, px_p, pool_len, pool_mode, n_act, pool_p, tok_p, lens32)
            tokens = tok_p.numpy().copy()
            if not capture:
                pooled = None
            elif own_out:
                pooled = pool_p.numpy()               # the caller's buffer: views, no copy
            else:
                pooled = torch.empty(pool_shape, dtype=torch.float32).copy_(pool_p).numpy()   # threaded copy out of staging
        final_ids = self._finish_sequences(ids, tokens, n_act, lens)
        actions = self._detokenize(final_ids, unnorm_key)
        if B == 1:
            actions = actions[0]                      # reference returns shape (action_dim,) (robot_utils.py:78)
        if return_tokens:
            return (actions, final_ids), pooled
        return actions, pooled

    @torch.no_grad()
    def forward(self, input_ids: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                pixel_values: Optional[torch.Tensor] = None, labels=None, inputs_embeds=None, past_key_values=None,
                use_cache=None, output_attentions=None, output_hidden_states: Optional[bool] = None,
                output_projector_features: Optional[bool] = None, return_dict: Optional[bool] = None):
        """Multimodal branch of PrismaticForConditionalGeneration.forward (modeling_prismatic.py:362-447): returns
        logits [B, T, V] fp32, hidden_states (L+1 x [B, T, D] bf16) and projector features as device tensors."""
        if labels is not None or inputs_embeds is not None or past_key_values is not None or output_attentions:
            raise ValueError("only the inference multimodal forward (input_ids + pixel_values) is implemented")
        if pixel_values is None:
            raise ValueError("language-only forward is outside the predict_action path")
        self._check_inputs(input_ids, pixel_values, attention_mask)
        import ctypes as C

        want_h = bool(output_hidden_states) if output_hidden_states is not None else self.config.output_hidden_states
        r = self.engine.run(input_ids, pixel_values, 0, 0, 0, want_hidden=True,
                            want_projector=bool(output_projector_features))
        hidden = r["hidden"]                                       # [L+1, B, T, D]
        Lp1, B, T, D = hidden.shape
        V = self.config.text_config.vocab_size
        logits = torch.empty(B, T, V, dtype=torch.float32, device=self.device)
        epi = _lib.GemmEpilogue()
        epi.round_bf16 = 1
        lm_head = self._lm_head_dev
        _lib.check(self.engine.lib.ovla_gemm(C.c_void_p(hidden[-1].data_ptr()), C.c_longlong(D),
                                             C.c_void_p(lm_head.data_ptr()), C.c_longlong(D), B * T, V, D, 2, 0,
                                             C.c_void_p(logits.data_ptr()), C.c_longlong(V), C.byref(epi), 0, 0,
                                             _lib.stream_ptr()))
        out = PrismaticCausalLMOutputWithPast(
            logits=logits,
            hidden_states=tuple(hidden[i] for i in range(Lp1)) if want_h else None,
            projector_features=r.get("projector"),
        )
        if return_dict is False:
            return (out.logits, out.hidden_states)
        return out

    __call__ = forward


def _attach_lm_head(model: OpenVLAForActionPrediction, state_dict: Dict[str, torch.Tensor]) -> None:
    model._lm_head_dev = state_dict["language_model.lm_head.weight"].to(model.device, torch.bfloat16).contiguous()


def from_state_dict(config: OpenVLAConfig, state_dict: Dict[str, torch.Tensor], max_batch: int = 1,
                    max_prompt_len: int = 40, device: int = 0) -> OpenVLAForActionPrediction:
    """Counterpart of `AutoModelForVision2Seq.from_pretrained(...)` (openvla_utils.py:42-51) for an in-memory
    HF-named state dict (no checkpoint download offline)."""
    m = OpenVLAForActionPrediction(config, max_batch=max_batch, max_prompt_len=max_prompt_len, device=device)
    m.load_state_dict(state_dict)
    _attach_lm_head(m, state_dict)
    return m
