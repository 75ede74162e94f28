"""Capture + act shim: mirror of the reference's `experiments/robot/openvla_utils.py:126-207` and
`experiments/robot/robot_utils.py:63-79` on top of the B200 engine.

`get_vla_action` keeps the reference signature and return convention (`action` or `(embeds, action)`), but the two
reference passes (capture forward + predict_action) run as one fused prefill (`predict_action_and_capture`).
Image pre-processing (PIL resize / TF center-crop, processing_prismatic.py:128-145, openvla_utils.py:149-175) is an
input producer outside the hot path: `processor(prompt, image)` must return a mapping with `input_ids` and
`pixel_values` exactly as the reference's PrismaticProcessor does; `SyntheticProcessor` below produces the
synthetic equivalents used by the benchmarks (no tokenizer / PIL files offline).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch

ACTION_DIM = 7

OPENVLA_V01_SYSTEM_PROMPT = (
    "A chat between a curious user and an artificial intelligence assistant. "
    "The assistant gives helpful, detailed, and polite answers to the user's questions."
)


def pool_tokens(tokens: torch.Tensor, method: str = "mean") -> torch.Tensor:
    """openvla_utils.py:126-137 (kept for callers that pool materialised hidden states themselves)."""
    pooled = tokens.mean(1) if method == "mean" else tokens[:, -1]
    assert pooled.shape[0] == 1, (
        f"Expected batch=1, got {pooled.shape[0]}. "
        "Either vectorise downstream or change pooling.")
    return pooled.squeeze(0)


def build_prompt(base_vla_name: str, task_label: str) -> str:
    """openvla_utils.py:177-183."""
    if "openvla-v01" in base_vla_name:
        return f"{OPENVLA_V01_SYSTEM_PROMPT} USER: What action should the robot take to {task_label.lower()}? ASSISTANT:"
    return f"In: What action should the robot take to {task_label.lower()}?\nOut:"


class SyntheticProcessor:
    """Stand-in for PrismaticProcessor.__call__ (processing_prismatic.py:187-216) when no tokenizer files exist:
    deterministic pseudo-token ids for the prompt and per-tower normalisation of an already 224x224 uint8 frame
    (resize / crop need PIL and are outside the hot path)."""

    def __init__(self, config, prompt_len: int = 31, image_processor=None):
        """`image_processor`: a processing_prismatic.PrismaticImageProcessor -- frames of any size then take the real
        image transform (letterbox / bicubic resize / center crop / normalize) on the device."""
        self.config = config
        self.prompt_len = prompt_len
        self.image_processor = image_processor

    def __call__(self, prompt: str, image) -> Dict[str, torch.Tensor]:
        img = np.asarray(image)
        c = self.config
        if self.image_processor is not None:
            pixel_values = self.image_processor.apply_transform(img)[None]
        else:
            if img.shape != (c.image_size, c.image_size, 3) or img.dtype != np.uint8:
                raise ValueError(f"SyntheticProcessor expects a uint8 [{c.image_size},{c.image_size},3] frame "
                                 "(pass image_processor= for other sizes)")
            x = torch.from_numpy(img.copy()).permute(2, 0, 1).float() / 255.0
            stats = [((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)), ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))]
            if not c.use_fused_vision_backbone:
                stats = stats[1:]
            chans = [(x - torch.tensor(m).view(3, 1, 1)) / torch.tensor(s).view(3, 1, 1) for m, s in stats]
            pixel_values = torch.cat(chans, 0)[None]
        h = np.frombuffer(prompt.encode("utf-8"), dtype=np.uint8).astype(np.int64)
        rng = np.random.default_rng(int(h.sum()) * 7919 + len(h))
        ids = np.concatenate([[1], rng.integers(3, 31744, self.prompt_len - 1)]).astype(np.int64)
        input_ids = torch.from_numpy(ids)[None]
        return {"input_ids": input_ids, "attention_mask": torch.ones_like(input_ids), "pixel_values": pixel_values}


def get_vla_action(vla, processor, base_vla_name, obs, task_label, unnorm_key, center_crop=False, *,
                   layer_indices: Optional[Sequence[int]] = None, pooling_method: str = "mean",
                   return_embeddings: bool = False):
    """openvla_utils.py:140-207.  `obs["full_image"]` is a uint8 HxWx3 frame.  Returns `action` (float64 [7]) or
    `(embeds, action)` with embeds = {layer_idx: float32 [4096]}."""
    frame = obs["full_image"]
    if center_crop:
        # openvla_utils.py:155-175: crop scale 0.9 (side sqrt(0.9)), bilinear crop_and_resize back to 224 x 224, uint8;
        # done by the CUDA kernel behind `vla.center_crop_frames` instead of TensorFlow
        if not hasattr(vla, "center_crop_frames"):
            raise NotImplementedError("center_crop needs a model object with `center_crop_frames`")
        frame = vla.center_crop_frames(torch.from_numpy(np.ascontiguousarray(frame))[None], 0.9)[0].cpu().numpy()
    prompt = build_prompt(base_vla_name, task_label)
    inputs = processor(prompt, frame)
    input_ids = inputs["input_ids"]
    pixel_values = inputs["pixel_values"].to(torch.bfloat16)       # `.to(DEVICE, dtype=torch.bfloat16)`, :186
    attention_mask = inputs.get("attention_mask")
    if return_embeddings:
        embeds, action = vla.predict_action_and_capture(
            input_ids, unnorm_key=unnorm_key, layer_indices=layer_indices, pooling_method=pooling_method,
            pixel_values=pixel_values, attention_mask=attention_mask, do_sample=False)
        # reference pools a batch of one and squeezes to (D,) (:126-137)
        embeds = {k: np.ascontiguousarray(v[0]) for k, v in embeds.items()}
        return embeds, action
    return vla.predict_action(input_ids, unnorm_key=unnorm_key, pixel_values=pixel_values,
                              attention_mask=attention_mask, do_sample=False)


def get_action(cfg, model, obs, task_label, processor=None, return_embeddings=False, layer_indices=None,
               pooling_method="mean"):
    """robot_utils.py:63-79: always returns (embeds-or-None, action) and asserts the action shape."""
    if getattr(cfg, "model_family", "openvla") != "openvla":
        raise ValueError("Unexpected `model_family` found in config.")
    res = get_vla_action(model, processor, getattr(cfg, "pretrained_checkpoint", "openvla"), obs, task_label,
                         getattr(cfg, "unnorm_key", None), center_crop=getattr(cfg, "center_crop", False),
                         layer_indices=layer_indices, pooling_method=pooling_method,
                         return_embeddings=return_embeddings)
    embeds, action = res if return_embeddings else (None, res)
    assert action.shape == (ACTION_DIM,)
    return embeds, action
