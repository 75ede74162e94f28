"""Drop-in surface of the reference's `PrismaticImageProcessor` / `PrismaticProcessor`
(prismatic/extern/hf/processing_prismatic.py:31-260) with the image transform on the device.

The reference builds a timm / torchvision transform per tower and applies `TVF.resize -> TVF.center_crop -> TVF.to_tensor
-> TVF.normalize` to a PIL image on the host (:128-145).  Here the uint8 frame is uploaded once and
`ovla_resize_frames` (PIL's bicubic resampling, bit-identical) + `ovla_preprocess_frames` (to_tensor, per-tower normalize,
the bf16 cast of openvla_utils.py:186) produce `pixel_values` on the GPU.  Same constructor arguments, `apply_transform`,
`preprocess`, `__call__`; the float32 tensor of the reference is never materialised -- the bf16 result equals
`reference_tensor.to(torch.bfloat16)` bit for bit (tests/test_image_transform.py).  There is no host fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib

_STRATEGY = {"resize-naive": 0, "resize-crop": 1, "letterbox": 2}


class PrismaticImageProcessor:
    model_input_names = ["pixel_values"]

    def __init__(self, use_fused_vision_backbone: bool = False, image_resize_strategy: str = "letterbox",
                 input_sizes: Optional[List[Tuple[int, int, int]]] = None, interpolations: Optional[List[str]] = None,
                 means: Optional[List[Tuple[float, float, float]]] = None,
                 stds: Optional[List[Tuple[float, float, float]]] = None, device: int = 0, **kwargs: Any) -> None:
        """processing_prismatic.py:34-123 (same names and defaults).  All towers must share one square input size and
        bicubic interpolation, which is what every OpenVLA / Prismatic checkpoint uses."""
        self.use_fused_vision_backbone = use_fused_vision_backbone
        self.image_resize_strategy = image_resize_strategy
        input_sizes = [(3, 224, 224)] if input_sizes is None else [tuple(s) for s in input_sizes]
        interpolations = ["bicubic"] * len(input_sizes) if interpolations is None else list(interpolations)
        means = [(0.5, 0.5, 0.5)] if means is None else [tuple(m) for m in means]
        stds = [(0.5, 0.5, 0.5)] if stds is None else [tuple(s) for s in stds]
        self.input_sizes, self.interpolations, self.means, self.stds = input_sizes, interpolations, means, stds
        if image_resize_strategy not in _STRATEGY:
            raise ValueError(f"Image resize strategy `{image_resize_strategy}` is not supported!")     # :120-121
        if not (len(input_sizes) == len(interpolations) == len(means) == len(stds)) or len(input_sizes) not in (1, 2):
            raise ValueError("one (input_size, interpolation, mean, std) per vision tower, 1 or 2 towers")
        if len(set(input_sizes)) != 1 or input_sizes[0][0] != 3 or input_sizes[0][1] != input_sizes[0][2]:
            raise ValueError(f"towers must share one square RGB input size, got {input_sizes}")
        if any(i != "bicubic" for i in interpolations):
            raise ValueError("only bicubic interpolation is implemented (the timm data configs of DINOv2 / SigLIP)")
        self.size = input_sizes[0][-1]
        # the reference sets the letterbox fill inside its per-tower loop: the LAST tower's mean wins (:116-117)
        self.tvf_do_letterbox = image_resize_strategy == "letterbox"
        self.tvf_letterbox_fill = tuple(int(x * 255) for x in means[-1]) if self.tvf_do_letterbox else None
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.OvlaError("PrismaticImageProcessor needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device)
        self._mean = torch.tensor([v for m in means for v in m], dtype=torch.float32, device=self.device)
        self._std = torch.tensor([v for s in stds for v in s], dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------ frames
    @staticmethod
    def _as_u8(img) -> torch.Tensor:
        """PIL image (`img.convert("RGB")`, :163) or uint8 array / tensor [H, W, 3] -> CPU uint8 tensor."""
        if hasattr(img, "convert"):
            img = np.array(img.convert("RGB"))            # a writable copy
        t = torch.as_tensor(np.asarray(img) if not isinstance(img, torch.Tensor) else img)
        if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[-1] != 3:
            raise ValueError("images must be PIL images or uint8 [H, W, 3] arrays")
        return t

    @torch.no_grad()
    def transform_frames(self, frames_u8: torch.Tensor) -> torch.Tensor:
        """uint8 [B, H, W, 3] (host or device) -> bf16 `pixel_values` [B, 3 * towers, S, S] on the device."""
        if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[-1] != 3:
            raise ValueError("frames must be uint8 [B, H, W, 3]")
        fr = frames_u8.to(self.device, non_blocking=True).contiguous()
        B, H, W, _ = fr.shape
        S, n_t = self.size, len(self.means)
        u8 = torch.empty(B, S, S, 3, dtype=torch.uint8, device=self.device)
        out = torch.empty(B, 3 * n_t, S, S, dtype=torch.bfloat16, device=self.device)
        if B == 0:
            return out
        fill = self.tvf_letterbox_fill or (0, 0, 0)
        with torch.cuda.device(self.device):
            st = _lib.stream_ptr()
            _lib.check(self.lib.ovla_resize_frames(fr.data_ptr(), B, H, W, _STRATEGY[self.image_resize_strategy],
                                                   fill[0], fill[1], fill[2], u8.data_ptr(), S, st))
            _lib.check(self.lib.ovla_preprocess_frames(u8.data_ptr(), B, S, n_t, self._mean.data_ptr(), self._std.data_ptr(),
                                                       out.data_ptr(), st))
        return out

    def apply_transform(self, img) -> torch.Tensor:
        """processing_prismatic.py:128-145 for one image: bf16 [3 * towers, S, S] on the device."""
        return self.transform_frames(self._as_u8(img)[None])[0]

    def preprocess(self, images, return_tensors: Optional[str] = None, **_: Any) -> Dict[str, torch.Tensor]:
        """:146-169.  Images of one size go through the device in one batch; mixed sizes one launch pair per size."""
        if not isinstance(images, (list, tuple)):
            images = [images]
        frames = [self._as_u8(im) for im in images]
        out: List[Optional[torch.Tensor]] = [None] * len(frames)
        by_shape: Dict[Tuple[int, int], List[int]] = {}
        for i, f in enumerate(frames):
            by_shape.setdefault((int(f.shape[0]), int(f.shape[1])), []).append(i)
        for idxs in by_shape.values():
            px = self.transform_frames(torch.stack([frames[i] for i in idxs]))
            for j, i in enumerate(idxs):
                out[i] = px[j]
        S, n_t = self.size, len(self.means)
        pixel_values = torch.stack(out) if out else torch.empty(0, 3 * n_t, S, S, dtype=torch.bfloat16, device=self.device)
        return {"pixel_values": pixel_values}

    def __call__(self, images, **kwargs) -> Dict[str, torch.Tensor]:
        return self.preprocess(images, **kwargs)


class PrismaticProcessor:
    """processing_prismatic.py:174-260: image processor + tokenizer.  The tokenizer is whatever HF-style callable the
    caller has (`tokenizer(text, return_tensors=..., padding=..., truncation=..., max_length=...)` returning
    `input_ids` / `attention_mask`); right-padded ragged batches are what `predict_action` accepts."""
    attributes = ["image_processor", "tokenizer"]

    def __init__(self, image_processor: Optional[PrismaticImageProcessor] = None, tokenizer=None) -> None:
        self.image_processor, self.tokenizer = image_processor, tokenizer

    def __call__(self, text, images, padding=False, truncation=None, max_length=None, return_tensors="pt") -> Dict[str, Any]:
        pixel_values = self.image_processor(images, return_tensors=return_tensors)["pixel_values"]
        text_inputs = self.tokenizer(text, return_tensors=return_tensors, padding=padding, truncation=truncation,
                                     max_length=max_length)
        ids = text_inputs["input_ids"] if isinstance(text_inputs, dict) else text_inputs.input_ids
        if pixel_values.shape[0] != ids.shape[0]:
            raise ValueError("Batch is malformed; expected same number of images and text inputs!")      # :225-226
        return {**dict(text_inputs), "pixel_values": pixel_values}

    def batch_decode(self, sequences, skip_special_tokens: bool = False, clean_up_tokenization_spaces=None, **kwargs):
        return self.tokenizer.batch_decode(sequences=sequences, skip_special_tokens=skip_special_tokens,
                                           clean_up_tokenization_spaces=clean_up_tokenization_spaces, **kwargs)

    def decode(self, token_ids, skip_special_tokens: bool = False, clean_up_tokenization_spaces=None, **kwargs):
        return self.tokenizer.decode(token_ids=token_ids, skip_special_tokens=skip_special_tokens,
                                     clean_up_tokenization_spaces=clean_up_tokenization_spaces, **kwargs)

    @property
    def model_input_names(self) -> List[str]:
        return list(dict.fromkeys(list(self.tokenizer.model_input_names) + self.image_processor.model_input_names))


def openvla_image_processor(config, device: int = 0) -> PrismaticImageProcessor:
    """The processor of an OpenVLA checkpoint (convert_openvla_weights_to_hf.py:193-207): timm data configs of the
    towers, `image_resize_strategy` from the config (openvla-7b: "resize-naive")."""
    fused = config.use_fused_vision_backbone
    means = [(0.485, 0.456, 0.406), (0.5, 0.5, 0.5)] if fused else [(0.5, 0.5, 0.5)]
    stds = [(0.229, 0.224, 0.225), (0.5, 0.5, 0.5)] if fused else [(0.5, 0.5, 0.5)]
    S = config.image_size
    return PrismaticImageProcessor(use_fused_vision_backbone=fused,
                                   image_resize_strategy=getattr(config, "image_resize_strategy", "resize-naive"),
                                   input_sizes=[(3, S, S)] * len(means), interpolations=["bicubic"] * len(means),
                                   means=means, stds=stds, device=device)
