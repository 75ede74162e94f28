"""Native-style twin of the model surface: mirror of `prismatic/models/vlas/openvla.py:35-103`
(`OpenVLA.predict_action(image, instruction, unnorm_key)`) and of `PurePromptBuilder`
(`prismatic/models/backbones/llm/prompting/base_prompter.py:28-73`) on top of the same B200 engine.

The reference builds the prompt, tokenises it with the LLM tokenizer, runs the image transform, generates
`action_dim` tokens under bf16 autocast and de-tokenises with `ActionTokenizer` -- the same arithmetic as the HF
surface.  Tokenizer and PIL transforms are input producers outside the hot path: they are injected
(`tokenizer(prompt) -> list[int]`, `image_transform(image) -> uint8 [S, S, 3]`), with deterministic stand-ins offline.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from .modeling_prismatic import EMPTY_TOKEN_ID, OpenVLAForActionPrediction


class PurePromptBuilder:
    """base_prompter.py:28-73: `In: {message}\\nOut: {reply}` turns, BOS/EOS handled by the tokenizer."""

    def __init__(self, model_family: str = "openvla", system_prompt: Optional[str] = None) -> None:
        self.model_family, self.system_prompt = model_family, system_prompt
        self.bos, self.eos = "<s>", "</s>"
        self.wrap_human = lambda msg: f"In: {msg}\nOut: "
        self.wrap_gpt = lambda msg: f"{msg if msg != '' else ' '}{self.eos}"
        self.prompt, self.turn_count = "", 0

    def add_turn(self, role: str, message: str) -> str:
        assert (role == "human") if (self.turn_count % 2 == 0) else (role == "gpt")
        message = message.replace("<image>", "").strip()
        wrapped = self.wrap_human(message) if (self.turn_count % 2) == 0 else self.wrap_gpt(message)
        self.prompt += wrapped
        self.turn_count += 1
        return wrapped

    def get_potential_prompt(self, message: str) -> str:
        return (self.prompt + self.wrap_human(message)).removeprefix(self.bos).rstrip()

    def get_prompt(self) -> str:
        return self.prompt.removeprefix(self.bos).rstrip()


def hash_tokenizer(prompt: str, length: int = 30) -> List[int]:
    """Offline stand-in for the Llama tokenizer (no tokenizer files here): BOS + deterministic pseudo ids."""
    h = np.frombuffer(prompt.encode("utf-8"), dtype=np.uint8).astype(np.int64)
    rng = np.random.default_rng(int(h.sum()) * 7919 + len(h))
    return [1] + rng.integers(3, 31744, length - 1).tolist()


class OpenVLA:
    """`OpenVLA(PrismaticVLM)` inference surface (vlas/openvla.py:19-103)."""

    def __init__(self, model: OpenVLAForActionPrediction, tokenizer: Callable[[str], Sequence[int]] = hash_tokenizer,
                 image_transform: Optional[Callable] = None) -> None:
        self.model = model
        self.tokenizer = tokenizer
        self.image_transform = image_transform or (lambda im: np.asarray(im, dtype=np.uint8))
        self.norm_stats = model.norm_stats

    def get_prompt_builder(self, system_prompt: Optional[str] = None) -> PurePromptBuilder:
        return PurePromptBuilder("openvla", system_prompt)

    def get_action_dim(self, unnorm_key: Optional[str] = None) -> int:
        return self.model.get_action_dim(unnorm_key)

    def get_action_stats(self, unnorm_key: Optional[str] = None):
        return self.model.get_action_stats(unnorm_key)

    @torch.inference_mode()
    def predict_action(self, image, instruction: str, unnorm_key: Optional[str] = None, **kwargs) -> np.ndarray:
        """vlas/openvla.py:35-103: build the prompt, tokenise, append the empty token 29871 for Llama tokenizers
        (:58-66), transform the image, generate `action_dim` tokens greedily, de-tokenise and un-normalise."""
        pb = self.get_prompt_builder()
        pb.add_turn(role="human", message=f"What action should the robot take to {instruction.lower()}?")
        ids = list(self.tokenizer(pb.get_prompt()))
        if ids[-1] != EMPTY_TOKEN_ID:
            ids.append(EMPTY_TOKEN_ID)
        input_ids = torch.tensor([ids], dtype=torch.int64)
        frame = torch.from_numpy(np.ascontiguousarray(self.image_transform(image)))[None]
        pixel_values = self.model.preprocess_frames(frame)
        return self.model.predict_action(input_ids, unnorm_key=unnorm_key, pixel_values=pixel_values, **kwargs)
