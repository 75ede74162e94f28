"""REST front-end with request batching -- counterpart of the reference's `vla-scripts/deploy.py:66-123`
(`OpenVLAServer`: POST /act with {"image": ndarray, "instruction": str, "unnorm_key": optional} -> action).

The reference handles one request at a time on one GPU (deploy.py:120-123).  Here concurrent requests are
micro-batched: a worker thread drains the queue for at most `max_wait_ms` (or until `max_batch` requests are waiting)
and serves them with ONE fused, ragged `predict_action` pass per un-normalisation key -- the natural consumer of the
batch-B capability of the engine.  Each request still receives exactly the B = 1 result (batch / ragged invariance).
`run()` is the reference's FastAPI + uvicorn shell (POST /act); `make_app()` returns the app without starting it.
"""
from __future__ import annotations

import json
import logging
import queue
import threading
import time
import traceback
from concurrent.futures import Future
from typing import Any, Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

SYSTEM_PROMPT = (
    "A chat between a curious user and an artificial intelligence assistant. "
    "The assistant gives helpful, detailed, and polite answers to the user's questions."
)


def get_openvla_prompt(instruction: str, openvla_path: str) -> str:
    """deploy.py:58-62."""
    if "v01" in str(openvla_path):
        return f"{SYSTEM_PROMPT} USER: What action should the robot take to {instruction.lower()}? ASSISTANT:"
    return f"In: What action should the robot take to {instruction.lower()}?\nOut:"


class OpenVLAServer:
    def __init__(self, vla, tokenizer: Callable[[str], Sequence[int]], openvla_path: str = "openvla/openvla-7b",
                 max_batch: int = 16, max_wait_ms: float = 3.0, pad_token_id: int = 0) -> None:
        """`vla`: an `OpenVLAForActionPrediction` (its engine must be created with max_batch >= `max_batch` and a
        max_prompt_len that covers the longest instruction); `tokenizer(prompt) -> ids` (first id = BOS);
        `pad_token_id` fills the masked tail of shorter prompts (never attended).  Frames must already have the model
        resolution (uint8 HxWx3)."""
        self.vla, self.tokenizer, self.openvla_path = vla, tokenizer, openvla_path
        self.pad_token_id = pad_token_id
        self.max_batch, self.max_wait = max_batch, max_wait_ms / 1e3
        self._q: "queue.Queue" = queue.Queue()
        self._stop = threading.Event()
        self._worker = threading.Thread(target=self._serve, daemon=True)
        self._worker.start()
        self.batches_served: List[int] = []

    # ------------------------------------------------------------------ batching worker
    def _serve(self) -> None:
        while not self._stop.is_set():
            try:
                first = self._q.get(timeout=0.05)
            except queue.Empty:
                continue
            batch = [first]
            deadline = time.perf_counter() + self.max_wait
            while len(batch) < self.max_batch:
                left = deadline - time.perf_counter()
                if left <= 0:
                    break
                try:
                    batch.append(self._q.get(timeout=left))
                except queue.Empty:
                    break
            # one fused pass per unnorm_key: instructions of different lengths go into ONE ragged batch, right-padded
            # with an attention mask of ones then zeros (the tokenizer's padding_side="right"; causal attention never
            # looks at the pads, so every request still gets its B = 1 answer -- DESIGN.md "Ragged batches")
            groups: Dict[Any, list] = {}
            for item in batch:
                groups.setdefault(item["unnorm_key"], []).append(item)
            for key, items in groups.items():
                try:
                    longest = max(len(it["ids"]) for it in items)
                    ids = torch.full((len(items), longest), self.pad_token_id, dtype=torch.int64)
                    mask = torch.zeros((len(items), longest), dtype=torch.int64)
                    for r, it in enumerate(items):
                        ids[r, :len(it["ids"])] = torch.tensor(it["ids"], dtype=torch.int64)
                        mask[r, :len(it["ids"])] = 1
                    frames = torch.from_numpy(np.stack([it["image"] for it in items]))
                    px = self.vla.preprocess_frames(frames)
                    actions = self.vla.predict_action(ids, unnorm_key=key, pixel_values=px, attention_mask=mask,
                                                      do_sample=False)
                    actions = np.asarray(actions).reshape(len(items), -1)
                    self.batches_served.append(len(items))
                    for it, a in zip(items, actions):
                        it["future"].set_result(a)
                except Exception as ex:  # noqa: BLE001
                    for it in items:
                        it["future"].set_exception(ex)

    def submit(self, image: np.ndarray, instruction: str, unnorm_key: Optional[str] = None) -> Future:
        prompt = get_openvla_prompt(instruction, self.openvla_path)
        fut: Future = Future()
        self._q.put({"ids": list(self.tokenizer(prompt)), "image": np.ascontiguousarray(image, dtype=np.uint8),
                     "unnorm_key": unnorm_key, "future": fut})
        return fut

    # ------------------------------------------------------------------ reference surface
    def predict_action(self, payload: Dict[str, Any]):
        """deploy.py:91-118: same payload / return conventions ("error" string on failure); returns a plain list so
        that any JSON layer can serialise it."""
        try:
            if double_encode := "encoded" in payload:
                assert len(payload.keys()) == 1, "Only uses encoded payload!"
                payload = json.loads(payload["encoded"])
            image, instruction = np.asarray(payload["image"], dtype=np.uint8), payload["instruction"]
            action = self.submit(image, instruction, payload.get("unnorm_key", None)).result(timeout=60)
            return json.dumps(action.tolist()) if double_encode else action.tolist()
        except Exception:  # noqa: BLE001
            logging.error(traceback.format_exc())
            return "error"

    def make_app(self):
        """The reference's FastAPI shell (deploy.py:120-123): POST /act -> predict_action(payload)."""
        from fastapi import FastAPI

        self.app = FastAPI()
        self.app.post("/act")(self.predict_action)
        return self.app

    def run(self, host: str = "0.0.0.0", port: int = 8000) -> None:
        import uvicorn

        uvicorn.run(self.make_app(), host=host, port=port)

    def close(self) -> None:
        self._stop.set()
        self._worker.join(timeout=2)
