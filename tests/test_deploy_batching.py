"""Request batching of the REST front-end (deploy.py counterpart): host logic with a stand-in model on CPU, and the
real engine on the GPU (each request gets exactly its B = 1 answer)."""
import threading

import numpy as np
import pytest
import torch

from openvla_probe_b200.deploy import OpenVLAServer, get_openvla_prompt
from openvla_probe_b200.vlas import hash_tokenizer


class _FakeVLA:
    def __init__(self):
        self.calls = []

    def preprocess_frames(self, frames):
        return frames.float().mean(dim=(1, 2, 3))

    def predict_action(self, ids, unnorm_key=None, pixel_values=None, do_sample=False):
        self.calls.append(ids.shape[0])
        out = np.stack([np.full(7, float(pixel_values[b]) + (1.0 if unnorm_key == "k2" else 0.0)) for b in range(ids.shape[0])])
        return out[0] if ids.shape[0] == 1 else out


def test_prompt_and_payload_conventions():
    assert get_openvla_prompt("Pick Up X", "openvla/openvla-7b") == "In: What action should the robot take to pick up x?\nOut:"
    assert get_openvla_prompt("a", "openvla-v01-7b").endswith("ASSISTANT:")
    srv = OpenVLAServer(_FakeVLA(), hash_tokenizer, max_batch=4, max_wait_ms=1.0)
    try:
        img = np.full((8, 8, 3), 10, np.uint8)
        assert srv.predict_action({"image": img, "instruction": "go"}) == [10.0] * 7
        assert srv.predict_action({"instruction": "no image"}) == "error"          # reference returns "error" (deploy.py:110-118)
        import json
        enc = srv.predict_action({"encoded": json.dumps({"image": img.tolist(), "instruction": "go", "unnorm_key": "k2"})})
        assert json.loads(enc) == [11.0] * 7
    finally:
        srv.close()


def test_concurrent_requests_are_batched_and_routed_back():
    fake = _FakeVLA()
    srv = OpenVLAServer(fake, hash_tokenizer, max_batch=8, max_wait_ms=50.0)
    try:
        futs = [srv.submit(np.full((8, 8, 3), v, np.uint8), "same instruction") for v in range(8)]
        res = [f.result(timeout=10) for f in futs]
        assert [float(r[0]) for r in res] == [float(v) for v in range(8)]
        assert max(fake.calls) > 1 and sum(fake.calls) == 8                      # served in fewer, batched passes
    finally:
        srv.close()


@pytest.mark.gpu
def test_batched_server_matches_single_requests_on_gpu():
    import dataclasses

    from openvla_probe_b200 import config as cfgmod
    from openvla_probe_b200.modeling_prismatic import from_state_dict
    from oracle import openvla_oracle as O

    od = O.tiny_dims()
    cfg = dataclasses.replace(cfgmod.tiny(), norm_stats={"synthetic": {"action": O.default_stats()}})
    model = from_state_dict(cfg, O.make_weights(od, seed=0), max_batch=4, max_prompt_len=40)
    rng = np.random.default_rng(3)
    imgs = rng.integers(0, 256, (4, od.image_size, od.image_size, 3), dtype=np.uint8)
    single = []
    for i in range(4):
        ids = torch.tensor([hash_tokenizer(get_openvla_prompt("pick up the cup", "openvla"))])
        single.append(model.predict_action(ids, unnorm_key="synthetic", pixel_values=model.preprocess_frames(torch.from_numpy(imgs[i:i + 1]))))
    srv = OpenVLAServer(model, hash_tokenizer, max_batch=4, max_wait_ms=200.0)
    try:
        futs = [srv.submit(imgs[i], "Pick up the cup", "synthetic") for i in range(4)]
        res = [f.result(timeout=60) for f in futs]
        for a, b in zip(res, single):
            assert np.array_equal(a, b)
        assert max(srv.batches_served) > 1
    finally:
        srv.close()
