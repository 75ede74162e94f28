"""Request batching of the REST front-end (deploy.py counterpart): host logic with a stand-in model on CPU, and the
real engine on the GPU (each request gets exactly its B = 1 answer)."""
import threading

import numpy as np
import pytest
import torch

from openvla_probe_b200.deploy import OpenVLAServer, get_openvla_prompt
from openvla_probe_b200.vlas import hash_tokenizer


def word_tokenizer(prompt):
    """BOS + one pseudo id per word: prompts of different lengths (hash_tokenizer always returns 30 ids)."""
    return hash_tokenizer(prompt, length=1 + len(prompt.split()))


class _FakeVLA:
    def __init__(self):
        self.calls = []
        self.masks = []

    def preprocess_frames(self, frames):
        return frames.float().mean(dim=(1, 2, 3))

    def predict_action(self, ids, unnorm_key=None, pixel_values=None, attention_mask=None, do_sample=False):
        self.calls.append(ids.shape[0])
        self.masks.append(None if attention_mask is None else attention_mask.clone())
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


def test_mixed_length_instructions_share_one_ragged_batch():
    """Instructions of different lengths are right-padded into ONE pass with a ones-then-zeros attention mask."""
    fake = _FakeVLA()
    srv = OpenVLAServer(fake, word_tokenizer, max_batch=8, max_wait_ms=100.0)
    try:
        texts = ["go", "pick up the red cup", "open the top drawer of the cabinet", "go"]
        futs = [srv.submit(np.full((8, 8, 3), v, np.uint8), t) for v, t in enumerate(texts)]
        res = [f.result(timeout=10) for f in futs]
        assert [float(r[0]) for r in res] == [0.0, 1.0, 2.0, 3.0]
        big = max(range(len(fake.calls)), key=lambda i: fake.calls[i])
        assert fake.calls[big] > 1
        m = fake.masks[big]
        lens = m.sum(dim=1)
        assert len(set(lens.tolist())) > 1                                   # really ragged
        for r in range(m.shape[0]):
            assert bool((m[r, :lens[r]] == 1).all()) and bool((m[r, lens[r]:] == 0).all())
    finally:
        srv.close()


def test_fastapi_shell_serves_act_over_http():
    """The reference's REST shell (vla-scripts/deploy.py:120-123: FastAPI app, POST /act) over a real socket: uvicorn in a
    thread on 127.0.0.1, JSON payloads as the reference's client sends them (plain and double-encoded)."""
    import json
    import socket
    import time
    import urllib.request

    uvicorn = pytest.importorskip("uvicorn")
    pytest.importorskip("fastapi")
    srv = OpenVLAServer(_FakeVLA(), hash_tokenizer, max_batch=4, max_wait_ms=1.0)
    try:
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
    except OSError as ex:                      # a sandbox without loopback sockets: nothing to test here
        srv.close()
        pytest.skip(f"cannot bind a loopback port: {ex}")
    server = uvicorn.Server(uvicorn.Config(srv.make_app(), host="127.0.0.1", port=port, log_level="error"))
    th = threading.Thread(target=server.run, daemon=True)
    th.start()
    try:
        for _ in range(200):
            if server.started:
                break
            time.sleep(0.05)
        assert server.started

        def post(obj):
            req = urllib.request.Request(f"http://127.0.0.1:{port}/act", data=json.dumps(obj).encode(),
                                         headers={"Content-Type": "application/json"})
            with urllib.request.urlopen(req, timeout=20) as r:
                return json.loads(r.read())

        img = np.full((8, 8, 3), 7, np.uint8).tolist()
        assert post({"image": img, "instruction": "lift the lid"}) == [7.0] * 7
        enc = post({"encoded": json.dumps({"image": img, "instruction": "lift the lid", "unnorm_key": "k2"})})
        assert json.loads(enc) == [8.0] * 7
        assert post({"instruction": "no image"}) == "error"
    finally:
        server.should_exit = True
        th.join(timeout=10)
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
        # instructions of different lengths share one ragged pass; every request still gets its own B = 1 answer
        srv.tokenizer = word_tokenizer
        texts = ["go", "pick up the cup", "open the top drawer of the cabinet now", "push it"]
        single2 = []
        for i, t in enumerate(texts):
            ids = torch.tensor([word_tokenizer(get_openvla_prompt(t, "openvla"))])
            single2.append(model.predict_action(ids, unnorm_key="synthetic",
                                                pixel_values=model.preprocess_frames(torch.from_numpy(imgs[i:i + 1]))))
        n0 = len(srv.batches_served)
        futs = [srv.submit(imgs[i], t, "synthetic") for i, t in enumerate(texts)]
        res = [f.result(timeout=60) for f in futs]
        for a, b in zip(res, single2):
            assert np.array_equal(a, b)
        assert max(srv.batches_served[n0:]) > 1
    finally:
        srv.close()
