"""Image transform in front of the model (processing_prismatic.py:128-145): letterbox / resize (PIL bicubic) / center crop /
to_tensor / normalize / bf16.

CPU half (`-m "not gpu"`): the numpy oracle (oracle/image_oracle.py) against the golden vectors produced by the
reference's own PrismaticImageProcessor (tests/golden/make_image_golden.py), bit for bit, and against the installed
Pillow.  GPU half: the CUDA path (ovla_resize_frames + ovla_preprocess_frames through the C ABI) against the oracle and
the same golden digests -- uint8 and bf16 results are compared EXACTLY (integer arithmetic; float32 ops in torch order).
"""
import hashlib
import json
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import image_cases as IC  # noqa: E402
from oracle import image_oracle as IO  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "image_transform_golden.json")))


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _oracle_case(name, strategy, h, w, seed, key):
    g = GOLD["cases"][name]
    kw = IC.KWARGS[key]
    fill = g["letterbox_fill"] or (127, 127, 127)
    u8 = IO.transform_u8(IC.frame(h, w, seed), strategy, 224, fill)
    t = IO.to_tensor_normalize(u8, kw["means"], kw["stds"])
    return g, u8, t


@pytest.mark.parametrize("case", IC.CASES, ids=[c[0] for c in IC.CASES])
def test_oracle_equals_reference_processor_golden(case):
    g, u8, t = _oracle_case(*case)
    assert u8.shape == (224, 224, 3) and list(t.shape) == g["tensor_shape"]
    assert u8[::37, ::41].reshape(-1).tolist() == g["frame_u8_strided"]
    assert sha(u8) == g["frame_u8_sha256"]
    assert sha(t) == g["tensor_f32_sha256"]                                  # float32, bit for bit
    assert sha(torch.from_numpy(t).to(torch.bfloat16).view(torch.int16).numpy()) == g["tensor_bf16_sha256"]


def test_oracle_resample_equals_installed_pillow():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(0)
    for h, w, ow, oh in [(256, 256, 224, 224), (480, 640, 224, 224), (100, 130, 224, 224), (333, 224, 224, 333),
                         (5, 7, 224, 224), (224, 224, 224, 224), (1080, 1920, 224, 224), (300, 301, 17, 9)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
        assert np.array_equal(IO.resize_bicubic_u8(img, ow, oh), ref), (h, w, ow, oh)


def test_oracle_size_and_crop_rules():
    assert IO.resized_output_size(200, 300, 224) == (224, 336) and IO.resized_output_size(300, 200, 224) == (336, 224)
    assert IO.resized_output_size(257, 255, 224) == (225, 224)
    assert IO.center_crop_offsets(225, 224, 224, 224) == (0, 0)              # round(0.5) = 0: half to even
    assert IO.center_crop_offsets(227, 224, 224, 224) == (2, 0)              # round(1.5) = 2
    assert IO.letterbox_pad(np.zeros((201, 300, 3), np.uint8), (127,) * 3).shape == (299, 300, 3)   # odd difference
    _, b, k = IO.resample_coeffs(224, 224)                                   # identity resize: one tap of 2^22
    assert all(int(k[i].sum()) == 1 << 22 and int(k[i].max()) == 1 << 22 for i in range(224))
    with pytest.raises(ValueError):
        IO.transform_u8(np.zeros((8, 8, 3), np.uint8), "stretch")


# ----------------------------------------------------------------------------------------------- CUDA path
def _device_transform(frames, strategy, fill, n_towers):
    import ctypes as C

    from openvla_probe_b200 import _lib

    lib = _lib.load()
    fr = torch.from_numpy(np.ascontiguousarray(frames)).cuda()
    B, H, W, _ = fr.shape
    out = torch.empty(B, 224, 224, 3, dtype=torch.uint8, device="cuda")
    sid = {"resize-naive": 0, "resize-crop": 1, "letterbox": 2}[strategy]
    _lib.check(lib.ovla_resize_frames(fr.data_ptr(), B, H, W, sid, int(fill[0]), int(fill[1]), int(fill[2]), out.data_ptr(),
                                      224, _lib.stream_ptr()))
    stats = [((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)), ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))][2 - n_towers:]
    mean = torch.tensor([v for m, _ in stats for v in m], dtype=torch.float32, device="cuda")
    std = torch.tensor([v for _, s in stats for v in s], dtype=torch.float32, device="cuda")
    px = torch.empty(B, 3 * n_towers, 224, 224, dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.ovla_preprocess_frames(out.data_ptr(), B, 224, n_towers, mean.data_ptr(), std.data_ptr(), px.data_ptr(),
                                          _lib.stream_ptr()))
    torch.cuda.synchronize()
    return out.cpu().numpy(), px.cpu()


@pytest.mark.gpu
@pytest.mark.parametrize("case", IC.CASES, ids=[c[0] for c in IC.CASES])
def test_device_transform_equals_reference_processor_golden(case):
    name, strategy, h, w, seed, key = case
    g, u8_ref, t_ref = _oracle_case(*case)
    n_towers = 2 if key == "fused" else 1
    frames = np.stack([IC.frame(h, w, seed), IC.frame(h, w, seed + 100), IC.frame(h, w, seed)])
    u8, px = _device_transform(frames, strategy, g["letterbox_fill"] or (127, 127, 127), n_towers)
    assert np.array_equal(u8[0], u8_ref) and np.array_equal(u8[2], u8_ref)
    assert sha(u8[0]) == g["frame_u8_sha256"]
    assert sha(px[0].view(torch.int16).numpy()) == g["tensor_bf16_sha256"]
    other = IO.transform_u8(frames[1], strategy, 224, g["letterbox_fill"] or (127, 127, 127))
    assert np.array_equal(u8[1], other)


@pytest.mark.gpu
def test_device_resize_edge_cases():
    import ctypes as C

    from openvla_probe_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(3)
    for h, w, strategy in [(5, 7, "resize-naive"), (1080, 1920, "resize-crop"), (224, 225, "resize-crop"), (3, 500, "letterbox"),
                           (640, 641, "letterbox")]:
        img = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
        u8, _ = _device_transform(img, strategy, (10, 200, 77), 2)
        for b in range(2):
            assert np.array_equal(u8[b], IO.transform_u8(img[b], strategy, 224, (10, 200, 77))), (h, w, strategy)
    fr = torch.zeros(1, 8, 8, 3, dtype=torch.uint8, device="cuda")
    out = torch.empty(1, 224, 224, 3, dtype=torch.uint8, device="cuda")
    assert lib.ovla_resize_frames(fr.data_ptr(), 1, 8, 8, 7, 0, 0, 0, out.data_ptr(), 224, None) != 0       # unknown strategy
    assert lib.ovla_resize_frames(fr.data_ptr(), 0, 8, 8, 0, 0, 0, 0, out.data_ptr(), 224, None) == 0       # empty batch


def test_processor_validation_and_no_cpu_fallback():
    from openvla_probe_b200 import _lib
    from openvla_probe_b200.processing_prismatic import PrismaticImageProcessor

    with pytest.raises(ValueError, match="not supported"):
        PrismaticImageProcessor(image_resize_strategy="stretch")                      # processing_prismatic.py:120-121
    with pytest.raises(ValueError):
        PrismaticImageProcessor(image_resize_strategy="resize-naive", input_sizes=[(3, 224, 224), (3, 384, 384)],
                                interpolations=["bicubic"] * 2, means=[(0.5,) * 3] * 2, stds=[(0.5,) * 3] * 2)
    with pytest.raises(ValueError):
        PrismaticImageProcessor(image_resize_strategy="resize-naive", interpolations=["bilinear"])
    if not torch.cuda.is_available():
        with pytest.raises(_lib.OvlaError):
            PrismaticImageProcessor(image_resize_strategy="resize-naive")


@pytest.mark.gpu
def test_processor_surface_matches_reference_golden():
    """The drop-in PrismaticImageProcessor / PrismaticProcessor (same constructor arguments and methods) on PIL images:
    bf16 pixel_values equal to the reference tensor cast to bf16 (digest), mixed frame sizes in one call, letterbox fill
    taken from the last tower's mean, tokenizer outputs passed through and the batch-size check of :225-226."""
    Image = pytest.importorskip("PIL.Image")
    from openvla_probe_b200.processing_prismatic import PrismaticImageProcessor, PrismaticProcessor

    for strategy in ("resize-naive", "resize-crop", "letterbox"):
        cases = [c for c in IC.CASES if c[1] == strategy and c[5] == "fused"]
        proc = PrismaticImageProcessor(image_resize_strategy=strategy, **IC.FUSED)
        if strategy == "letterbox":
            assert proc.tvf_letterbox_fill == (127, 127, 127)
        imgs = [Image.fromarray(IC.frame(h, w, seed)) for _, _, h, w, seed, _ in cases]
        px = proc(imgs)["pixel_values"]
        assert px.shape == (len(cases), 6, 224, 224) and px.dtype == torch.bfloat16 and px.is_cuda
        for i, c in enumerate(cases):
            assert sha(px[i].cpu().view(torch.int16).numpy()) == GOLD["cases"][c[0]]["tensor_bf16_sha256"], c[0]
        one = proc.apply_transform(imgs[0])
        assert torch.equal(one, px[0])

    class Tok:
        model_input_names = ["input_ids", "attention_mask"]

        def __call__(self, text, return_tensors=None, padding=False, truncation=None, max_length=None):
            n = len(text) if isinstance(text, list) else 1
            return {"input_ids": torch.ones(n, 5, dtype=torch.long), "attention_mask": torch.ones(n, 5, dtype=torch.long)}

    p = PrismaticProcessor(PrismaticImageProcessor(image_resize_strategy="resize-naive", **IC.SIGLIP), Tok())
    img = Image.fromarray(IC.frame(256, 256, 12))
    out = p(["a", "b"], [img, img])
    assert set(out) == {"input_ids", "attention_mask", "pixel_values"} and out["pixel_values"].shape == (2, 3, 224, 224)
    assert sha(out["pixel_values"][1].cpu().view(torch.int16).numpy()) == GOLD["cases"]["libero_256_naive_siglip"]["tensor_bf16_sha256"]
    assert p.model_input_names == ["input_ids", "attention_mask", "pixel_values"]
    with pytest.raises(ValueError, match="malformed"):
        p(["a", "b"], [img])
