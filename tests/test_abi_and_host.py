"""CPU-side checks: the C-ABI library builds/loads and exports every symbol include/ovla_b200.h declares (no compute
calls without a GPU), and the host logic of the drop-in surface (29871 rule, EOS replay, error conventions)."""
import ctypes
import dataclasses
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from openvla_probe_b200 import _lib
    from openvla_probe_b200.build import build

    build(verbose=False)
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "ovla_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(ovla_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ovla_b200.h but not exported by libovla_b200.so"
    assert lib.ovla_abi_version() == 2


def test_every_entry_point_has_header_derived_argtypes(lib):
    """no call may rely on the call site wrapping 64-bit values by hand -- argtypes / restype of every symbol
    come from include/ovla_b200.h itself (`_lib.header_prototypes`)."""
    import ctypes as C

    from openvla_probe_b200 import _lib

    protos = _lib.header_prototypes()
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "ovla_b200.h")).read(), flags=re.S)
    assert sorted(protos) == sorted(set(re.findall(r"\b(ovla_[a-z0-9_]+)\s*\(", hdr)))
    for name, (restype, argtypes) in protos.items():
        fn = getattr(lib, name)
        assert fn.argtypes is not None and list(fn.argtypes) == argtypes, name
        assert fn.restype is restype, name
    # spot checks of the mapping: pitches are 64-bit, handles and buffers are pointers, floats stay floats
    g = protos["ovla_gemm"][1]
    assert g[1] is C.c_longlong and g[3] is C.c_longlong and g[10] is C.c_longlong and g[0] is C.c_void_p
    assert protos["ovla_probe_adamw"][1][4] is C.c_longlong and protos["ovla_probe_adamw"][1][9] is C.c_float
    assert protos["ovla_workspace_bytes"][0] is C.c_longlong and protos["ovla_last_error"][0] is C.c_char_p
    assert protos["ovla_bind_weight"][1][1] is C.c_char_p


def test_gemm_tile_walk_visits_every_tile_once(lib):
    """The persistent GEMM's rasterisation (gemm.cuh `gemm_tile_coords`: row groups inside column super-groups,
    optionally serpentine) evaluated on the HOST through ovla_debug_gemm_tile_order: for every knob combination the
    walk must be a bijection onto the num_m x num_n tiles -- a tile visited twice or never would be a silent wrong
    result -- and must keep the locality it is there for: consecutive tiles of a group stay inside the group's rows and
    the super-group's columns, and a serpentine walk enters each row group at the column the previous one left."""
    def walk(num_m, num_n, G, NC, serp):
        mb = (ctypes.c_int * (num_m * num_n))()
        nb = (ctypes.c_int * (num_m * num_n))()
        assert lib.ovla_debug_gemm_tile_order(num_m, num_n, G, NC, serp, mb, nb) == 0
        return list(zip(mb, nb))

    for num_m, num_n in [(1, 1), (1, 7), (9, 1), (3, 5), (16, 16), (17, 5), (36, 11), (283, 16), (72, 86)]:
        for G in (1, 2, 3, 8, 16, 500):
            for NC in (0, 1, 3, 8, 1000):
                for serp in (0, 1):
                    tiles = walk(num_m, num_n, G, NC, serp)
                    assert sorted(tiles) == [(m, n) for m in range(num_m) for n in range(num_n)], (num_m, num_n, G, NC, serp)
                    ncs = NC if 0 < NC < num_n else num_n
                    per_sg = num_m * ncs
                    for t, (m, n) in enumerate(tiles):
                        sg, r = divmod(t, per_sg)
                        assert sg * ncs <= n < min(num_n, (sg + 1) * ncs)          # inside its column super-group
                        width = min(ncs, num_n - sg * ncs)
                        assert m // G == r // (G * width)                          # inside its row group
    # the default walks of the four Llama prefill GEMMs at bs = 256 (283 x {48, 16, 86, 16} tiles)
    t = walk(283, 86, 16, 0, 1)
    assert t[:3] == [(0, 0), (1, 0), (2, 0)] and t[16] == (0, 1)
    assert t[16 * 86 - 1] == (15, 85) and t[16 * 86] == (16, 85) and t[2 * 16 * 86] == (32, 0)    # serpentine turn-around
    t = walk(283, 16, 2, 8, 0)
    assert t[:4] == [(0, 0), (1, 0), (0, 1), (1, 1)] and t[16] == (2, 0) and t[283 * 8] == (0, 8)
    assert lib.ovla_debug_gemm_tile_order(0, 4, 1, 0, 0, None, None) != 0


def test_library_contains_blackwell_instructions():
    """The shipped binary is sm_100a SASS with tcgen05 / TMA / TMEM instructions (no PTX-JIT, no fallback arch)."""
    import shutil
    import subprocess

    from openvla_probe_b200 import _lib

    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", str(_lib.lib_path())], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic


def test_no_cpu_fallback_without_gpu():
    from openvla_probe_b200 import _lib, config
    from openvla_probe_b200.modeling_prismatic import OpenVLAForActionPrediction

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.OvlaError):
        OpenVLAForActionPrediction(config.tiny())


def _bare_model():
    """Host-logic-only instance (no engine): the methods under test never touch the device."""
    from openvla_probe_b200 import config
    from openvla_probe_b200.modeling_prismatic import OpenVLAForActionPrediction

    stats = {"a": {"action": {"q01": [0.0] * 7, "q99": [1.0] * 7}}, "b": {"action": {"q01": [0.0] * 3, "q99": [1.0] * 3}}}
    m = object.__new__(OpenVLAForActionPrediction)
    m.config = dataclasses.replace(config.tiny(), norm_stats=stats)
    m.norm_stats = stats
    m.pad_token_id = m.config.pad_token_id
    return m


def test_unnorm_key_conventions():
    m = _bare_model()
    with pytest.raises(AssertionError):
        m.get_action_dim(None)                 # more than one dataset and no key (modeling_prismatic.py:539-546)
    with pytest.raises(AssertionError):
        m.get_action_dim("missing")
    assert m.get_action_dim("a") == 7 and m.get_action_dim("b") == 3
    assert m.get_action_stats("b")["q99"] == [1.0] * 3


def test_append_29871_is_batch_wide():
    m = _bare_model()
    ids = torch.tensor([[1, 5, 29871], [1, 6, 29871]])
    assert m._append_empty(ids)[0].shape == (2, 3)
    ids[0, -1] = 9
    out, lens = m._append_empty(ids)
    assert lens is None and out.shape == (2, 4) and out[:, -1].tolist() == [29871, 29871]


def test_append_29871_ragged_rows_get_it_behind_their_last_real_token():
    """Right-padded rows: the batch-wide test of modeling_prismatic.py:512 looks at each row's last real token, 29871
    lands right behind it and the lengths count it."""
    m = _bare_model()
    pad = m.pad_token_id
    ids = torch.tensor([[1, 5, 6, 7], [1, 8, pad, pad]])
    out, lens = m._append_empty(ids, torch.tensor([4, 2]))
    assert lens.tolist() == [5, 3]
    assert out.tolist() == [[1, 5, 6, 7, 29871], [1, 8, 29871, pad, pad]]
    # every row already ends with 29871 (before its pads): nothing appended
    ids = torch.tensor([[1, 5, 6, 29871], [1, 29871, pad, pad]])
    out, lens = m._append_empty(ids, torch.tensor([4, 2]))
    assert out.shape == (2, 4) and lens.tolist() == [4, 2]
    # the EOS replay reaches back into the row's REAL prompt, not into the pads
    ids = torch.tensor([[1, 5, 6, 7, 29871], [1, 8, 29871, pad, pad]])
    toks = np.array([[9, 2, 4, 4, 4, 4, 4], [9, 9, 9, 9, 9, 9, 3]])
    got = m._finish_sequences(ids, toks, 7, lens=torch.tensor([5, 3]))
    assert got[0].tolist() == [1, 5, 6, 7, 29871, 9, 2] and got[1].tolist() == toks[1].tolist()


def test_eos_replay_matches_hf_semantics():
    m = _bare_model()
    ids = torch.arange(10, 20).view(1, 10)
    toks = np.array([[7, 8, 2, 4, 5, 6, 7]])
    got = m._finish_sequences(ids, toks, 7)
    assert got.tolist() == [[16, 17, 18, 19, 7, 8, 2]]
    toks = np.array([[7, 8, 9, 4, 5, 6, 3]])
    assert m._finish_sequences(ids, toks, 7).tolist() == toks.tolist()


def test_input_validation_errors():
    m = _bare_model()
    ids = torch.ones(2, 5, dtype=torch.long)
    px = torch.zeros(2, 6, 56, 56)
    with pytest.raises(ValueError):
        m._check_inputs(ids, px[:1], None)
    with pytest.raises(ValueError):
        m._check_inputs(ids, px[:, :3], None)
    mask = torch.ones(2, 5, dtype=torch.long)
    mask[1, -1] = 0
    assert m._check_inputs(ids, px, mask).tolist() == [5, 4]          # right-padded: the rows' true lengths
    assert m._check_inputs(ids, px, torch.ones(2, 5)) is None
    mask[1] = torch.tensor([0, 1, 1, 1, 1])                           # left padding is rejected
    with pytest.raises(ValueError):
        m._check_inputs(ids, px, mask)
    mask[1] = torch.tensor([1, 0, 1, 1, 1])                           # holes are rejected
    with pytest.raises(ValueError):
        m._check_inputs(ids, px, mask)
    mask[1] = 0                                                       # an empty row is rejected
    with pytest.raises(ValueError):
        m._check_inputs(ids, px, mask)


def test_state_dict_schema_matches_oracle_and_golden():
    from helpers import pair
    from openvla_probe_b200 import weights
    from oracle import openvla_oracle as O

    for kind, fused in (("tiny", True), ("tiny", False), ("openvla-7b", True), ("siglip-7b", False)):
        od, pc = pair(kind, fused=fused)
        assert weights.state_dict_shapes(pc) == O.weight_shapes(od)
    od, pc = pair("openvla-7b")
    n_params = sum(int(np.prod(s)) for s in weights.state_dict_shapes(pc).values())
    assert 7.5e9 < n_params < 7.7e9


def test_synthetic_processor_and_prompt():
    from openvla_probe_b200 import config
    from openvla_probe_b200.openvla_utils import SyntheticProcessor, build_prompt, pool_tokens

    assert build_prompt("openvla", "Pick Up The Cup") == "In: What action should the robot take to pick up the cup?\nOut:"
    assert build_prompt("openvla-v01-7b", "x").endswith("ASSISTANT:")
    p = SyntheticProcessor(config.tiny(), prompt_len=11)
    out = p("hello", np.zeros((56, 56, 3), np.uint8))
    assert out["input_ids"].shape == (1, 11) and out["input_ids"][0, 0] == 1
    assert out["pixel_values"].shape == (1, 6, 56, 56)
    assert torch.allclose(out["pixel_values"][0, 3:], torch.full((3, 56, 56), -1.0))
    with pytest.raises(AssertionError):
        pool_tokens(torch.zeros(2, 3, 4))
    assert pool_tokens(torch.ones(1, 3, 4), "final").shape == (4,)


def test_get_vla_action_center_crop_wiring():
    """`center_crop=True` (openvla_utils.py:155-175): the frame goes through the model object's `center_crop_frames`
    (crop scale 0.9) before the processor, and the cropped uint8 frame is what gets normalised.  A fake model backed by
    the oracle's restatement stands in for the CUDA kernel here; the kernel itself is compared byte for byte on the GPU."""
    from oracle import openvla_oracle as O
    from openvla_probe_b200 import config
    from openvla_probe_b200.openvla_utils import SyntheticProcessor, get_vla_action

    class Fake:
        def __init__(self):
            self.seen = None

        def center_crop_frames(self, frames, scale):
            assert frames.dtype == torch.uint8 and frames.dim() == 4 and abs(scale - 0.9) < 1e-12
            return torch.from_numpy(O.center_crop_frames(frames.numpy(), scale, 56))

        def predict_action(self, input_ids, unnorm_key=None, pixel_values=None, **kw):
            self.seen = pixel_values
            return np.zeros(7)

    rng = np.random.default_rng(0)
    frame = rng.integers(0, 256, (56, 56, 3), dtype=np.uint8)
    proc = SyntheticProcessor(config.tiny(), prompt_len=9)
    vla = Fake()
    get_vla_action(vla, proc, "openvla", {"full_image": frame}, "do it", "k", center_crop=True)
    cropped = O.center_crop_frames(frame[None], 0.9, 56)[0]
    want = proc("x", cropped)["pixel_values"].to(torch.bfloat16)
    assert torch.equal(vla.seen, want) and not np.array_equal(cropped, frame)
    vla2 = Fake()
    get_vla_action(vla2, proc, "openvla", {"full_image": frame}, "do it", "k", center_crop=False)
    assert torch.equal(vla2.seen, proc("x", frame)["pixel_values"].to(torch.bfloat16))

    class NoCrop:
        pass

    with pytest.raises(NotImplementedError):
        get_vla_action(NoCrop(), proc, "openvla", {"full_image": frame}, "do it", "k", center_crop=True)
