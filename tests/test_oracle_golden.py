"""Pin the CPU oracle against fixtures produced by the REFERENCE's own code (tests/golden/make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import torch

from oracle import openvla_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def _tok():
    with open(os.path.join(G, "action_tokenizer_golden.json")) as f:
        return json.load(f)


def test_detokenizer_known_answers():
    """values obtained from the reference's ActionTokenizer (action_tokenizer.py:49-68)."""
    g = _tok()
    ids = np.array([31744, 31745, 31999, 32000, 32063, 31743, 0, 2, 31872])
    want = np.array([0.99607843, 0.99607843, -0.99607843, -0.99607843, -0.99607843, 0.99607843, 0.99607843,
                     0.99607843, 0.0])
    got = O.detokenize(ids)
    assert np.allclose(got, want, atol=1e-8)
    assert g["known_ids"] == ids.tolist()
    assert np.array_equal(got, np.array(g["known_actions"]))           # bit-exact float64


def test_detokenizer_all_ids_bit_exact_vs_reference():
    g = _tok()
    got = O.detokenize(np.arange(0, 32064))
    assert hashlib.sha256(np.ascontiguousarray(got, dtype=np.float64).tobytes()).hexdigest() == g["decode_all_sha256"]
    assert np.array_equal(got[g["decode_tail_first_id"]:], np.array(g["decode_tail"]))
    assert np.array_equal(got[:: g["decode_stride"]], np.array(g["decode_strided"]))


def test_bins_and_digitize_round_trip():
    g = _tok()
    bins, centers = O.action_bins(256)
    assert np.array_equal(bins, np.array(g["bins"])) and np.array_equal(centers, np.array(g["bin_centers"]))
    assert g["action_token_begin_idx"] == 31743 and len(centers) == 255
    assert np.digitize(np.array(g["digitize_in"]), bins).tolist() == g["digitize_out"]
    assert np.digitize(np.array([-1, -0.999, 0, 0.999, 1]), bins).tolist() == [1, 1, 128, 255, 256]
    # encode -> decode lands in the bin that holds the value (action_tokenizer.py:38-47 then :49-68)
    x = np.linspace(-1, 1, 1001)
    ids = 32000 - np.digitize(np.clip(x, -1, 1), bins)
    assert np.all(np.abs(O.detokenize(ids) - x) <= (bins[1] - bins[0]) + 1e-12)


def test_unnormalize_formula():
    stats = {"q01": [-1.0, 0.0, 2.0], "q99": [1.0, 4.0, 3.0], "mask": [True, True, False]}
    n = np.array([0.5, -1.0, 0.25])
    got = O.unnormalize(n, stats)
    assert np.array_equal(got, np.array([0.5, 0.0, 0.25]))
    no_mask = O.unnormalize(n, {"q01": [-1.0, 0.0, 2.0], "q99": [1.0, 4.0, 3.0]})   # default mask all True (:528)
    assert np.array_equal(no_mask, np.array([0.5, 0.0, 0.5 * 1.25 * 1.0 + 2.0]))


def _wiring():
    return torch.load(os.path.join(G, "hf_wiring_golden.pt"), weights_only=False)


def _dims_from(fx):
    towers = tuple(O.TowerDims(*t) for t in fx["dims"]["towers"])
    ld, li, ll, lh = fx["dims"]["llm"]
    return O.VLADims(image_size=fx["dims"]["image_size"], towers=towers, llm_dim=ld, llm_inter=li, llm_layers=ll,
                     llm_heads=lh)


def test_state_dict_names_match_reference_model():
    fx = _wiring()
    d = _dims_from(fx)
    assert sorted(O.weight_shapes(d).keys()) == fx["state_dict_names"]


def test_forward_matches_reference_modeling_prismatic():
    """Oracle multimodal forward == the reference's PrismaticForConditionalGeneration.forward (fp32, tiny dims)."""
    fx = _wiring()
    d = _dims_from(fx)
    W = O.make_weights(d, seed=fx["weight_seed"], dtype=torch.float32)
    ids, px = O.make_inputs(d, fx["B"], prompt_len=fx["P"], seed=fx["input_seed"])
    with torch.no_grad():
        out = O.multimodal_forward(W, d, ids, px.float(), dtype=torch.float32)
    assert len(out.hidden_states) == fx["n_hidden_states"] == d.llm_layers + 1
    assert out.logits.shape[1] == fx["seq_len"] == d.n_patches + fx["P"]
    tol = dict(rtol=2e-4, atol=2e-5)
    assert torch.allclose(out.projector_features.mean(1), fx["projector_features_mean"], **tol)
    assert torch.allclose(out.projector_features[:, 7], fx["projector_features_row7"], **tol)
    pooled = torch.stack([h.float().mean(1) for h in out.hidden_states])
    last = torch.stack([h[:, -1].float() for h in out.hidden_states])
    assert torch.allclose(pooled, fx["hidden_pooled"], **tol)
    assert torch.allclose(last, fx["hidden_last_token"], rtol=5e-4, atol=5e-5)
    assert torch.allclose(out.logits[:, -1], fx["logits_last"], rtol=5e-4, atol=5e-5)


def test_predict_action_matches_reference():
    """29871 append + greedy tokens + de-tokenise + un-normalise == reference predict_action (modeling_prismatic.py:506-536)."""
    fx = _wiring()
    d = _dims_from(fx)
    W = O.make_weights(d, seed=fx["weight_seed"], dtype=torch.float32)
    ids, px = O.make_inputs(d, fx["B"], prompt_len=fx["P"], seed=fx["input_seed"])
    with torch.no_grad():
        actions, tokens = O.predict_action(W, d, ids, px.float(), fx["norm_stats"], dtype=torch.float32,
                                           return_tokens=True)
    assert tokens[0].tolist() == fx["generated_sequence"][0, -7:].tolist()
    assert np.array_equal(actions[0], fx["action"].numpy())
