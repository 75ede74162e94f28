"""Oracle self-consistency: the third-party halves it restates, and the invariants the CUDA path relies on."""
import numpy as np
import pytest
import torch

from oracle import openvla_oracle as O


@pytest.fixture(scope="module")
def tiny():
    d = O.tiny_dims()
    return d, O.make_weights(d, seed=0, dtype=torch.float32)


def test_llama_restatement_matches_installed_transformers():
    """Pins oracle.llama_forward (+ KV-cache decode, hidden_states semantics) to transformers' LlamaForCausalLM."""
    from transformers import LlamaConfig, LlamaForCausalLM

    d = O.VLADims(towers=O.tiny_dims().towers, image_size=56, llm_dim=256, llm_inter=704, llm_layers=3, llm_heads=2,
                  vocab=320, pad_token_id=300)
    cfg = LlamaConfig(hidden_size=256, intermediate_size=704, num_hidden_layers=3, num_attention_heads=2,
                      num_key_value_heads=2, vocab_size=320, rms_norm_eps=1e-6, pad_token_id=300,
                      attn_implementation="eager")
    torch.manual_seed(0)
    m = LlamaForCausalLM(cfg).float().eval()
    W = {"language_model." + k: v.detach().clone() for k, v in m.state_dict().items()}
    x = torch.randn(2, 11, 256)
    with torch.no_grad():
        ref = m(inputs_embeds=x, output_hidden_states=True, use_cache=True, return_dict=True)
        hs, last, kv = O.llama_forward(W, d, x)
        assert len(hs) == len(ref.hidden_states) == 4
        for a, b in zip(hs, ref.hidden_states):
            assert torch.allclose(a, b, rtol=1e-4, atol=1e-5)
        assert torch.allclose(O.lm_head(W, last), ref.logits, rtol=1e-4, atol=1e-4)
        # hidden_states[L] is post-final-norm
        assert torch.allclose(hs[-1], ref.hidden_states[-1], atol=1e-5)
        # cached single-token step: position = cache length (modeling_prismatic.py:330-341)
        tok = torch.tensor([[5], [7]])
        ref2 = m(input_ids=tok, past_key_values=ref.past_key_values, use_cache=True, return_dict=True)
        e = torch.nn.functional.embedding(tok, W["language_model.model.embed_tokens.weight"])
        _, last2, _ = O.llama_forward(W, d, e, kv)
        assert torch.allclose(O.lm_head(W, last2), ref2.logits, rtol=1e-4, atol=1e-4)


def test_vit_restatement_matches_timm_style_module():
    """oracle.vit_tower == the nn.Module restatement used to run the reference (tests/golden/_timm_stub.py):
    block selection depth-2, prefix strip, no final norm (modeling_prismatic.py:85-87,119-123)."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import _timm_stub as T

    d = O.tiny_dims()
    W = O.make_weights(d, seed=2, dtype=torch.float32)
    img = torch.randn(2, 3, d.image_size, d.image_size)
    for ti, t in enumerate(d.towers):
        vt = T.VisionTransformer(d.image_size, d.patch, t.dim, t.depth, t.heads, t.mlp, t.n_prefix, t.layerscale)
        sd = {}
        for k, v in W.items():
            if k.startswith(O.TOWER_PREFIX[ti] + "."):
                sd[k[len(O.TOWER_PREFIX[ti]) + 1:].replace("scale_factor", "gamma")] = v
        vt.load_state_dict(sd, strict=True)
        with torch.no_grad():
            ref = vt.get_intermediate_layers(img, n={t.depth - 2})[0]
            got = O.vit_tower(W, O.TOWER_PREFIX[ti], t, d, img)
        assert got.shape == (2, d.n_patches, t.dim)
        assert torch.allclose(got, ref, rtol=1e-4, atol=1e-5)


def test_splice_order_and_hidden_state_semantics(tiny):
    d, W = tiny
    ids, px = O.make_inputs(d, 2, prompt_len=7)
    with torch.no_grad():
        out = O.multimodal_forward(W, d, ids, px.float(), dtype=torch.float32)
    emb = O.embed(W, ids, torch.float32)
    h0 = out.hidden_states[0]
    T = d.n_patches + 7
    assert h0.shape == (2, T, d.llm_dim) and len(out.hidden_states) == d.llm_layers + 1
    assert torch.equal(h0[:, 0], emb[:, 0])                         # BOS first (:383-385)
    assert torch.equal(h0[:, 1:1 + d.n_patches], out.projector_features)
    assert torch.equal(h0[:, 1 + d.n_patches:], emb[:, 1:])
    # last entry is post-final-norm
    pre = out.hidden_states[-1]
    assert torch.allclose(pre.pow(2).mean(-1).sqrt(), (W["language_model.model.norm.weight"].pow(2).mean().sqrt()).expand(2, T),
                          rtol=0.2)


def test_single_pass_equals_two_pass(tiny):
    """F9: hidden states at positions [0, T-1) of the predict pass (29871 appended) equal the capture pass."""
    d, W = tiny
    ids, px = O.make_inputs(d, 2, prompt_len=7)
    ids29 = torch.cat([ids, torch.full((2, 1), 29871)], 1)
    with torch.no_grad():
        cap = O.multimodal_forward(W, d, ids, px.float(), dtype=torch.float32)
        pred = O.multimodal_forward(W, d, ids29, px.float(), dtype=torch.float32)
    for a, b in zip(cap.hidden_states, pred.hidden_states):
        assert torch.allclose(a, b[:, :-1], rtol=0, atol=2e-6)


def test_batch_invariance(tiny):
    d, W = tiny
    ids, px = O.make_inputs(d, 3, prompt_len=6)
    st = O.default_stats()
    with torch.no_grad():
        a_all, t_all = O.predict_action(W, d, ids, px.float(), st, dtype=torch.float32, return_tokens=True)
        a1, t1 = O.predict_action(W, d, ids[1:2], px[1:2].float(), st, dtype=torch.float32, return_tokens=True)
    assert np.array_equal(t_all[1], t1[0]) and np.array_equal(a_all[1], a1[0])


def test_token_29871_append_rule(tiny):
    d, W = tiny
    ids, px = O.make_inputs(d, 2, prompt_len=6)
    with_tok, _ = O.make_inputs(d, 2, prompt_len=7, append_empty=True)
    assert with_tok.shape[1] == 7 and bool((with_tok[:, -1] == 29871).all())
    st = O.default_stats()
    mixed = with_tok.clone()
    mixed[0, -1] = 5        # not ALL rows end in 29871 -> appended to every row (batch-wide test, :512)
    with torch.no_grad():
        s1, _, _ = O.greedy_generate(W, d, torch.cat([mixed, torch.full((2, 1), 29871)], 1), px.float(), 2,
                                     dtype=torch.float32)
    assert s1.shape[1] == 7 + 1 + 2


def test_eos_policy_matches_hf_generate(tiny, monkeypatch):
    """If id 2 (EOS) is generated early, HF greedy search stops and `generated_ids[0, -7:]` (modeling_prismatic.py:521)
    reaches back into the prompt.  EOS is forced by biasing its logit."""
    d, W = tiny
    real_lm_head = O.lm_head

    def biased(Wd, x):
        lg = real_lm_head(Wd, x)
        lg[..., 2] += 1e4
        return lg

    monkeypatch.setattr(O, "lm_head", biased)
    ids, px = O.make_inputs(d, 1, prompt_len=9)
    st = O.default_stats()
    with torch.no_grad():
        seq, logits, _ = O.greedy_generate(W, d, torch.cat([ids, torch.tensor([[29871]])], 1), px.float(), 7,
                                           dtype=torch.float32)
        acts, toks = O.predict_action(W, d, ids, px.float(), st, dtype=torch.float32, return_tokens=True)
    assert seq.shape[1] == 11 and int(seq[0, 10]) == 2                # stopped after one generated token
    assert toks[0].tolist() == torch.cat([ids[0], torch.tensor([29871, 2])])[-7:].tolist()
    assert np.array_equal(acts[0], O.unnormalize(O.detokenize(toks[0], d), st))


def test_pool_modes_and_layer_indices(tiny):
    d, W = tiny
    ids, px = O.make_inputs(d, 1, prompt_len=6)
    st = O.default_stats()
    with torch.no_grad():
        e_mean, _ = O.get_vla_action(W, d, ids, px.float(), st, [0, -1, d.llm_layers], "mean", dtype=torch.float32)
        e_fin, _ = O.get_vla_action(W, d, ids, px.float(), st, None, "final", dtype=torch.float32)
        out = O.multimodal_forward(W, d, ids, px.float(), dtype=torch.float32)
    assert np.array_equal(e_mean[-1], e_mean[d.llm_layers])
    assert list(e_fin.keys()) == [-1]                                  # default layers (-1,), openvla_utils.py:198
    assert np.allclose(e_fin[-1][0], out.hidden_states[-1][0, -1].numpy())
    assert np.allclose(e_mean[0][0], out.hidden_states[0][0].mean(0).numpy(), atol=1e-6)


def test_tiny_dims_fp64_cross_check(tiny):
    d, W = tiny
    ids, px = O.make_inputs(d, 1, prompt_len=5)
    W64 = {k: v.double() for k, v in W.items()}
    with torch.no_grad():
        o32 = O.multimodal_forward(W, d, ids, px.float(), dtype=torch.float32)
        o64 = O.multimodal_forward(W64, d, ids, px.double(), dtype=torch.float64)
    for a, b in zip(o32.hidden_states, o64.hidden_states):
        assert float((a.double() - b).abs().max()) < 1e-3 * float(b.abs().max())


def test_center_crop_restatement_properties():
    """center_crop input branch (openvla_utils.py:81-124,155-175): the full box at native size is the identity, a
    constant frame stays constant, the crop is centred (commutes with a 180 degree rotation) and a 0.9-area crop of a
    horizontal ramp keeps the ramp's centre value at the centre."""
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (2, 224, 224, 3), dtype=np.uint8)
    assert np.array_equal(O.center_crop_frames(img, 1.0), img)
    assert np.unique(O.center_crop_frames(np.full((1, 256, 256, 3), 201, np.uint8), 0.9)).tolist() == [201]
    a = O.center_crop_frames(img, 0.9)
    b = O.center_crop_frames(img[:, ::-1, ::-1].copy(), 0.9)[:, ::-1, ::-1]
    assert np.abs(a.astype(int) - b.astype(int)).max() <= 1          # float32 lerp order differs by at most one code
    ramp = np.broadcast_to(np.arange(256, dtype=np.uint8)[None, None, :, None], (1, 256, 256, 3)).copy()
    out = O.center_crop_frames(ramp, 0.9)
    side = np.sqrt(np.float32(0.9))
    assert abs(int(out[0, 100, 0, 0]) - (1 - side) / 2 * 255) <= 1 and abs(int(out[0, 100, 223, 0]) - (1 + side) / 2 * 255) <= 1



def _tower_sd(W, prefix):
    return {k[len(prefix) + 1:]: v for k, v in W.items() if k.startswith(prefix + ".")}


def test_vit_towers_match_independent_transformers_implementations():
    """timm 0.9.10 (the reference's ViT code) cannot be installed here, so the oracle's tower arithmetic is pinned to
    the OTHER public implementations of the same two architectures: transformers' Dinov2WithRegistersModel
    (= vit_large_patch14_reg4_dinov2: cls + 4 register tokens, position embedding on the patches, LayerScale, erf-GELU)
    and SiglipVisionModel (= vit_so400m_patch14_siglip with hidden_act="gelu", as timm 0.9.10 runs it).  Weights are
    mapped name by name; the oracle's output (blocks 0..depth-2, prefix stripped, no final norm,
    modeling_prismatic.py:85-87,119-123) equals hidden_states[depth-1] of the HF encoders."""
    from transformers import Dinov2WithRegistersConfig, Dinov2WithRegistersModel, SiglipVisionConfig, SiglipVisionModel

    d = O.tiny_dims()
    W = O.make_weights(d, seed=5, dtype=torch.float32)
    img = torch.randn(2, 3, d.image_size, d.image_size)
    np_ = d.n_patches

    # ---- DINOv2 with registers
    t = d.towers[0]
    sd = _tower_sd(W, O.TOWER_PREFIX[0])
    cfg = Dinov2WithRegistersConfig(hidden_size=t.dim, num_hidden_layers=t.depth, num_attention_heads=t.heads,
                                    mlp_ratio=t.mlp // t.dim, hidden_act="gelu", layer_norm_eps=1e-6,
                                    image_size=d.image_size, patch_size=d.patch, num_register_tokens=4, qkv_bias=True,
                                    layerscale_value=1.0, use_swiglu_ffn=False, hidden_dropout_prob=0.0,
                                    attention_probs_dropout_prob=0.0, drop_path_rate=0.0, attn_implementation="eager")
    m = Dinov2WithRegistersModel(cfg).float().eval()
    hs = {}
    D = t.dim
    hs["embeddings.cls_token"] = sd["cls_token"]
    hs["embeddings.mask_token"] = torch.zeros(1, D)
    hs["embeddings.register_tokens"] = sd["reg_token"]
    hs["embeddings.position_embeddings"] = torch.cat([torch.zeros(1, 1, D), sd["pos_embed"]], 1)   # no_embed_class
    hs["embeddings.patch_embeddings.projection.weight"] = sd["patch_embed.proj.weight"]
    hs["embeddings.patch_embeddings.projection.bias"] = sd["patch_embed.proj.bias"]
    for i in range(t.depth):
        b, h = f"blocks.{i}.", f"encoder.layer.{i}."
        for n in ("norm1", "norm2"):
            hs[h + n + ".weight"], hs[h + n + ".bias"] = sd[b + n + ".weight"], sd[b + n + ".bias"]
        qw, kw, vw = sd[b + "attn.qkv.weight"].chunk(3, 0)
        qb, kb, vb = sd[b + "attn.qkv.bias"].chunk(3, 0)
        for nm, w_, b_ in (("query", qw, qb), ("key", kw, kb), ("value", vw, vb)):
            hs[h + f"attention.attention.{nm}.weight"], hs[h + f"attention.attention.{nm}.bias"] = w_, b_
        hs[h + "attention.output.dense.weight"], hs[h + "attention.output.dense.bias"] = sd[b + "attn.proj.weight"], sd[b + "attn.proj.bias"]
        hs[h + "layer_scale1.lambda1"], hs[h + "layer_scale2.lambda1"] = sd[b + "ls1.scale_factor"], sd[b + "ls2.scale_factor"]
        for n in ("fc1", "fc2"):
            hs[h + f"mlp.{n}.weight"], hs[h + f"mlp.{n}.bias"] = sd[b + f"mlp.{n}.weight"], sd[b + f"mlp.{n}.bias"]
    hs["layernorm.weight"], hs["layernorm.bias"] = torch.ones(D), torch.zeros(D)
    missing, unexpected = m.load_state_dict(hs, strict=False)
    assert not unexpected and not [k for k in missing if "mask_token" not in k], (missing, unexpected)
    with torch.no_grad():
        ref = m(pixel_values=img, output_hidden_states=True).hidden_states[t.depth - 1][:, t.n_prefix:]
        got = O.vit_tower(W, O.TOWER_PREFIX[0], t, d, img)
    assert got.shape == ref.shape == (2, np_, D)
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-5), float((got - ref).abs().max())

    # ---- SigLIP
    t = d.towers[1]
    sd = _tower_sd(W, O.TOWER_PREFIX[1])
    D = t.dim
    cfg = SiglipVisionConfig(hidden_size=D, intermediate_size=t.mlp, num_hidden_layers=t.depth, num_attention_heads=t.heads,
                             image_size=d.image_size, patch_size=d.patch, hidden_act="gelu", layer_norm_eps=1e-6,
                             attention_dropout=0.0, attn_implementation="eager")
    m = SiglipVisionModel(cfg).float().eval()
    hs = {"vision_model.embeddings.patch_embedding.weight": sd["patch_embed.proj.weight"],
          "vision_model.embeddings.patch_embedding.bias": sd["patch_embed.proj.bias"],
          "vision_model.embeddings.position_embedding.weight": sd["pos_embed"][0]}
    for i in range(t.depth):
        b, h = f"blocks.{i}.", f"vision_model.encoder.layers.{i}."
        hs[h + "layer_norm1.weight"], hs[h + "layer_norm1.bias"] = sd[b + "norm1.weight"], sd[b + "norm1.bias"]
        hs[h + "layer_norm2.weight"], hs[h + "layer_norm2.bias"] = sd[b + "norm2.weight"], sd[b + "norm2.bias"]
        qw, kw, vw = sd[b + "attn.qkv.weight"].chunk(3, 0)
        qb, kb, vb = sd[b + "attn.qkv.bias"].chunk(3, 0)
        for nm, w_, b_ in (("q_proj", qw, qb), ("k_proj", kw, kb), ("v_proj", vw, vb)):
            hs[h + f"self_attn.{nm}.weight"], hs[h + f"self_attn.{nm}.bias"] = w_, b_
        hs[h + "self_attn.out_proj.weight"], hs[h + "self_attn.out_proj.bias"] = sd[b + "attn.proj.weight"], sd[b + "attn.proj.bias"]
        for n in ("fc1", "fc2"):
            hs[h + f"mlp.{n}.weight"], hs[h + f"mlp.{n}.bias"] = sd[b + f"mlp.{n}.weight"], sd[b + f"mlp.{n}.bias"]
    missing, unexpected = m.load_state_dict(hs, strict=False)
    assert not unexpected and all(("post_layernorm" in k or ".head." in k) for k in missing), (missing, unexpected)
    with torch.no_grad():
        ref = m(pixel_values=img, output_hidden_states=True).hidden_states[t.depth - 1]
        got = O.vit_tower(W, O.TOWER_PREFIX[1], t, d, img)
    assert got.shape == ref.shape == (2, np_, D)
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-5), float((got - ref).abs().max())
