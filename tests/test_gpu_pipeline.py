"""GPU parity: the fused CUDA pass (through the C ABI) against the CPU oracle on shared seeded weights.

Tolerance policy (written here): for every checked tensor, our bf16 result must stay within
c = 2x the error envelope of the oracle's own bf16 evaluation against the oracle in fp32 (rel-L2 and max-abs,
plus a small absolute floor); token ids are compared exactly where the oracle's bf16 and fp32 runs agree on them.
"""
import numpy as np
import pytest
import torch

from helpers import envelope_ok, pair, rel_l2, to_f32
from oracle import openvla_oracle as O

pytestmark = pytest.mark.gpu


def _build(kind="tiny", fused=True, B=3, P=12, llm_layers=2, depth=(3, 3), seed=0):
    from openvla_probe_b200.modeling_prismatic import from_state_dict

    od, pc = pair(kind, fused=fused, llm_layers=llm_layers, depth=depth)
    W = O.make_weights(od, seed=seed)
    stats = {"synthetic": {"action": O.default_stats()}}
    import dataclasses

    pc = dataclasses.replace(pc, norm_stats=stats)
    model = from_state_dict(pc, W, max_batch=max(B, 1), max_prompt_len=P + 2)
    ids, px = O.make_inputs(od, B, prompt_len=P, seed=seed + 1)
    return od, pc, W, model, ids, px


@pytest.mark.parametrize("fused", [True, False])
def test_forward_hidden_states_and_logits(fused):
    od, pc, W, model, ids, px = _build(fused=fused, B=2, P=10)
    out = model.forward(input_ids=ids.cuda(), pixel_values=px.cuda(), output_hidden_states=True,
                        output_projector_features=True)
    with torch.no_grad():
        ref32 = O.multimodal_forward(to_f32(W), od, ids, px, dtype=torch.float32)
        ref16 = O.multimodal_forward(W, od, ids, px, dtype=torch.bfloat16)
    assert len(out.hidden_states) == od.llm_layers + 1
    ok, info = envelope_ok(out.projector_features.float().cpu(), ref32.projector_features, ref16.projector_features.float())
    assert ok, f"projector: {info}"
    for i in range(od.llm_layers + 1):
        ok, info = envelope_ok(out.hidden_states[i].float().cpu(), ref32.hidden_states[i], ref16.hidden_states[i].float())
        assert ok, f"hidden[{i}]: {info}"
    ok, info = envelope_ok(out.logits.cpu(), ref32.logits, ref16.logits)
    assert ok, f"logits: {info}"
    # hidden_states[0] is the spliced embedding: [BOS | patches | text[1:]] -- exact copy semantics
    emb = O.embed(W, ids, torch.bfloat16)
    hs0 = out.hidden_states[0].cpu()
    assert torch.equal(hs0[:, 0], emb[:, 0]) and torch.equal(hs0[:, od.n_patches + 1:], emb[:, 1:])


def test_vision_backbone_patches():
    od, pc, W, model, ids, px = _build(B=2, P=6)
    r = model.engine.run(ids, px, 0, 0, 0, want_patches=True)
    with torch.no_grad():
        ref32 = O.vision_backbone(to_f32(W), od, px.float())
        ref16 = O.vision_backbone(W, od, px)
    ok, info = envelope_ok(r["patches"].float().cpu(), ref32, ref16.float())
    assert ok, info
    # concat order DINO || SigLIP (modeling_prismatic.py:123)
    d0 = od.towers[0].dim
    assert rel_l2(r["patches"][..., :d0].float().cpu(), ref32[..., :d0]) < 0.05
    assert rel_l2(r["patches"][..., d0:].float().cpu(), ref32[..., d0:]) < 0.05


@pytest.mark.parametrize("pooling", ["mean", "final"])
def test_capture_and_action_single_pass_equals_two_pass(pooling):
    od, pc, W, model, ids, px = _build(B=3, P=9)
    layers = list(range(od.llm_layers + 1))
    embeds, actions = model.predict_action_and_capture(ids, unnorm_key="synthetic", layer_indices=layers + [-1],
                                                       pooling_method=pooling, pixel_values=px)
    stats = O.default_stats()
    with torch.no_grad():
        e32, a32 = O.get_vla_action(to_f32(W), od, ids, px, stats, layers, pooling, dtype=torch.float32)
        e16, a16 = O.get_vla_action(W, od, ids, px, stats, layers, pooling, dtype=torch.bfloat16)
    assert actions.shape == (3, 7) and actions.dtype == np.float64
    for L in layers:
        assert embeds[L].shape == (3, od.llm_dim) and embeds[L].dtype == np.float32
        ok, info = envelope_ok(embeds[L], e32[L], e16[L])
        assert ok, f"pooled layer {L} ({pooling}): {info}"
    assert np.array_equal(embeds[-1], embeds[od.llm_layers])


def test_tokens_match_oracle_given_identical_logits_and_agreement_rate():
    od, pc, W, model, ids, px = _build(B=4, P=8)
    ids29 = torch.cat([ids, torch.full((4, 1), 29871)], 1)
    r = model.engine.run(ids29, px, 0, 0, 7, want_logits=True)
    tokens = r["tokens"].cpu()
    logits = r["step_logits"].cpu()           # [7, B, V]
    # bit-exact argmax given identical logits (lowest index on ties)
    assert torch.equal(tokens, torch.argmax(logits, dim=-1).t())
    with torch.no_grad():
        seq16, lg16, _ = O.greedy_generate(W, od, ids29, px, 7, dtype=torch.bfloat16, stop_on_eos=False)
        seq32, lg32, _ = O.greedy_generate(to_f32(W), od, ids29, px, 7, dtype=torch.float32, stop_on_eos=False)
    # first generated token: same prefill => envelope on the logits
    ok, info = envelope_ok(logits[0], lg32[0], lg16[0])
    assert ok, info
    agree = (tokens == seq16[:, -7:]).float().mean().item()
    print(f"action-token agreement vs bf16 oracle: {agree:.3f}")


def test_batch_invariance():
    od, pc, W, model, ids, px = _build(B=4, P=8)
    (a_all, t_all), pooled_all = model._predict(ids, "synthetic", capture=True, pixel_values=px, return_tokens=True)
    for b in (0, 3):
        (a_b, t_b), pooled_b = model._predict(ids[b:b + 1], "synthetic", capture=True, pixel_values=px[b:b + 1],
                                              return_tokens=True)
        assert np.array_equal(t_all[b], t_b[0])
        assert np.array_equal(a_all[b], a_b)
        assert np.allclose(pooled_all[:, b], pooled_b[:, 0], rtol=0, atol=1e-6)


def test_host_and_device_entry_points_agree():
    od, pc, W, model, ids, px = _build(B=2, P=8)
    a_host, p_host = model._predict(ids, "synthetic", capture=True, pixel_values=px)
    a_dev, p_dev = model._predict(ids.cuda(), "synthetic", capture=True, pixel_values=px.cuda())
    assert np.array_equal(a_host, a_dev) and np.array_equal(p_host, p_dev)


def test_host_path_streams_pooled_states_and_fills_caller_buffer():
    """B above the CUDA-graph limit (16) takes the eager pass, where ovla_run_host copies each layer's pooled block to
    the host behind its pooling kernel; with `pooled_out=` the block lands in the caller's (pinned) buffer and the
    returned arrays are views of it.  Repeated calls and the device entry point give identical bits."""
    od, pc, W, model, ids, px = _build(B=20, P=8)
    tc = pc.text_config
    a_dev, p_dev = model._predict(ids.cuda(), "synthetic", capture=True, pixel_values=px.cuda())
    out = torch.full((tc.num_hidden_layers + 1, 20, tc.hidden_size), float("nan"), dtype=torch.float32).pin_memory()
    for _ in range(3):
        out.fill_(float("nan"))
        a_host, p_host = model._predict(ids, "synthetic", capture=True, pixel_values=px, pooled_out=out)
        assert p_host.ctypes.data == out.data_ptr()                       # views of the caller's buffer
        assert np.array_equal(a_host, a_dev) and np.array_equal(p_host, p_dev)
    a2, p2 = model._predict(ids, "synthetic", capture=True, pixel_values=px)      # fresh-result path
    assert np.array_equal(p2, p_dev) and p2.ctypes.data != out.data_ptr()
    embeds, _ = model.predict_action_and_capture(ids, unnorm_key="synthetic", layer_indices=[0, -1], pixel_values=px,
                                                 pooled_out=out)
    assert np.array_equal(embeds[-1], p_dev[-1]) and np.array_equal(embeds[0], p_dev[0])
    with pytest.raises(ValueError):
        model._predict(ids, "synthetic", capture=True, pixel_values=px, pooled_out=torch.zeros(3, 3))


def test_error_behaviour():
    from openvla_probe_b200 import _lib

    od, pc, W, model, ids, px = _build(B=2, P=8)
    with pytest.raises(AssertionError):
        model.predict_action(ids, unnorm_key="nope", pixel_values=px)
    with pytest.raises(ValueError):
        model.predict_action(ids, unnorm_key="synthetic", pixel_values=px[:1])
    with pytest.raises(_lib.OvlaError):
        big = torch.cat([ids] * 3)
        model.predict_action(big, unnorm_key="synthetic", pixel_values=torch.cat([px] * 3))   # > max_batch
    bad = ids.clone()
    bad[0, 3] = 40000
    with pytest.raises(_lib.OvlaError):
        model.predict_action(bad, unnorm_key="synthetic", pixel_values=px)
    # empty batch
    a = model.predict_action(ids[:0], unnorm_key="synthetic", pixel_values=px[:0])
    assert a.shape == (0, 7)


def test_full_width_shapes_against_fp32_oracle():
    """Real layer widths / 224-px frames / T = 288 (depth cut to 3 ViT blocks + 2 Llama layers so the CPU oracle
    finishes in seconds).  Stated tolerance vs the fp32 oracle: rel-L2 <= 1.5e-2 on pooled hidden states and on the
    vision / projector outputs, <= 3e-2 on last-position logits (bf16 pipeline, ~30 rounding points deep)."""
    od, pc, W, model, ids, px = _build(kind="full-width", B=2, P=31, llm_layers=2)
    ids29 = torch.cat([ids, torch.full((2, 1), 29871)], 1)
    r = model.engine.run(ids29, px, od.n_patches + 31, 0, 2, want_logits=True, want_patches=True, want_projector=True)
    with torch.no_grad():
        W32 = to_f32(W)
        patches = O.vision_backbone(W32, od, px.float())
        out = O.multimodal_forward(W32, od, ids29, px, dtype=torch.float32)
    assert rel_l2(r["patches"].float().cpu(), patches) < 1.5e-2
    assert rel_l2(r["projector"].float().cpu(), out.projector_features) < 1.5e-2
    pooled = r["pooled"].cpu()
    for i, h in enumerate(out.hidden_states):
        ref = h[:, : od.n_patches + 31].mean(1)
        assert rel_l2(pooled[i], ref) < 1.5e-2, i
    assert rel_l2(r["step_logits"][0].cpu(), out.logits[:, -1]) < 3e-2
    assert torch.equal(r["tokens"].cpu(), torch.argmax(r["step_logits"].cpu(), -1).t())


@pytest.mark.parametrize("B", [1, 2, 3])
def test_persistent_decode_step_equals_the_per_layer_kernels(B):
    """Cached decode steps at batch <= 4 run as ONE persistent kernel (csrc/decode_mega.cu: shared-memory weight rings
    fed by cp.async.bulk across grid barriers, RMSNorm recomputed while staging, attention inside).  Same summation
    orders and rounding points as the per-layer kernels: every step's logits, the greedy ids and the pooled states are
    bit-identical between the two, eagerly and under CUDA-graph replay; real Llama widths (D 4096, I 11008, 32 heads
    of 128, vocab 32064), 3 layers, T = 288."""
    od, pc, W, model, ids, px = _build(kind="full-width", B=B, P=31, llm_layers=3)
    ids29 = torch.cat([ids, torch.full((B, 1), 29871)], 1)
    eng = model.engine
    res = {}
    for mega in (0, 1):
        eng.set_option("decode_mega", mega)
        runs = [eng.run(ids29, px, od.n_patches + 31, 0, 7, want_logits=True) for _ in range(3)]   # eager, capture, replay
        torch.cuda.synchronize()
        for r in runs[1:]:
            assert torch.equal(r["step_logits"], runs[0]["step_logits"]) and torch.equal(r["tokens"], runs[0]["tokens"])
        res[mega] = {k: v.cpu() for k, v in runs[0].items()}
    assert bool(torch.isfinite(res[1]["step_logits"]).all())
    for s in range(7):
        assert torch.equal(res[0]["step_logits"][s], res[1]["step_logits"][s]), f"decode step {s}"
    assert torch.equal(res[0]["tokens"], res[1]["tokens"]) and torch.equal(res[0]["pooled"], res[1]["pooled"])
    assert torch.equal(res[1]["tokens"], torch.argmax(res[1]["step_logits"], -1).t())
    eng.close()


def test_full_size_7b_size_independent_properties():
    """BASELINE config [2] at full size (24 + 27 ViT blocks, 32 Llama layers, 224 px, real widths; random-init weights
    as in bench.py -- the CPU oracle cannot run this in test time), checked through properties that do not need it:
    run-to-run determinism (bit-identical), batch invariance of the captured states (a row computed alone -- CUDA-graph
    replay, GEMV decode, split-K GEMMs -- against the same row inside a batch of three, an eager batch of 20 and the benchmark's batch of 256),
    host and device entry points agreeing bit for bit, pooling semantics, and actions inside the un-normalisation box."""
    import dataclasses

    from bench import synthetic_inputs
    from openvla_probe_b200 import config as cfgmod, weights
    from openvla_probe_b200.modeling_prismatic import OpenVLAForActionPrediction

    stats = {"synthetic": {"action": {"q01": [-1.0] * 7, "q99": [2.0] * 7}}}
    cfg = dataclasses.replace(cfgmod.openvla_7b(), norm_stats=stats)
    model = OpenVLAForActionPrediction(cfg, max_batch=256, max_prompt_len=24)
    weights.bind_random(model)
    ids256, px256 = synthetic_inputs(cfg, 256, 20, 7)
    ids20, px20 = ids256[:20].contiguous(), px256[:20].contiguous()
    ids, px = ids20[:3].contiguous(), px20[:3].contiguous()
    L, D = cfg.text_config.num_hidden_layers, cfg.text_config.hidden_size

    (a3, t3), p3 = model._predict(ids.cuda(), "synthetic", capture=True, pixel_values=px.cuda(), return_tokens=True)
    (a3b, t3b), p3b = model._predict(ids.cuda(), "synthetic", capture=True, pixel_values=px.cuda(), return_tokens=True)
    assert p3.shape == (L + 1, 3, D) and np.isfinite(p3).all()
    assert np.array_equal(t3, t3b) and np.array_equal(p3, p3b) and np.array_equal(a3, a3b)          # deterministic
    (ah, th), ph = model._predict(ids, "synthetic", capture=True, pixel_values=px, return_tokens=True)
    assert np.array_equal(th, t3) and np.array_equal(ph, p3)                                         # host == device entry
    assert ((a3 >= -1.0 - 1e-9) & (a3 <= 2.0 + 1e-9)).all() and a3.shape == (3, 7)

    for b in range(3):
        for _ in range(3):                                   # third call replays the captured CUDA graph
            (a1, t1), p1 = model._predict(ids[b:b + 1].cuda(), "synthetic", capture=True, pixel_values=px[b:b + 1].cuda(),
                                          return_tokens=True)
        for layer in (0, 1, 8, 16, 24, L):
            err = np.linalg.norm(p1[layer, 0] - p3[layer, b]) / np.linalg.norm(p3[layer, b])
            assert err < 2e-2, (b, layer, err)               # same math, different tiling / reduction order in bf16
    # the eager large-batch pass (no CUDA graph, one stream, pooled states streamed to the host layer by layer)
    (a20, t20), p20 = model._predict(ids20, "synthetic", capture=True, pixel_values=px20, return_tokens=True)
    assert p20.shape == (L + 1, 20, D) and np.isfinite(p20).all()
    for b in range(3):
        for layer in (0, 1, 8, 16, 24, L):
            err = np.linalg.norm(p20[layer, b] - p3[layer, b]) / np.linalg.norm(p3[layer, b])
            assert err < 2e-2, ("batch 20 vs 3", b, layer, err)
    # the benchmark's own shape: 256 observations per pass (CTA-pair 256x256 tiles everywhere, M = 256 decode GEMMs)
    (a256, t256), p256 = model._predict(ids256, "synthetic", capture=True, pixel_values=px256, return_tokens=True)
    assert p256.shape == (L + 1, 256, D) and np.isfinite(p256).all()
    for b in range(3):
        for layer in (0, 1, 8, 16, 24, L):
            err = np.linalg.norm(p256[layer, b] - p3[layer, b]) / np.linalg.norm(p3[layer, b])
            assert err < 2e-2, ("batch 256 vs 3", b, layer, err)
    # (greedy ids are not compared across batch sizes: over random-weight logits near-ties flip between tilings and the
    # continuation then diverges; argmax / de-tokenisation are checked bit-exactly given identical logits elsewhere)
    assert a256.shape == (256, 7) and ((a256 >= -1.0 - 1e-9) & (a256 <= 2.0 + 1e-9)).all()
    # layer 0 is the embedding stream (BOS | projected patches | text): pooling fewer rows changes it, "final" != "mean"
    _, p_final = model._predict(ids.cuda(), "synthetic", capture=True, pooling_method="final", pixel_values=px.cuda())
    assert not np.allclose(p_final[0], p3[0]) and np.isfinite(p_final).all()
    model.engine.close()


def test_full_size_siglip_single_backbone_properties():
    """BASELINE config [4] at full size (SigLIP SO400M only, 2-layer projector, 32 Llama layers): the single-backbone
    wiring (modeling_prismatic.py:135-137,148-150) with head_dim-72 attention -- determinism and a row alone (CUDA-graph
    replay) against the same row in a batch."""
    import dataclasses

    from bench import synthetic_inputs
    from openvla_probe_b200 import config as cfgmod, weights
    from openvla_probe_b200.modeling_prismatic import OpenVLAForActionPrediction

    stats = {"synthetic": {"action": {"q01": [0.0] * 7, "q99": [1.0] * 7}}}
    cfg = dataclasses.replace(cfgmod.siglip_7b(), norm_stats=stats)
    model = OpenVLAForActionPrediction(cfg, max_batch=4, max_prompt_len=24)
    weights.bind_random(model)
    ids, px = synthetic_inputs(cfg, 4, 18, 3)
    assert px.shape[1] == 3                                              # one tower: 3 channels
    L = cfg.text_config.num_hidden_layers
    ids_d, px_d = ids.cuda(), px.cuda()                                  # same pointers every call => call 2 captures + replays
    (a4, t4), p4 = model._predict(ids_d, "synthetic", capture=True, pixel_values=px_d, return_tokens=True)
    assert np.isfinite(p4).all(), "non-finite pooled states"
    for rep in range(3):                                                 # eager, captured + replayed, replayed
        (a4b, t4b), p4b = model._predict(ids_d, "synthetic", capture=True, pixel_values=px_d, return_tokens=True)
        assert np.isfinite(p4b).all(), f"non-finite pooled states (repeat {rep})"
        assert np.array_equal(t4, t4b), f"token ids differ between repeats ({rep})"
        diff = [(layer, int((p4[layer] != p4b[layer]).sum()), float(np.abs(p4[layer] - p4b[layer]).max()))
                for layer in range(L + 1) if not np.array_equal(p4[layer], p4b[layer])]
        assert not diff, f"repeat {rep}: pooled states differ run to run; (layer, n_diff, max_abs) first = {diff[:3]}"
    for b in (0, 3):
        for _ in range(3):
            (a1, t1), p1 = model._predict(ids[b:b + 1].cuda(), "synthetic", capture=True, pixel_values=px[b:b + 1].cuda(),
                                          return_tokens=True)
        for layer in (0, 1, 16, L):
            err = np.linalg.norm(p1[layer, 0] - p4[layer, b]) / np.linalg.norm(p4[layer, b])
            assert err < 2e-2, (b, layer, err)
    model.engine.close()


def test_full_size_single_observation_against_fp32_oracle():
    """BASELINE config [1]/[2] at FULL size -- every ViT block, 32 Llama layers, real widths, 224-px frame, T = 277 --
    for one observation against the CPU oracle in fp32 on identical (bf16-representable) weights.  Stated tolerance:
    rel-L2 <= 1.5e-2 on each of the 33 mean-pooled hidden states, <= 2.5e-2 on the vision / projector outputs (a bf16
    pipeline ~300 rounding points deep against fp32), <= 6e-2 on the last-position logits (measured: 0.16-0.81 % on the
    pooled states growing with depth, 1.3 % vision / projector, 2.9 % logits); the 6 cached decode steps are compared
    the same way, teacher-forced with the device's own tokens.  Weights are drawn on the GPU and
    copied to the host (30 GB fp32); skipped when the host has less than 48 GB of free memory."""
    import dataclasses

    import psutil

    if psutil.virtual_memory().available < 48e9:
        pytest.skip("needs 48 GB of free host memory for the fp32 oracle weights")
    from openvla_probe_b200.modeling_prismatic import from_state_dict

    od, pc = pair("openvla-7b")
    pc = dataclasses.replace(pc, norm_stats={"synthetic": {"action": O.default_stats()}})
    g = torch.Generator(device="cuda").manual_seed(1234)
    W = {}
    for name, shape in O.weight_shapes(od).items():        # same init rules as O.make_weights, drawn on the device
        affine = name.endswith(("norm1.weight", "norm2.weight", "layernorm.weight", "model.norm.weight", "scale_factor"))
        w = torch.randn(shape, device="cuda", generator=g)
        w = (1.0 + 0.1 * w) if affine else 0.02 * w
        if name.endswith("embed_tokens.weight"):
            w[od.pad_token_id] = 0.0
        W[name] = w.to(torch.bfloat16).float().cpu()         # bf16-representable fp32 on the host
    model = from_state_dict(pc, W, max_batch=1, max_prompt_len=24)
    ids, px = O.make_inputs(od, 1, prompt_len=20, seed=3)
    ids29 = torch.cat([ids, torch.full((1, 1), 29871)], 1)
    pool_len = od.n_patches + 20
    r = model.engine.run(ids29, px, pool_len, 0, 7, want_logits=True, want_patches=True, want_projector=True)
    torch.set_num_threads(max(1, (__import__("os").cpu_count() or 1)))
    with torch.no_grad():
        patches = O.vision_backbone(W, od, px.float())
        out = O.multimodal_forward(W, od, ids29, px, dtype=torch.float32)
    errs = {"patches": rel_l2(r["patches"].float().cpu(), patches),
            "projector": rel_l2(r["projector"].float().cpu(), out.projector_features)}
    pooled = r["pooled"].cpu()
    for i, h in enumerate(out.hidden_states):
        errs[f"pooled[{i}]"] = rel_l2(pooled[i], h[:, :pool_len].mean(1))
    errs["logits"] = rel_l2(r["step_logits"][0].cpu(), out.logits[:, -1])
    # the 6 cached decode steps (GEMV linears, fused RoPE + KV append + attention), teacher-forced with OUR tokens so that
    # a near-tie flip cannot desynchronise the two sequences
    toks = r["tokens"].cpu()                                  # [1, 7]
    step_logits = r["step_logits"].cpu()                      # [7, 1, V]
    assert torch.equal(toks, torch.argmax(step_logits, -1).t())
    past = out.past_key_values
    with torch.no_grad():
        for k in range(1, 7):
            o = O.cached_forward(W, od, toks[:, k - 1:k], past, torch.float32)
            past = o.past_key_values
            errs[f"decode_logits[{k}]"] = rel_l2(step_logits[k], o.logits[:, -1])
            top2 = torch.topk(o.logits[0, -1], 2).values
            if float(top2[0] - top2[1]) > 4 * float((step_logits[k, 0] - o.logits[0, -1]).abs().max()):
                assert int(toks[0, k]) == int(torch.argmax(o.logits[0, -1]))     # clear margin -> same greedy token
    print("full-size rel-L2 vs fp32 oracle:", {k: round(v, 5) for k, v in errs.items()})
    assert len(out.hidden_states) == 33
    for k, v in errs.items():
        assert v < (6e-2 if "logits" in k else 1.5e-2 if k.startswith("pooled") else 2.5e-2), (k, v)
    model.engine.close()


# ----------------------------------------------------------------------------------------------- ragged prompts
def _ragged(ids, lens, pad):
    """Right-pad: row b keeps its first lens[b] ids; returns (padded ids, attention mask)."""
    ids = ids.clone()
    mask = torch.zeros_like(ids)
    for b, n in enumerate(lens):
        ids[b, n:] = pad
        mask[b, :n] = 1
    return ids, mask


@pytest.mark.parametrize("B,pooling,entry", [(3, "mean", "host"), (4, "final", "device"), (6, "mean", "host"),
                                             (12, "mean", "device")])
def test_ragged_batch_rows_equal_their_batch1_results(B, pooling, entry):
    """Right-padded prompts of different lengths (attention_mask = ones then zeros; reference splice:
    modeling_prismatic.py:388-390, SURVEY Appendix B "pool over each sample's true length").  Row b of the batch must
    be what a batch-1 call on its un-padded prompt returns: greedy ids and actions exactly, pooled states to 1e-6 (as
    test_batch_invariance), and inside the oracle's error envelope.  B = 3 / 4 decode in the persistent kernel, B = 6
    with the weight-streaming GEMVs + the per-layer decode attention, B = 12 with tcgen05 GEMMs."""
    P = 12
    od, pc, W, model, ids, px = _build(B=B, P=P)
    lens = [P, 5, 9, 1, 7, 12, 3, 2, 11, 6, 4, 8][:B]
    rid, mask = _ragged(ids, lens, pc.pad_token_id)
    layers = list(range(od.llm_layers + 1))
    dev = (lambda t: t.cuda()) if entry == "device" else (lambda t: t)
    for rep in range(3):                                     # eager, graph capture, graph replay
        (a_all, t_all), pooled_all = model._predict(dev(rid), "synthetic", capture=True, pooling_method=pooling,
                                                    pixel_values=dev(px), attention_mask=mask, return_tokens=True)
        if rep == 0:
            first = (a_all.copy(), t_all.copy(), pooled_all.copy())
        else:
            assert np.array_equal(first[0], a_all) and np.array_equal(first[1], t_all) and np.array_equal(first[2], pooled_all)
    stats = O.default_stats()
    for b in range(B):
        one = ids[b:b + 1, :lens[b]]
        (a_b, t_b), pooled_b = model._predict(dev(one), "synthetic", capture=True, pooling_method=pooling,
                                              pixel_values=dev(px[b:b + 1]), return_tokens=True)
        assert np.array_equal(t_all[b], t_b[0]), (b, lens[b])
        assert np.array_equal(a_all[b], a_b), (b, lens[b])
        assert np.allclose(pooled_all[:, b], pooled_b[:, 0], rtol=0, atol=1e-6), (b, lens[b])
    for b in (1, 3 % B):                                     # oracle envelope on two short rows
        one = ids[b:b + 1, :lens[b]]
        with torch.no_grad():
            e32, _ = O.get_vla_action(to_f32(W), od, one, px[b:b + 1], stats, layers, pooling, dtype=torch.float32)
            e16, _ = O.get_vla_action(W, od, one, px[b:b + 1], stats, layers, pooling, dtype=torch.bfloat16)
        for L in layers:
            ok, info = envelope_ok(pooled_all[L, b:b + 1], e32[L], e16[L])
            assert ok, f"row {b} (len {lens[b]}) layer {L}: {info}"


def test_ragged_lengths_are_read_at_replay_time():
    """The lengths live in a device buffer the captured pass reads when it runs: the same CUDA graph replayed with other
    lengths (same shapes, host entry point => same staging pointers) gives the other lengths' results."""
    B, P = 3, 10
    od, pc, W, model, ids, px = _build(B=B, P=P)
    before = model.engine.lib.ovla_graph_replays(model.engine._h)
    results = {}
    patterns = [(10, 4, 7), (3, 10, 5), (10, 4, 7), (6, 6, 10)]
    for lens in patterns:
        rid, mask = _ragged(ids, lens, pc.pad_token_id)
        (a, t), pooled = model._predict(rid, "synthetic", capture=True, pixel_values=px, attention_mask=mask,
                                        return_tokens=True)
        if lens in results:
            assert np.array_equal(results[lens][0], t) and np.array_equal(results[lens][1], pooled)
        results[lens] = (t.copy(), pooled.copy())
    assert model.engine.lib.ovla_graph_replays(model.engine._h) - before >= 2
    for lens in patterns[1:2] + patterns[3:]:
        for b in range(B):
            (a_b, t_b), pooled_b = model._predict(ids[b:b + 1, :lens[b]], "synthetic", capture=True,
                                                  pixel_values=px[b:b + 1], return_tokens=True)
            assert np.array_equal(results[lens][0][b], t_b[0])
            assert np.allclose(results[lens][1][:, b], pooled_b[:, 0], rtol=0, atol=1e-6)


def test_ragged_full_width_against_fp32_oracle():
    """Real layer widths, T = 288 for the longest row: each ragged row against the fp32 oracle run on its un-padded
    prompt (rel-L2 <= 1.5e-2 on the pooled states, <= 3e-2 on its first-token logits, as the uniform full-width test),
    and the decode positions / KV appends of the short rows through the equality of their greedy ids with a batch-1
    run of the engine."""
    P = 31
    od, pc, W, model, ids, px = _build(kind="full-width", B=3, P=P, llm_layers=2)
    lens = [P, 12, 23]
    rid, _ = _ragged(ids, lens, pc.pad_token_id)
    ids29 = torch.cat([rid, torch.full((3, 1), pc.pad_token_id)], 1)
    for b, n in enumerate(lens):
        ids29[b, n] = 29871
    lens29 = torch.tensor([n + 1 for n in lens], dtype=torch.int32)
    r = model.engine.run(ids29, px, od.n_patches + P, 0, 7, want_logits=True, prompt_lens=lens29)
    pooled, logits0, tokens = r["pooled"].cpu(), r["step_logits"][0].cpu(), r["tokens"].cpu()
    for b, n in enumerate(lens):
        one29 = ids29[b:b + 1, :n + 1]
        with torch.no_grad():
            out = O.multimodal_forward(to_f32(W), od, one29, px[b:b + 1], dtype=torch.float32)
        for i, h in enumerate(out.hidden_states):
            assert rel_l2(pooled[i, b:b + 1], h[:, : od.n_patches + n].mean(1)) < 1.5e-2, (b, i)
        assert rel_l2(logits0[b:b + 1], out.logits[:, -1]) < 3e-2, b
        r1 = model.engine.run(one29, px[b:b + 1], od.n_patches + n, 0, 7)
        assert torch.equal(r1["tokens"].cpu()[0], tokens[b]), b


# ----------------------------------------------------------------------------------------------- RMSNorm fused into the GEMMs
@pytest.mark.parametrize("fused_towers", [True, False])
def test_prefill_with_rmsnorm_fused_into_the_gemms(fused_towers, monkeypatch):
    """Large-batch prefill: o_proj / down_proj leave per-row sums of squares from their epilogues and the QKV / gate-up
    GEMMs (norm weight folded into W) apply 1/rms to their accumulators -- no stand-alone RMSNorm kernel between the
    layers (LlamaRMSNorm behind modeling_prismatic.py:404-415).  The threshold is lowered so the tiny model takes the
    path; checked against the oracle's error envelope on every hidden state / pooled state / the logits, against the
    un-fused kernels of the same engine, and by the launch count (2 norm launches per layer gone, 1 sum-of-squares
    kernel added)."""
    monkeypatch.setenv("OVLA_FUSE_NORM_MIN_ROWS", "1")
    od, pc, W, model, ids, px = _build(fused=fused_towers, B=3, P=10, llm_layers=3)
    monkeypatch.delenv("OVLA_FUSE_NORM_MIN_ROWS")
    eng, lib = model.engine, model.engine.lib
    eng.set_option("graph_max_batch", 0)
    ids29 = torch.cat([ids, torch.full((3, 1), 29871)], 1)
    res, launches = {}, {}
    for fuse in (1, 0):
        eng.set_option("fuse_norm", fuse)
        lib.ovla_reset_launch_count()
        res[fuse] = eng.run(ids29, px, od.n_patches + 10, 0, 3, want_hidden=True, want_logits=True)
        torch.cuda.synchronize()
        launches[fuse] = int(lib.ovla_launch_count())
    assert launches[0] - launches[1] == 2 * od.llm_layers - 1, launches
    with torch.no_grad():
        ref32 = O.multimodal_forward(to_f32(W), od, ids29, px, dtype=torch.float32)
        ref16 = O.multimodal_forward(W, od, ids29, px, dtype=torch.bfloat16)
    hid = res[1]["hidden"].float().cpu()
    for i in range(od.llm_layers + 1):
        ok, info = envelope_ok(hid[i], ref32.hidden_states[i], ref16.hidden_states[i].float())
        assert ok, f"hidden[{i}]: {info}"
        assert rel_l2(hid[i], res[0]["hidden"][i].float().cpu()) < 1e-2, i
    pooled = res[1]["pooled"].cpu()
    for i in range(od.llm_layers + 1):
        ok, info = envelope_ok(pooled[i], ref32.hidden_states[i][:, : od.n_patches + 10].mean(1),
                               ref16.hidden_states[i][:, : od.n_patches + 10].float().mean(1))
        assert ok, f"pooled[{i}]: {info}"
    ok, info = envelope_ok(res[1]["step_logits"][0].cpu(), ref32.logits[:, -1], ref16.logits[:, -1])
    assert ok, f"logits: {info}"
    # repeatable bit for bit (the partial sums are plain stores in a fixed layout)
    eng.set_option("fuse_norm", 1)
    again = eng.run(ids29, px, od.n_patches + 10, 0, 3, want_hidden=True)
    assert torch.equal(again["hidden"], res[1]["hidden"]) and torch.equal(again["tokens"], res[1]["tokens"])


def test_get_vla_action_with_the_device_image_transform():
    """get_vla_action on a simulator-sized frame (256 x 256) with the PrismaticImageProcessor mirror doing the resize /
    normalize on the device: same action and embeddings, bit for bit, as feeding the frame the image oracle resized on
    the host through the plain path (the resize is bit-identical to PIL, the rest is the same engine)."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import image_cases as IC
    from oracle import image_oracle as IO
    from openvla_probe_b200.openvla_utils import SyntheticProcessor, get_vla_action
    from openvla_probe_b200.processing_prismatic import openvla_image_processor

    od, pc, W, model, ids, px = _build(B=1, P=9)
    frame = IC.frame(256, 256, 21)
    ip = openvla_image_processor(pc)
    assert ip.image_resize_strategy == "resize-naive" and ip.size == pc.image_size
    layers = list(range(od.llm_layers + 1))
    e1, a1 = get_vla_action(model, SyntheticProcessor(pc, prompt_len=9, image_processor=ip), "openvla", {"full_image": frame},
                            "pick up the bowl", "synthetic", layer_indices=layers, return_embeddings=True)
    small = IO.transform_u8(frame, "resize-naive", pc.image_size)
    e2, a2 = get_vla_action(model, SyntheticProcessor(pc, prompt_len=9), "openvla", {"full_image": small},
                            "pick up the bowl", "synthetic", layer_indices=layers, return_embeddings=True)
    assert a1.shape == (7,) and np.array_equal(a1, a2)
    for L in layers:
        assert np.array_equal(e1[L], e2[L]), L
