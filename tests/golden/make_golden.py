"""Generate the golden fixtures that pin the CPU oracle against the REFERENCE's own code.

Run in the build container (needs /root/reference; the fixtures it writes are committed, the tests never read
/root/reference):

    python tests/golden/make_golden.py

Fixtures
  action_tokenizer_golden.json  -- the reference's `prismatic/vla/action_tokenizer.py` (loaded by file path, unmodified):
                                   decode of every token id 0..32063, digitize round trip, bin tables.
  hf_wiring_golden.pt           -- the reference's `prismatic/extern/hf/modeling_prismatic.py` (unmodified) executed in
                                   fp32 on a tiny architecture: `OpenVLAForActionPrediction.forward` (vision split, concat
                                   order, projector, splice, LLM call, hidden_states) and `.predict_action` (29871 append,
                                   de-tokenise, un-normalise).  Two third-party pieces are absent here and substituted:
                                   `timm` by tests/golden/_timm_stub.py (a restatement of the timm-0.9.10 ViT surface), and
                                   transformers-4.40.1 `generate` by a plain greedy loop that drives the reference's own
                                   `forward` branches (transformers 5.5's generate no longer matches the reference's
                                   `prepare_inputs_for_generation` contract).  The LLM is the installed transformers
                                   `LlamaForCausalLM`, created by the reference via AutoModelForCausalLM.from_config.
                                   Weights are `oracle.make_weights(seed)` loaded with strict=True, which also pins the
                                   HF state-dict names.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import openvla_oracle as O  # noqa: E402


def _load(modname, path):
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def golden_dims():
    """Tiny towers at the reference's fixed 224-px resolution (VISION_BACKBONE_TO_RESOLUTION), tiny Llama."""
    dino = O.TowerDims(128, 3, 2, 512, 5, True)
    sig = O.TowerDims(144, 4, 2, 536, 0, False)
    return O.VLADims(image_size=224, patch=14, towers=(dino, sig), llm_dim=256, llm_inter=704, llm_layers=2,
                     llm_heads=2)


def make_action_tokenizer_golden():
    at = _load("ref_action_tokenizer", f"{REF}/prismatic/vla/action_tokenizer.py")

    class Tok:
        vocab_size = 32000

    t = at.ActionTokenizer(Tok())
    ids = np.arange(0, 32064)
    dec = t.decode_token_ids_to_actions(ids)
    probe = np.array([-1.0, -0.999, -0.5, 0.0, 0.3, 0.999, 1.0])
    out = {
        "action_token_begin_idx": int(t.action_token_begin_idx),
        "n_bins": int(t.n_bins),
        "bins": t.bins.tolist(),
        "bin_centers": t.bin_centers.tolist(),
        # full table pinned by digest; the interesting tail (action tokens + padding rows) and a strided sample verbatim
        "decode_all_sha256": __import__("hashlib").sha256(np.ascontiguousarray(dec, dtype=np.float64).tobytes()).hexdigest(),
        "decode_tail_first_id": 31700,
        "decode_tail": dec[31700:].tolist(),
        "decode_stride": 997,
        "decode_strided": dec[::997].tolist(),
        "digitize_in": probe.tolist(),
        "digitize_out": np.digitize(np.clip(probe, -1.0, 1.0), t.bins).tolist(),
        "known_ids": [31744, 31745, 31999, 32000, 32063, 31743, 0, 2, 31872],
        "known_actions": t.decode_token_ids_to_actions(
            np.array([31744, 31745, 31999, 32000, 32063, 31743, 0, 2, 31872])).tolist(),
    }
    with open(os.path.join(HERE, "action_tokenizer_golden.json"), "w") as f:
        json.dump(out, f)
    print("wrote action_tokenizer_golden.json:", len(dec), "decoded ids")


def make_hf_wiring_golden():
    import _timm_stub

    _timm_stub.install()
    for name in ["prismatic", "prismatic.extern", "prismatic.extern.hf"]:   # skip prismatic/__init__.py (needs draccus)
        m = types.ModuleType(name)
        m.__path__ = [f"{REF}/" + name.replace(".", "/")]
        sys.modules[name] = m
    cfgm = _load("prismatic.extern.hf.configuration_prismatic", f"{REF}/prismatic/extern/hf/configuration_prismatic.py")
    mm = _load("prismatic.extern.hf.modeling_prismatic", f"{REF}/prismatic/extern/hf/modeling_prismatic.py")

    d = golden_dims()
    ids_timm = cfgm.VISION_BACKBONE_TO_TIMM_ID["dinosiglip-vit-so-224px"]
    for tid, t in zip(ids_timm, d.towers):
        _timm_stub.TOWERS[tid] = dict(dim=t.dim, depth=t.depth, heads=t.heads, mlp=t.mlp, n_prefix=t.n_prefix,
                                      layerscale=t.layerscale, patch=d.patch)
    stats = O.default_stats()
    norm_stats = {"synthetic": {"action": stats}}
    cfg = cfgm.OpenVLAConfig(
        vision_backbone_id="dinosiglip-vit-so-224px", llm_backbone_id="llama2-7b-pure",
        arch_specifier="no-align+fused-gelu-mlp", norm_stats=norm_stats,
        text_config=dict(hidden_size=d.llm_dim, intermediate_size=d.llm_inter, num_hidden_layers=d.llm_layers,
                         num_attention_heads=d.llm_heads, num_key_value_heads=d.llm_heads, vocab_size=d.vocab,
                         rms_norm_eps=d.rms_eps, pad_token_id=d.pad_token_id, max_position_embeddings=2048),
    )
    cfg._attn_implementation = "eager"
    torch.manual_seed(0)

    class Ref(mm.OpenVLAForActionPrediction):
        """API-drift shim only: transformers 5.5 calls tie_weights(recompute_mapping=...); Llama-2 ties nothing
        (reference comment at modeling_prismatic.py:277).  No arithmetic is touched."""

        def tie_weights(self, *a, **k):
            return None

    model = Ref(cfg).float().eval()
    W = O.make_weights(d, seed=3, dtype=torch.float32)
    missing, unexpected = model.load_state_dict(W, strict=False)
    # the only tolerated extras are non-persistent / buffer-like entries of the installed transformers
    assert not unexpected, unexpected
    missing = [m for m in missing if "rotary_emb" not in m]
    assert not missing, missing

    B, P = 1, 9
    input_ids, pixel_values = O.make_inputs(d, B, prompt_len=P, seed=4)
    pixel_values = pixel_values.float()
    with torch.no_grad():
        out = model(input_ids=input_ids, attention_mask=torch.ones_like(input_ids), pixel_values=pixel_values,
                    output_hidden_states=True, output_projector_features=True, return_dict=True)

        def greedy(ids, max_new_tokens, pixel_values=None, attention_mask=None, do_sample=False, **kw):
            """Greedy search (transformers 4.40.1 semantics, EOS=2 stop) driving the reference's forward branches."""
            o = model(input_ids=ids, attention_mask=torch.ones_like(ids), pixel_values=pixel_values, use_cache=True,
                      return_dict=True)
            seq, past = ids, o.past_key_values
            for step in range(max_new_tokens):
                nxt = o.logits[:, -1].argmax(-1, keepdim=True)
                seq = torch.cat([seq, nxt], 1)
                if int(nxt) == 2 or step == max_new_tokens - 1:
                    break
                o = model(input_ids=nxt, past_key_values=past, use_cache=True, return_dict=True)
                past = o.past_key_values
            return seq

        model.generate = greedy
        action = model.predict_action(input_ids, unnorm_key="synthetic", pixel_values=pixel_values,
                                      attention_mask=torch.ones_like(input_ids), do_sample=False)
        ids29 = torch.cat([input_ids, torch.tensor([[29871]])], 1)
        seq = greedy(ids29, 7, pixel_values=pixel_values)
    fx = {
        "dims": {"image_size": d.image_size, "towers": [list(vars(t).values()) for t in d.towers],
                 "llm": [d.llm_dim, d.llm_inter, d.llm_layers, d.llm_heads]},
        "weight_seed": 3, "input_seed": 4, "B": B, "P": P,
        "state_dict_names": sorted(W.keys()),
        "projector_features_mean": out.projector_features.mean(1),            # [B, D]
        "projector_features_row7": out.projector_features[:, 7].clone(),
        "hidden_pooled": torch.stack([h.float().mean(1) for h in out.hidden_states]),   # [L+1, B, D]
        "hidden_last_token": torch.stack([h[:, -1].float() for h in out.hidden_states]),
        "logits_last": out.logits[:, -1].float().clone(),                      # [B, V]
        "n_hidden_states": len(out.hidden_states),
        "seq_len": out.logits.shape[1],
        "generated_sequence": seq.clone(),
        "action": torch.from_numpy(np.asarray(action)),
        "norm_stats": stats,
    }
    torch.save(fx, os.path.join(HERE, "hf_wiring_golden.pt"))
    print("wrote hf_wiring_golden.pt: seq_len", fx["seq_len"], "tokens", seq[0, -7:].tolist(), "action", action)


if __name__ == "__main__":
    make_action_tokenizer_golden()
    make_hf_wiring_golden()
