"""Action-token agreement rate of the CUDA path against the CPU oracle (north_star: "action-token agreement rate
reported"): tiny openvla architecture, random-init weights, several weight / input seeds.  Random-init logits are nearly
flat over the vocabulary, so bf16 evaluation order alone flips near-ties: the oracle's own bf16-vs-fp32 agreement is
printed beside ours as the yardstick; the first-token rate isolates the prefill from the divergence that follows a first
flipped token.  Run under gpurun; prints one JSON line."""
import dataclasses, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from helpers import pair, to_f32
from oracle import openvla_oracle as O
from openvla_probe_b200.modeling_prismatic import from_state_dict

B, P, SEEDS = 32, 10, (0, 1, 2, 3)
tot = {k: [0, 0] for k in ("ours_vs_bf16", "ours_vs_fp32", "bf16_vs_fp32", "ours_first_vs_fp32", "bf16_first_vs_fp32",
                           "ours_seq_vs_bf16", "ours_seq_vs_fp32", "bf16_seq_vs_fp32")}
act_err = []
for seed in SEEDS:
    od, pc = pair("tiny", fused=True, llm_layers=2, depth=(3, 3))
    W = O.make_weights(od, seed=seed)
    pc = dataclasses.replace(pc, norm_stats={"synthetic": {"action": O.default_stats()}})
    model = from_state_dict(pc, W, max_batch=B, max_prompt_len=P + 2)
    ids, px = O.make_inputs(od, B, prompt_len=P, seed=seed + 100)
    ids29 = torch.cat([ids, torch.full((B, 1), 29871)], 1)
    r = model.engine.run(ids29, px, 0, 0, 7)
    ours = r["tokens"].cpu()
    with torch.no_grad():
        s16, _, _ = O.greedy_generate(W, od, ids29, px, 7, dtype=torch.bfloat16, stop_on_eos=False)
        s32, _, _ = O.greedy_generate(to_f32(W), od, ids29, px, 7, dtype=torch.float32, stop_on_eos=False)
    t16, t32 = s16[:, -7:], s32[:, -7:]
    for k, a, b in (("ours_vs_bf16", ours, t16), ("ours_vs_fp32", ours, t32), ("bf16_vs_fp32", t16, t32)):
        tot[k][0] += int((a == b).sum()); tot[k][1] += a.numel()
    for k, a, b in (("ours_first_vs_fp32", ours, t32), ("bf16_first_vs_fp32", t16, t32)):
        tot[k][0] += int((a[:, 0] == b[:, 0]).sum()); tot[k][1] += B
    for k, a, b in (("ours_seq_vs_bf16", ours, t16), ("ours_seq_vs_fp32", ours, t32), ("bf16_seq_vs_fp32", t16, t32)):
        tot[k][0] += int((a == b).all(dim=1).sum()); tot[k][1] += B
    del model
print(json.dumps({"model": "tiny openvla architecture, random init", "observations": B * len(SEEDS), "tokens_per_observation": 7,
                  **{k: round(v[0] / v[1], 4) for k, v in tot.items()},
                  "note": "per-token / first-token / whole-7-token-sequence agreement; bf16_vs_fp32 = the oracle against itself"}))
