"""Row-alone vs row-in-batch comparison at the full 7B size (debug aid)."""
import dataclasses, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic_inputs
from openvla_probe_b200 import config as cfgmod, weights
from openvla_probe_b200.modeling_prismatic import OpenVLAForActionPrediction
stats = {"synthetic": {"action": {"q01": [-1.0] * 7, "q99": [2.0] * 7}}}
cfg = dataclasses.replace(cfgmod.openvla_7b(), norm_stats=stats)
model = OpenVLAForActionPrediction(cfg, max_batch=3, max_prompt_len=24)
weights.bind_random(model)
ids, px = synthetic_inputs(cfg, 3, 20, 7)
ids29 = torch.cat([ids, torch.full((3, 1), 29871)], 1)
print("ids tail", ids[:, -3:].tolist(), "px equal rows?", bool(torch.equal(px[0], px[1])))
idd, pxd = ids29.cuda(), px.cuda()
pl = 256 + 20
rb = model.engine.run(idd, pxd, pl, 0, 2, want_patches=True, want_projector=True)
def rel(a, b): return float((a.float() - b.float()).norm() / b.float().norm())
for b in range(3):
    i1, p1 = idd[b:b + 1].contiguous(), pxd[b:b + 1].contiguous()
    for rep in range(3):
        r1 = model.engine.run(i1, p1, pl, 0, 2, want_patches=True, want_projector=True)
        print(f"b={b} rep={rep} patches {rel(r1['patches'][0], rb['patches'][b]):.4f} projector {rel(r1['projector'][0], rb['projector'][b]):.4f} "
              f"pooled0 {rel(r1['pooled'][0, 0], rb['pooled'][0, b]):.4f} pooled1 {rel(r1['pooled'][1, 0], rb['pooled'][1, b]):.4f} "
              f"pooled32 {rel(r1['pooled'][32, 0], rb['pooled'][32, b]):.4f} tokens {r1['tokens'].tolist()} vs {rb['tokens'][b].tolist()}")
