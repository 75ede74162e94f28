"""Phase timeline of the persistent bs=1 decode-step kernel (csrc/decode_mega.cu), from %globaltimer stamps of the first
and last CTA: per phase kind, the mean time spent staging, streaming (linear / attention) and waiting at the grid
barrier.   OVLA_MEGA_TRACE=1 python tools/decode_trace.py
(the per-piece timeline at the end needs a library built with -DOVLA_MEGA_PIECE_TRACE=1:
 python -m openvla_probe_b200.build -DOVLA_MEGA_PIECE_TRACE=1 --variant=trace; OVLA_B200_LIB=.../libovla_b200_trace.so)"""
import ctypes as C
import dataclasses
import os
import sys

import numpy as np
import torch

os.environ["OVLA_MEGA_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic_inputs  # noqa: E402
from openvla_probe_b200 import _lib, config as cfgmod, weights  # noqa: E402
from openvla_probe_b200.modeling_prismatic import OpenVLAForActionPrediction  # noqa: E402

lib = _lib.load()
stats = {"synthetic": {"action": {"q01": [0.0] * 7, "q99": [1.0] * 7}}}
cfg = dataclasses.replace(cfgmod.openvla_7b(), norm_stats=stats)
model = OpenVLAForActionPrediction(cfg, max_batch=1, max_prompt_len=32)
weights.bind_random(model)
model.engine.set_option("graph_max_batch", 0)
ids, px = synthetic_inputs(cfg, 1, 31, 1)
ids = torch.cat([ids, torch.full((1, 1), 29871)], 1).cuda()
px = px.cuda()
for _ in range(3):
    model.engine.run(ids, px, 256 + 31, 0, 7)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 2048)()
_lib.check(lib.ovla_debug_decode_trace(model.engine._h, buf, 2048))
t = np.array(buf[:], dtype=np.int64).reshape(2, 1024)
L = cfg.text_config.num_hidden_layers
# stamps per layer: [stage1 end, lin end, bar end] qkv | [attn end, bar end] | [stage, lin, bar] x 3  => see decode_mega.cu
names = ["qkv.stage", "qkv.stream", "qkv.barrier", "attn.compute", "attn.barrier", "o.stage", "o.stream", "o.barrier",
         "gate_up.stage", "gate_up.stream", "gate_up.barrier", "down.stage", "down.stream", "down.barrier"]
for who, row in (("first CTA", t[0]), ("last CTA", t[1])):
    n = 1 + L * len(names) + 2
    d = np.diff(row[:n]).astype(np.float64) / 1e3
    per = d[: L * len(names)].reshape(L, len(names))
    print(f"== {who}: whole step {(row[n - 1] - row[0]) / 1e3:.1f} us; per-layer mean (us):")
    for i, nm in enumerate(names):
        print(f"   {nm:16s} {per[:, i].mean():7.2f}  (min {per[:, i].min():6.2f}  max {per[:, i].max():6.2f})")
    print(f"   per-layer total  {per.sum(1).mean():7.2f};  lm_head stage+stream {d[L * len(names):].sum():.1f}")

# piece-level timeline of warp 0 in layer 1: (wait, consume) per 8 KB piece, in consumption order (qkv | o | gate_up | down)
for who, row in (("first CTA", t[0]), ("last CTA", t[1])):
    p = row[512:1012]
    n = int((p > 0).sum()) // 3
    if n:
        q = p[: 3 * n].reshape(n, 3).astype(np.float64) / 1e3
        print(f"== {who}, warp 0, layer 1: {n} pieces; (wait us, consume us, gap to next us)")
        print("   " + "  ".join(f"({q[i,1]-q[i,0]:.2f},{q[i,2]-q[i,1]:.2f},{(q[i+1,0]-q[i,2]) if i+1<n else 0:.2f})" for i in range(n)))
