"""bs=1 latency breakdown on a B200: per-kernel-family device time (library profiler) vs wall time of one pass."""
import ctypes as C
import dataclasses
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic_inputs  # noqa: E402
from openvla_probe_b200 import _lib, config as cfgmod, weights  # noqa: E402
from openvla_probe_b200.modeling_prismatic import OpenVLAForActionPrediction  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
lib = _lib.load()
stats = {"synthetic": {"action": {"q01": [0.0] * 7, "q99": [1.0] * 7}}}
cfg = dataclasses.replace(cfgmod.openvla_7b(), norm_stats=stats)
model = OpenVLAForActionPrediction(cfg, max_batch=B, max_prompt_len=32)
weights.bind_random(model)
ids, px = synthetic_inputs(cfg, B, 31, 1)
ids = torch.cat([ids, torch.full((B, 1), 29871)], 1).cuda()
px = px.cuda()
pool_len = 256 + 31
for _ in range(3):
    model.engine.run(ids, px, pool_len, 0, 7)
torch.cuda.synchronize()
wall = []
for _ in range(10):
    t0 = time.perf_counter()
    model.engine.run(ids, px, pool_len, 0, 7)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    wall.append(((t1 - t0) * 1e3, (time.perf_counter() - t0) * 1e3))
print("host-issue ms / total ms per pass:", [f"{a:.1f}/{b:.1f}" for a, b in wall])
n = (C.c_longlong * 7)(); ms = (C.c_double * 7)(); fl = (C.c_double * 7)(); by = (C.c_double * 7)()
lib.ovla_profile_enable(1)
model.engine.run(ids, px, pool_len, 0, 7)
torch.cuda.synchronize()
lib.ovla_profile_enable(0)
lib.ovla_profile_collect(n, ms, fl, by)
names = ["gemm_tcgen05", "gemv", "flash_attn", "decode_attn", "norm", "pool", "other"]
tot = 0.0
for i, nm in enumerate(names):
    tot += ms[i]
    print(f"{nm:14s} launches={n[i]:5d} ms={ms[i]:8.3f} TF/s={fl[i] / max(ms[i], 1e-9) / 1e9:8.1f} GB/s={by[i] / max(ms[i], 1e-9) / 1e6:8.1f}")
print("sum of kernel ms:", tot)

# phase split in graph mode: prefill only (n_new=0), +lm_head/argmax (n_new=1), full (n_new=7)
for n_new in (0, 1, 7):
    for _ in range(3):
        model.engine.run(ids, px, pool_len, 0, n_new)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); model.engine.run(ids, px, pool_len, 0, n_new); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"n_new={n_new}: p50 {sorted(ts)[5]:.2f} ms")
