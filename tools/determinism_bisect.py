"""Run-to-run determinism bisect of one full-size pass (round-2 item 1: the eager call and the CUDA-graph replay of the
SigLIP single-backbone model at B = 4 returned different pooled states).

    python tools/determinism_bisect.py                  # parent: one child per knob setting, JSON lines on stdout
    python tools/determinism_bisect.py --child NAME     # one setting in this process

Each child runs the SAME pass `--runs` times into persistent buffers (identical pointers => run 0 eager, run 1 captures
and replays, later runs replay) and reports, against run 0, the first stage that differs
(patches -> projector -> hidden[0..L] -> logits), how many elements differ and which rows / columns they cover.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SETTINGS = {
    "default": {},
    "graphs_off": {"OVLA_GRAPHS": "0"},
    "tma_epi_off": {"OVLA_GEMM_TMA_EPI": "0"},
    "pdl_off": {"OVLA_PDL": "0"},
    "attn_tc_off": {"OVLA_ATTN_TC": "0"},
    "nofence_lib": {"OVLA_B200_LIB": os.path.join(ROOT, "openvla_probe_b200", "libovla_b200_nofence.so")},
}


def child(name: str, model_kind: str, B: int, P: int, runs: int, plain: bool = False) -> None:
    import numpy as np
    import torch

    from bench import synthetic_inputs
    from openvla_probe_b200 import _lib, config as cfgmod, weights
    from openvla_probe_b200.engine import _RunArgs
    from openvla_probe_b200.modeling_prismatic import OpenVLAForActionPrediction

    stats = {"synthetic": {"action": {"q01": [0.0] * 7, "q99": [1.0] * 7}}}
    base = cfgmod.siglip_7b() if model_kind == "siglip" else cfgmod.openvla_7b()
    cfg = dataclasses.replace(base, norm_stats=stats)
    model = OpenVLAForActionPrediction(cfg, max_batch=B, max_prompt_len=P + 6)
    weights.bind_random(model)
    ids, px = synthetic_inputs(cfg, B, P, 3)
    ids = model._append_empty(ids).cuda().contiguous()
    px = px.cuda().to(torch.bfloat16).contiguous()
    tc = cfg.text_config
    L, D, V = tc.num_hidden_layers, tc.hidden_size, tc.vocab_size
    Pn = ids.shape[1]
    T = cfg.n_patches + Pn
    dev = model.device
    buf = {
        "patches": torch.empty(B, cfg.n_patches, cfg.vision_dim, dtype=torch.bfloat16, device=dev),
        "projector": torch.empty(B, cfg.n_patches, D, dtype=torch.bfloat16, device=dev),
        "hidden": torch.empty(L + 1, B, T, D, dtype=torch.bfloat16, device=dev),
        "pooled": torch.empty(L + 1, B, D, dtype=torch.float32, device=dev),
        "logits": torch.empty(7, B, V, dtype=torch.float32, device=dev),
        "tokens": torch.empty(B, 7, dtype=torch.int64, device=dev),
    }
    a = _RunArgs()
    a.input_ids_dev, a.pixel_values_dev = ids.data_ptr(), px.data_ptr()
    a.B, a.P, a.pool_len, a.pool_mode, a.n_new_tokens = B, Pn, T - 1, 0, 7
    a.pooled_out_dev, a.tokens_out_dev = buf["pooled"].data_ptr(), buf["tokens"].data_ptr()
    if not plain:   # plain = exactly the outputs predict_action_and_capture asks for (no extra copy nodes in the pass)
        a.step_logits_out_dev, a.hidden_out_dev = buf["logits"].data_ptr(), buf["hidden"].data_ptr()
        a.projector_out_dev, a.patches_out_dev = buf["projector"].data_ptr(), buf["patches"].data_ptr()
    lib = model.engine.lib
    lib.ovla_graph_replays.restype = C.c_longlong
    snaps = []
    for r in range(runs):
        for t in buf.values():
            t.fill_(0)
        _lib.check(lib.ovla_run(model.engine._h, C.byref(a), _lib.stream_ptr()))
        torch.cuda.synchronize()
        snaps.append({k: v.clone() for k, v in buf.items()})
    out = {"setting": name, "model": model_kind, "B": B, "T": T, "runs": runs,
           "graph_replays": int(lib.ovla_graph_replays(model.engine._h)), "finite": bool(torch.isfinite(snaps[0]["pooled"]).all()),
           "pairs": []}
    stages = [("patches", None), ("projector", None)] + [("hidden", i) for i in range(L + 1)] + [("logits", None), ("tokens", None)]
    if plain:
        stages = [("pooled", i) for i in range(L + 1)] + [("tokens", None)]
    out["plain"] = plain
    for r in range(1, runs):
        rec = {"run": r, "first_diff": None}
        n_stage_diff = 0
        for key, idx in stages:
            x, y = snaps[0][key], snaps[r][key]
            if idx is not None:
                x, y = x[idx], y[idx]
            ne = (x != y)
            n = int(ne.sum())
            if n:
                n_stage_diff += 1
                if rec["first_diff"] is None:
                    x2, ne2 = x.reshape(-1, x.shape[-1]).float(), ne.reshape(-1, ne.shape[-1])
                    y2 = y.reshape(-1, y.shape[-1]).float()
                    rows = torch.nonzero(ne2.any(1)).flatten().cpu().numpy()
                    cols = torch.nonzero(ne2.any(0)).flatten().cpu().numpy()
                    rec["first_diff"] = {
                        "stage": key if idx is None else f"{key}[{idx}]", "n_diff": n, "n_total": int(ne.numel()),
                        "max_abs": float((x2 - y2).abs().max()), "n_rows": int(rows.size), "n_cols": int(cols.size),
                        "rows_head": rows[:24].tolist(), "rows_minmax": [int(rows.min()), int(rows.max())],
                        "cols_head": cols[:24].tolist(), "cols_minmax": [int(cols.min()), int(cols.max())],
                    }
        rec["stages_differing"] = n_stage_diff
        out["pairs"].append(rec)
    print("BISECT " + json.dumps(out), flush=True)
    model.engine.close()


def main() -> None:
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--child", default=None)
    ap.add_argument("--model", default="siglip")
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--prompt", type=int, default=18)
    ap.add_argument("--runs", type=int, default=5)
    ap.add_argument("--plain", action="store_true")
    ap.add_argument("--settings", default=",".join(SETTINGS))
    args = ap.parse_args()
    if args.child:
        child(args.child, args.model, args.batch, args.prompt, args.runs, args.plain)
        return
    for name in args.settings.split(","):
        env = dict(os.environ)
        env.update(SETTINGS[name])
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", name, "--model", args.model,
                            "--batch", str(args.batch), "--prompt", str(args.prompt), "--runs", str(args.runs)] +
                           (["--plain"] if args.plain else []),
                           env=env, capture_output=True, text=True, timeout=900)
        lines = [ln[7:] for ln in r.stdout.splitlines() if ln.startswith("BISECT ")]
        if lines:
            print(lines[-1], flush=True)
        else:
            print(json.dumps({"setting": name, "error": (r.stderr or r.stdout)[-1500:]}), flush=True)


if __name__ == "__main__":
    main()
