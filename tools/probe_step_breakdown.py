#!/usr/bin/env python
"""Per-launch time of the grouped probe step (33 layers, batch 4096, D 4096, K 439): each of the four launches timed
alone with CUDA events over `--iters` back-to-back repetitions (inputs 4.4 GB per GEMM, far beyond L2).

    python tools/probe_step_breakdown.py [--kind object|dual] [--layers 33]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="object")
    ap.add_argument("--layers", type=int, default=33)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    from openvla_probe_b200 import _lib
    from openvla_probe_b200.probes import _KIND0, MultiLayerProbeTrainer

    G, B, D, K = args.layers, args.batch, 4096, 439
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn(G, 2 * B, D, generator=g, device="cuda")
    Y = (torch.rand(2 * B, 481, generator=g, device="cuda") < 0.5).to(torch.int8)
    Y[torch.rand(2 * B, 481, generator=g, device="cuda") < 0.5] = -1
    torch.manual_seed(0)
    tr = MultiLayerProbeTrainer(args.kind, G, D, K, torch.tensor(1.7) if args.kind == "dual" else torch.ones(K) * 2, batch=B)
    tr.load_epoch(X, Y, torch.arange(K), torch.arange(2 * B), True)
    del X
    f4, st, ck, lib = 4, _lib.stream_ptr(), _lib.check, tr.lib
    lo, n = 0, B

    def fwd():
        tr._grouped_gemm(tr.Xp.data_ptr() + lo * D * f4, D, tr.n_alloc * D, tr.P.data_ptr(), D, tr.n_total, G, n, tr.rows, D,
                         tr.Z.data_ptr(), tr.rows, tr.bmax * tr.rows, tr.P.data_ptr() + tr.n_w * f4, tr.n_total)

    def bce():
        ck(lib.ovla_probe_bce_grad_grouped(tr.Z.data_ptr(), tr.rows, tr.bmax * tr.rows, tr.Yp.data_ptr(), n, K, tr.Kpad,
                                           _KIND0[args.kind], tr.heads, tr.pw_vec.data_ptr() if tr.pw_vec is not None else None,
                                           tr.pw_scalar, tr.dZT.data_ptr(), tr.ldz_t, tr.rows * tr.ldz_t, G, tr.Gbuf.data_ptr(),
                                           tr.gs, tr.n_w, tr.n_total, tr.part.data_ptr(), tr.isplits, tr.ticket.data_ptr(), st))

    def dw():
        tr._grouped_gemm(tr.dZT.data_ptr(), tr.ldz_t, tr.rows * tr.ldz_t, tr.XpT.data_ptr() + lo * f4, tr.n_alloc,
                         D * tr.n_alloc, G, tr.rows, D, n, tr.Gbuf.data_ptr(), D, tr.gs)

    def adamw():
        tr._adamw_chunk(0, G, 1)

    out = {}
    for name, fn in (("gemm_logits", fwd), ("bce_grad_db_stats", bce), ("gemm_dW", dw), ("adamw", adamw)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[name] = {"us_per_layer": e0.elapsed_time(e1) * 1e3 / args.iters / G}
    rows = tr.rows
    flops = 2.0 * B * D * rows
    for k in ("gemm_logits", "gemm_dW"):
        out[k]["tf32_tflops"] = flops / (out[k]["us_per_layer"] * 1e-6) / 1e12
    out["gemm_logits"]["gbs"] = 4.0 * (B * D + rows * D + B * rows) / (out["gemm_logits"]["us_per_layer"] * 1e-6) / 1e9
    out["gemm_dW"]["gbs"] = 4.0 * (B * D + rows * D + B * rows) / (out["gemm_dW"]["us_per_layer"] * 1e-6) / 1e9
    out["bce_grad_db_stats"]["gbs"] = (8.0 * B * rows + B * tr.Kpad) / (out["bce_grad_db_stats"]["us_per_layer"] * 1e-6) / 1e9
    out["adamw"]["gbs"] = 28.0 * tr.n_total / (out["adamw"]["us_per_layer"] * 1e-6) / 1e9
    out["sum_us_per_layer"] = sum(v["us_per_layer"] for v in out.values() if isinstance(v, dict))
    print(json.dumps({"kind": args.kind, "layers": G, "batch": B, **out}))


if __name__ == "__main__":
    main()
