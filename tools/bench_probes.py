#!/usr/bin/env python
"""Probe-training benchmark (BASELINE.json configs[3]): object linear probe and dual-head probe on synthetic 4096-d
layer features, batch 4096, K = 439 kept labels of 481, AdamW; 1 GPU or N GPUs (torchrun, NCCL gradient allreduce).

    python tools/bench_probes.py [--steps 50] [--kind object|dual]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_probes.py

Prints one JSON line per kind on rank 0: optimisation steps/s (max over ranks, CUDA events), per-family kernel times
from the library profiler (TF32 tcgen05 GEMM TFLOP/s; HBM kernels GB/s), and a CPU baseline = the reference's step
(`nn.Linear` + BCEWithLogitsLoss + torch AdamW, train_object_probes.py:184-189) on the host cores for a few steps.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def synth(N, D, L, seed):
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(N, D, generator=g)
    Y = torch.randint(0, 2, (N, L), generator=g).to(torch.int8)
    Y[torch.rand(N, L, generator=g) < 0.5] = -1                       # p_missing ~ 0.5 (label_stats.csv)
    keep = torch.arange(L)[: 439 if L >= 439 else L]
    return X, Y, keep


def cpu_step_time(kind, X, Y, keep, batch, steps=3):
    torch.set_num_threads(os.cpu_count() or 1)
    K, D = len(keep), X.shape[1]
    pw = torch.ones(K) * 2.0
    heads = [torch.nn.Linear(D, K) for _ in range(2 if kind == "dual" else 1)]
    opt = torch.optim.AdamW([p for h in heads for p in h.parameters()], lr=1e-3, weight_decay=1e-4)
    bce = torch.nn.BCEWithLogitsLoss(reduction="none", pos_weight=pw)
    ts = []
    for s in range(steps + 1):
        xb, yb = X[s * batch:(s + 1) * batch], Y[s * batch:(s + 1) * batch][:, keep]
        t0 = time.time()
        mask = (yb != -1)
        if kind == "dual":
            lp = torch.nn.functional.binary_cross_entropy_with_logits(heads[0](xb), mask.float(), pos_weight=torch.tensor(1.0))
            lt = (torch.nn.functional.binary_cross_entropy_with_logits(heads[1](xb), (yb == 1).float(), reduction="none")
                  * mask.float()).sum() / mask.sum()
            loss = lp + lt
        else:
            loss = (bce(heads[0](xb), (yb == 1).float()) * mask.float()).sum() / mask.sum()
        opt.zero_grad(); loss.backward(); opt.step()
        if s > 0:
            ts.append(time.time() - t0)
    return sum(ts) / len(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--kind", default="both", choices=["object", "dual", "both"])
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    import torch.distributed as dist

    from openvla_probe_b200 import _lib
    from openvla_probe_b200.probes import ProbeTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    D, L = 4096, 481
    n_steps_total = args.steps + args.warmup
    N = args.batch * 8
    X, Y, keep = synth(N, D, L, 0)
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
        if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    for kind in (["object", "dual"] if args.kind == "both" else [args.kind]):
        K = len(keep)
        pw = torch.tensor(1.7) if kind == "dual" else torch.ones(K) * 2.0
        tr = ProbeTrainer(kind, D, K, pw, batch=args.batch, device=local)
        Xd, Yd = X.cuda(), Y.cuda()
        perm = torch.randperm(N, generator=torch.Generator().manual_seed(1))
        tr.load_epoch(Xd, Yd, keep, perm, drop_last=True)
        n_ep = len(tr.steps)
        for s in range(args.warmup):
            tr.train_step(s % n_ep)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        lib.ovla_profile_enable(1)
        lib.ovla_reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(args.steps):
            tr.train_step(s % n_ep)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        lib.ovla_profile_enable(0)
        n = (C.c_longlong * 7)(); t = (C.c_double * 7)(); fl = (C.c_double * 7)(); by = (C.c_double * 7)()
        lib.ovla_profile_collect(n, t, fl, by)
        ms_step = float(ms.item()) / args.steps
        if rank == 0:
            heads = 2 if kind == "dual" else 1
            rows = heads * tr.Kpad
            flops = 2 * 2.0 * (args.batch / world) * D * rows                      # forward + dW, per rank
            line = {
                "metric": "probe optimisation steps/s (batch 4096, D=4096, K=439)", "kind": kind, "value": 1e3 / ms_step,
                "unit": "steps/s", "samples_per_s": args.batch * 1e3 / ms_step, "n_gpus": world, "steps": args.steps,
                "ms_per_step": ms_step, "dtype": "tf32 GEMM / fp32 elsewhere", "data": "synthetic",
                "gpu_launches": int(lib.ovla_launch_count()),
                "gemm_tf32": {"ms_per_step": t[0] / args.steps, "tflops": fl[0] / max(t[0], 1e-9) / 1e9,
                              "launches_per_step": n[0] / args.steps},
                "hbm_kernels": {"ms_per_step": t[6] / args.steps, "gbs": by[6] / max(t[6], 1e-9) / 1e6,
                                "frac_of_hbm_peak": by[6] / max(t[6], 1e-9) / 1e6 / peaks["hbm_gbs"]},
                "step_tflops_per_gpu": flops / (ms_step * 1e-3) / 1e12,
                "allreduce_bytes_per_step": int(tr.G.numel() * 4) if world > 1 else 0,
                "loss": tr.step_loss(),
            }
            if not args.no_cpu and world == 1:
                cpu_s = cpu_step_time(kind, X, Y, keep, args.batch)
                line["cpu_baseline"] = {"value": 1.0 / cpu_s, "unit": "steps/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": "3 optimisation steps of the reference's torch fp32 step on the host"}
            print(json.dumps(line), flush=True)
        del tr
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
