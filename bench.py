#!/usr/bin/env python
"""Benchmark of the B200-native OpenVLA predict_action + all-layer hidden-state capture path.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through the C ABI)
    python bench.py --impl reference --steps K --warmup W     # the reference algorithm on the host CPU (oracle port)

A "step" = one fused pass over one batch of synthetic observations: bs=256 per GPU (BASELINE.json configs[2]),
224-px frames, 31-token LIBERO-style prompt (+29871 => T = 288), 7 greedy action tokens, 33 mean-pooled layers.
Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement" for every field).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "actions/sec w/ all-layer hidden capture (openvla-7b predict_action, 33 pooled layers)"
UNIT = "actions/s"
N_LAYERS_CAPTURED = 33


def workload_name(config: str, batch: int, prompt_len: int) -> str:
    return (f"{config} predict_action + {N_LAYERS_CAPTURED}-layer mean-pooled capture, bs={batch}/GPU, 224px frames, "
            f"T={256 + prompt_len + 1} (256 patches + {prompt_len + 1} prompt ids), 7 greedy action tokens")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="observations per GPU per step")
    ap.add_argument("--prompt-len", type=int, default=31, help="prompt ids incl. BOS, before 29871 is appended")
    ap.add_argument("--config", default="openvla-7b", choices=["openvla-7b", "siglip-7b", "tiny"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bs1", action="store_true")
    ap.add_argument("--no-probe", action="store_true", help="skip the probe-training leg (configs[3])")
    ap.add_argument("--no-siglip", action="store_true", help="skip the single-backbone SigLIP bs=64 leg (configs[4])")
    ap.add_argument("--probe-only", action="store_true", help="run only the probe-training leg and print its block")
    ap.add_argument("--probe-kind", default="object", choices=["object", "spatial", "dual"])
    ap.add_argument("--probe-layers", type=int, default=33)
    ap.add_argument("--probe-chunks", type=int, default=0)
    ap.add_argument("--probe-comm-sms", type=int, default=16, help="SMs the grouped GEMMs leave to the collective at N > 1")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0)
    ap.add_argument("--lite", action="store_true",
                    help="profiling mode for ncu: exactly --warmup untimed steps, no e2e / bs1 / CPU legs")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- helpers
def ncu_traffic():
    """DRAM bytes per launch of the largest GEMM instance of the step (Llama gate_up, SwiGLU epilogue: 32 % of the
    step), from the committed `ncu --set full` capture; None when the summary is not in the tree."""
    import csv
    import glob
    cands = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r*_ncu_gemm_summary.csv")))
    try:
        path = cands[-1]
        rows = list(csv.reader(open(path)))
        hdr = rows[0]
        ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
        for r in rows[2:]:
            if "<256, 2, 1, 0>" in r[ik]:
                return (float(r[ir]) + float(r[iw])) * 1e9, (
                    "bytes per launch of gemm_tcgen05_kernel<256,2,SwiGLU> (M=72448 N=22016 K=4096; algorithmic 2.37e9: "
                    "A 0.59 + W 0.18 + out 1.59 GB) from profiles/" + os.path.basename(path) + "; was 8.0e9 before the "
                    "producers of the persistent grid were aligned (profiles/r02t_gemm_raster.md); what is left is W "
                    "re-read once per 16-row-tile group (18 x 0.18 GB): a wider activation slab falls out of the L2")
    except (OSError, ValueError, IndexError):
        pass
    return None, "no ncu capture in profiles/"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def synthetic_inputs(cfg, batch: int, prompt_len: int, seed: int):
    """DummyDataset-style observations (prismatic/vla/datasets/datasets.py:207): uint8 frames normalised per tower
    exactly as PrismaticImageProcessor does, bf16, channel-stacked; prompt = [BOS] + random text ids."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (batch, cfg.image_size, cfg.image_size, 3), dtype=np.uint8)
    x = torch.from_numpy(img).permute(0, 3, 1, 2).float() / 255.0
    stats = [((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)), ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))]
    if not cfg.use_fused_vision_backbone:
        stats = stats[1:]
    px = torch.cat([(x - torch.tensor(m).view(1, 3, 1, 1)) / torch.tensor(s).view(1, 3, 1, 1) for m, s in stats], 1)
    ids = np.concatenate([np.ones((batch, 1), np.int64), rng.integers(3, 31744, (batch, prompt_len - 1))], 1)
    return torch.from_numpy(ids), px.to(torch.bfloat16).contiguous()


def algorithmic_flops_per_action(cfg, T: int) -> float:
    """2*M*N*K of every linear (discarded last ViT blocks skipped), causal attention counted at half,
    lm_head on one row; + the 6 cached decode steps."""
    f = 0.0
    np_ = cfg.n_patches
    for t in cfg.towers:
        N = np_ + t.n_prefix
        f += 2.0 * np_ * t.dim * 3 * cfg.patch ** 2
        per = 2.0 * N * (3 * t.dim * t.dim + t.dim * t.dim + 2 * t.dim * t.mlp) + 4.0 * N * N * t.dim
        f += per * (t.depth - 1)
    vd, D = cfg.vision_dim, cfg.text_config.hidden_size
    if cfg.use_fused_vision_backbone:
        f += 2.0 * np_ * (vd * 4 * vd + 4 * vd * D + D * D)
    else:
        f += 2.0 * np_ * (vd * D + D * D)
    tc = cfg.text_config
    lin = 4 * D * D + 3 * D * tc.intermediate_size
    f += tc.num_hidden_layers * (2.0 * T * lin + 2.0 * T * T * D) + 2.0 * D * tc.vocab_size
    f += 6 * (tc.num_hidden_layers * 2.0 * lin + 2.0 * D * tc.vocab_size)
    return f


# ----------------------------------------------------------------------------------------------- reference arm / CPU baseline
def cpu_reference_run(cfg_name: str, prompt_len: int, steps: int, warmup: int, budget_s: float):
    """The reference's own algorithm for the path on the host CPU: oracle port (oracle/openvla_oracle.py) of
    get_vla_action's TWO passes (capture forward + predict_action generate), fp32, bs=1, all host threads."""
    from oracle import openvla_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    d = {"openvla-7b": O.OPENVLA_7B, "siglip-7b": O.SIGLIP_7B, "tiny": O.tiny_dims()}[cfg_name]
    t0 = time.time()
    shapes = O.weight_shapes(d)
    need_gb = sum(int(np.prod(s)) for s in shapes.values()) * 4 / 1e9
    try:
        import psutil
        avail_gb = psutil.virtual_memory().available / 1e9
    except Exception:
        avail_gb = 0.0
    distinct = avail_gb > need_gb * 1.4 + 8.0          # else share one buffer per shape (same FLOPs, less RAM)
    W = {}
    cache = {}
    for name, shape in shapes.items():                 # values do not affect CPU time
        key = (shape, len(shape) == 1 and not name.endswith("bias"))
        if key not in cache:
            w = torch.empty(shape, dtype=torch.float32)
            if key[1]:
                w.fill_(1.0)
            else:
                w.uniform_(-0.035, 0.035)
            cache[key] = w
            W[name] = w
        else:
            W[name] = cache[key].clone() if distinct else cache[key]
    init_s = time.time() - t0
    ids, px = O.make_inputs(d, 1, prompt_len=prompt_len, seed=1)
    stats = O.default_stats()
    layers = list(range(d.llm_layers + 1))
    times = []
    with torch.no_grad():
        for _ in range(min(warmup, 1)):                # one untimed pass at most: a pass takes tens of seconds
            O.get_vla_action(W, d, ids, px, stats, layers, "mean", dtype=torch.float32)
        t_begin = time.time()
        for _ in range(steps):
            if times and (time.time() - t_begin) + times[-1] > budget_s:
                break
            t1 = time.time()
            O.get_vla_action(W, d, ids, px, stats, layers, "mean", dtype=torch.float32)
            times.append(time.time() - t1)
    ms = 1e3 * sum(times) / len(times)
    return {"ms_per_action": ms, "actions_per_s": 1e3 / ms, "steps_timed": len(times), "init_s": init_s,
            "cores": os.cpu_count() or 1, "distinct_weights": distinct}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.config, args.prompt_len, max(1, args.steps), max(1, args.warmup), args.cpu_budget_s)
    sample = ("1 observation per step (bs=1), full get_vla_action two-pass order (capture forward + 7-token greedy "
              "generate), fp32, all host threads; "
              + ("distinct weight buffers" if r["distinct_weights"] else "equal-shaped weights share one buffer (host RAM)"))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["actions_per_s"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps_timed"], "warmup": args.warmup, "ms_per_step": r["ms_per_action"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.config, args.batch, args.prompt_len),
                   "note": "reference arm: each step is a bs=1 sample of this workload on the host CPU"},
        "cpu_baseline": {"value": r["actions_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["actions_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- probe-training leg
def probe_cpu_baseline(kind: str, batch: int, D: int, K: int, budget_s: float = 15.0):
    """One layer's probe step as the reference trains it (train_object_probes.py:177-189: nn.Linear, masked
    BCEWithLogitsLoss with pos_weight, AdamW) through the oracle port's own loss (oracle/probe_oracle.py), fp32, all host
    threads, on a bounded sample: a few steps of the same batch of 4096 x 4096-d rows."""
    import torch.nn as nn

    from oracle import probe_oracle as PO

    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(7)
    X = torch.randn(batch, D, generator=g)
    Y = (torch.rand(batch, K, generator=g) < 0.5).to(torch.int8)
    Y[torch.rand(batch, K, generator=g) < 0.5] = -1
    model = PO.DualHeadProbe(D, K) if kind == "dual" else nn.Linear(D, K)
    pw = torch.tensor(1.7) if kind == "dual" else torch.ones(K) * 2.0
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
    ts, t_begin = [], time.time()
    for s in range(8):
        t0 = time.time()
        loss, _ = PO.loss_fn(kind, model, X, Y, pw)
        opt.zero_grad()
        loss.backward()
        opt.step()
        if s > 0:
            ts.append(time.time() - t0)
        if time.time() - t_begin > budget_s and len(ts) >= 2:
            break
    sec = sum(ts) / len(ts)
    return {"value": 1.0 / sec, "unit": "layer-steps/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"{len(ts)} timed AdamW steps of one layer's {kind} probe on one batch of {batch} x {D} fp32 rows, "
                      f"{K} labels ({sec * 1e3:.0f} ms per step), oracle/probe_oracle.py loss (the reference's own torch calls)"}


def probe_training_leg(world: int, rank: int, local: int, lib, peaks, kind: str = "object", G: int = 33, batch: int = 4096,
                       D: int = 4096, K: int = 439, steps: int = 12, warmup: int = 3, chunks: int = 0, comm_sms: int = 16,
                       cpu_baseline: bool = True):
    """BASELINE.json configs[3]: linear (object) probes of all 33 captured layers trained concurrently on synthetic
    4096-d features, multilabel BCE, AdamW; at N > 1 every rank takes its own batch of 4096 rows per layer-step (weak
    scaling, global batch 4096 x N) and the flat [dW | db | stats] gradients go through NCCL all-reduce, chunked and
    overlapped with the next chunk's compute (probes.MultiLayerProbeTrainer).  Device-timed with CUDA events, max over
    ranks.  Returns the `probe_training` block of the JSON line (rank 0) or None."""
    import torch.distributed as dist

    from openvla_probe_b200.probes import MultiLayerProbeTrainer

    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    n_ep = 2                                                   # steps per epoch held resident
    N = batch * n_ep                                           # this rank's own rows
    X = torch.randn(G, N, D, generator=g, device="cuda", dtype=torch.float32)
    Y = (torch.rand(N, 481, generator=g, device="cuda") < 0.5).to(torch.int8)
    Y[torch.rand(N, 481, generator=g, device="cuda") < 0.5] = -1
    keep = torch.arange(K)
    torch.manual_seed(0)
    pw = torch.tensor(1.7) if kind == "dual" else torch.ones(K) * 2.0
    tr = MultiLayerProbeTrainer(kind, G, D, K, pw, batch=batch, device=local, chunks=chunks, shard="local", comm_sms=comm_sms)
    perm = torch.randperm(N, generator=torch.Generator().manual_seed(1))
    tr.load_epoch(X, Y, keep, perm, drop_last=True)
    del X
    assert len(tr.steps) == n_ep
    for s in range(max(3, warmup)):
        tr.train_step(s % n_ep)
    tr.finish()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    lib.ovla_reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        tr.train_step(s % n_ep)
    tr.finish()
    e1.record()
    torch.cuda.synchronize()
    launches = int(lib.ovla_launch_count())
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # all-reduce alone (same buffers, no compute beside it), for the record
    ar_ms = None
    if world > 1:
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.all_reduce(tr.Gbuf)
        torch.cuda.synchronize()
        a0.record()
        for _ in range(3):
            dist.all_reduce(tr.Gbuf)
        a1.record()
        torch.cuda.synchronize()
        ar_ms = a0.elapsed_time(a1) / 3
    losses = tr.step_losses()
    ms_step = float(ms.item()) / steps
    rows = tr.rows
    # algorithmic HBM bytes of one layer-step (fp32): X_b read by the forward and again (transposed copy) by dW;
    # W read; Z and dZ^T written and read once; labels; dW written; AdamW reads p, g, m, v and writes p, m, v
    bytes_ls = 4.0 * (2 * batch * D + rows * D + 4 * batch * rows + rows * D + 7 * rows * D) + batch * tr.Kpad
    flops_ls = 2 * 2.0 * batch * D * rows
    us_ls = ms_step * 1e3 / G
    out = {
        "metric": "probe layer-steps/s (one AdamW step of one layer's probe on a batch of 4096 rows per GPU)",
        "kind": kind, "value": world * G * 1e3 / ms_step, "unit": "layer-steps/s", "n_gpus": world, "scaling": "weak",
        "layers_concurrent": G, "batch_per_gpu": batch, "global_batch": batch * world, "D": D, "K": K, "steps": steps,
        "ms_per_step_all_layers": ms_step, "us_per_layer_step": us_ls, "samples_per_s": world * batch * G * 1e3 / ms_step,
        "dtype": "tf32 GEMM (fp32 accumulate) / fp32 elsewhere", "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": bytes_ls / (us_ls * 1e-6) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": bytes_ls / (us_ls * 1e-6) / 1e9 / peaks["hbm_gbs"], "bytes_per_layer_step": bytes_ls,
                     "tf32_tflops": flops_ls / (us_ls * 1e-6) / 1e12,
                     "note": "whole layer-step (4 grouped launches), not a single kernel; the two TF32 GEMMs carry "
                             f"{flops_ls / 1e9:.1f} GFLOP per layer-step"},
        "allreduce": {"bytes_per_step": int(tr.Gbuf.numel() * 4) if world > 1 else 0, "chunks": len(tr.chunks),
                      "alone_ms": ar_ms, "overlapped": world > 1 and len(tr.chunks) > 1,
                      "sm_limit_of_gemms": tr.sm_limit, "NCCL_MAX_CTAS": os.environ.get("NCCL_MAX_CTAS")},
        "loss_layer0": losses[0],
    }
    del tr
    torch.cuda.empty_cache()
    if world == 1 and cpu_baseline:
        try:
            out["cpu_baseline"] = probe_cpu_baseline(kind, batch, D, K)
        except Exception as ex:  # noqa: BLE001
            out["cpu_baseline"] = {"value": None, "kind": "port", "sample": f"failed: {type(ex).__name__}: {ex}"}
    return out if rank == 0 else None


# ----------------------------------------------------------------------------------------------- configs[4] leg
def siglip_leg(world: int, rank: int, local: int, steps: int, warmup: int, prompt_len: int, batch: int = 64):
    """BASELINE.json configs[4]: prism-siglip-224px+7b (single SigLIP backbone, 2-layer gelu-mlp projector), bs = 64 per
    GPU, predict_action + 33-layer capture; same timing rules as the headline (device events, max over ranks; e2e through
    the public API with host buffers).  Returns the `siglip_bs64` block (rank 0) or None."""
    import dataclasses

    import torch.distributed as dist

    from openvla_probe_b200 import config as cfgmod, weights
    from openvla_probe_b200.modeling_prismatic import OpenVLAForActionPrediction

    stats = {"synthetic": {"action": {"q01": np.linspace(-0.9, -0.3, 7).tolist(), "q99": np.linspace(0.4, 1.0, 7).tolist(),
                                      "mask": [True] * 6 + [False]}}}
    cfg = dataclasses.replace(cfgmod.siglip_7b(), norm_stats=stats)
    model = OpenVLAForActionPrediction(cfg, max_batch=batch, max_prompt_len=prompt_len + 1, device=local)
    weights.bind_random(model, seed=0)
    ids, px = synthetic_inputs(cfg, batch, prompt_len, seed=11 + rank)
    ids29 = torch.cat([ids, torch.full((batch, 1), 29871, dtype=torch.int64)], 1)
    ids_dev, px_dev, px_pin = ids29.cuda(), px.cuda(), px.pin_memory()
    pool_len = cfg.n_patches + prompt_len

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        t = torch.tensor([e0.elapsed_time(e1), wall], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]) / n, float(t[1]) / n

    for _ in range(max(3, warmup)):
        model.engine.run(ids_dev, px_dev, pool_len, 0, 7)
    ms_dev, _ = timed(lambda: model.engine.run(ids_dev, px_dev, pool_len, 0, 7), steps)
    tc = cfg.text_config
    pool_pin = torch.empty(tc.num_hidden_layers + 1, batch, tc.hidden_size, dtype=torch.float32).pin_memory()
    api = lambda: model.predict_action_and_capture(ids, unnorm_key="synthetic", layer_indices=list(range(N_LAYERS_CAPTURED)),  # noqa: E731
                                                   pixel_values=px_pin, pooled_out=pool_pin)
    api(); api()
    _, ms_e2e = timed(api, steps)
    T = cfg.n_patches + prompt_len + 1
    step_tf = algorithmic_flops_per_action(cfg, T) * batch / (ms_dev * 1e-3) / 1e12
    out = {"workload": workload_name("siglip-7b", batch, prompt_len), "value": world * batch / (ms_dev / 1e3), "unit": UNIT,
           "ms_per_step": ms_dev, "steps": steps, "n_gpus": world, "step_tflops_per_gpu": step_tf,
           "e2e": {"value": world * batch / (ms_e2e / 1e3), "unit": UNIT,
                   "h2d_bytes_per_step": batch * (prompt_len + 1) * 8 + px.numel() * 2,
                   "d2h_bytes_per_step": (tc.num_hidden_layers + 1) * batch * tc.hidden_size * 4 + batch * 7 * 8}}
    model.engine.close()
    del model
    torch.cuda.empty_cache()
    return out if rank == 0 else None


# ----------------------------------------------------------------------------------------------- our arm
def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return

    import ctypes as C
    import dataclasses

    import torch.distributed as dist

    from openvla_probe_b200 import _lib, config as cfgmod, weights
    from openvla_probe_b200.modeling_prismatic import OpenVLAForActionPrediction

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # the probe-training all-reduce runs BESIDE persistent GEMMs that leave it `--probe-comm-sms` SMs: keep NCCL's
        # grid within that reservation (the data path of the capture benchmark itself has no collective)
        os.environ.setdefault("NCCL_MAX_CTAS", str(max(1, args.probe_comm_sms)))
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()

    if args.probe_only:
        blk = probe_training_leg(world, rank, local, lib, load_peaks(), kind=args.probe_kind, G=args.probe_layers,
                                 steps=args.steps, warmup=args.warmup, chunks=args.probe_chunks,
                                 comm_sms=args.probe_comm_sms)
        if rank == 0:
            print(json.dumps(blk), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    base = {"openvla-7b": cfgmod.openvla_7b, "siglip-7b": cfgmod.siglip_7b, "tiny": cfgmod.tiny}[args.config]()
    stats = {"synthetic": {"action": {"q01": np.linspace(-0.9, -0.3, 7).tolist(), "q99": np.linspace(0.4, 1.0, 7).tolist(),
                                      "mask": [True] * 6 + [False]}}}
    cfg = dataclasses.replace(base, norm_stats=stats)
    B, P0 = args.batch, args.prompt_len
    T = cfg.n_patches + P0 + 1
    model = OpenVLAForActionPrediction(cfg, max_batch=B, max_prompt_len=P0 + 1, device=local)
    weights.bind_random(model, seed=0)
    ids, px = synthetic_inputs(cfg, B, P0, seed=1 + rank)
    ids29 = torch.cat([ids, torch.full((B, 1), 29871, dtype=torch.int64)], 1)
    ids_dev, px_dev = ids29.cuda(), px.cuda()
    px_pin = px.pin_memory()
    pool_len = cfg.n_patches + P0
    n_act = 7

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return model.engine.run(ids_dev, px_dev, pool_len, 0, n_act)

    n_warm = args.warmup if args.lite else max(3, args.warmup)
    for _ in range(n_warm):
        step_device()
    # ---- device-resident timing (value): per-kernel event profiler OFF (it costs two cudaEventRecord per launch)
    lib.ovla_profile_enable(0)
    lib.ovla_reset_launch_count()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    launches = int(lib.ovla_launch_count())
    t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * B / (ms_step / 1e3)

    # ---- a SEPARATE profiled pass of the same steps: per-kernel CUDA events on the launch stream -> kernel_breakdown
    # and the roofline of the dominant kernel family (its own step time is reported as profiled_ms_per_step)
    cat_n = (C.c_longlong * 7)(); cat_ms = (C.c_double * 7)(); cat_fl = (C.c_double * 7)(); cat_by = (C.c_double * 7)()
    prof_steps = args.steps if args.lite else max(1, min(args.steps, 3))
    lib.ovla_profile_enable(1)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(prof_steps):
        step_device()
    p1.record()
    barrier()
    lib.ovla_profile_enable(0)
    _lib.check(lib.ovla_profile_collect(cat_n, cat_ms, cat_fl, cat_by))
    prof_ms_step = p0.elapsed_time(p1) / prof_steps

    # ---- end-to-end through the public API with HOST buffers (H2D + D2H inside the timed region)
    if args.lite:
        args.no_bs1 = args.no_cpu_baseline = True
    # The pooled states land in a caller-owned pinned buffer (`pooled_out=`, the C ABI's host-buffer contract:
    # ovla_run_host writes into the pointer it is given); `e2e_fresh` below is the same call returning freshly allocated
    # arrays (an extra 138 MB host allocation + copy per step).
    tc = cfg.text_config
    pool_pin = torch.empty(tc.num_hidden_layers + 1, B, tc.hidden_size, dtype=torch.float32).pin_memory()

    def e2e_loop(**kw):
        for _ in range(2):
            model.predict_action_and_capture(ids, unnorm_key="synthetic", layer_indices=list(range(N_LAYERS_CAPTURED)),
                                             pixel_values=px_pin, **kw)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            embeds, actions = model.predict_action_and_capture(
                ids, unnorm_key="synthetic", layer_indices=list(range(N_LAYERS_CAPTURED)), pixel_values=px_pin, **kw)
        torch.cuda.synchronize()
        t = torch.tensor([(time.perf_counter() - t0) * 1e3], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return world * B / (float(t.item()) / args.steps / 1e3)

    e2e_value = None if args.lite else e2e_loop(pooled_out=pool_pin)
    e2e_fresh = None if args.lite else e2e_loop()
    L, D = cfg.text_config.num_hidden_layers, cfg.text_config.hidden_size
    h2d = B * (P0 + 1) * 8 + px.numel() * 2
    d2h = (L + 1) * B * D * 4 + B * n_act * 8

    # ---- bs=1 latency (p50), device-resident and end-to-end
    bs1 = None
    if not args.no_bs1:
        i1, p1 = ids_dev[:1].contiguous(), px_dev[:1].contiguous()
        for _ in range(3):
            model.engine.run(i1, p1, pool_len, 0, n_act)
        lat = []
        for _ in range(15):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); model.engine.run(i1, p1, pool_len, 0, n_act); b.record(); torch.cuda.synchronize()
            lat.append(a.elapsed_time(b))
        lat_e2e = []
        px1 = px[:1].contiguous().pin_memory()
        for _ in range(15):
            t0 = time.perf_counter()
            model.predict_action_and_capture(ids[:1], unnorm_key="synthetic", layer_indices=list(range(N_LAYERS_CAPTURED)),
                                             pixel_values=px1)
            lat_e2e.append((time.perf_counter() - t0) * 1e3)
        bs1 = {"p50_ms": statistics.median(lat), "e2e_p50_ms": statistics.median(lat_e2e)}

    peaks = load_peaks()
    # ---- configs[4]: the single-backbone SigLIP model at bs = 64, on the same GPUs, once the headline model is gone
    siglip_block = None
    if not args.lite and not args.no_siglip and args.config == "openvla-7b":
        model.engine.close()
        del model
        model = None
        torch.cuda.empty_cache()
        try:
            siglip_block = siglip_leg(world, rank, local, args.steps, args.warmup, P0)
        except Exception as ex:  # noqa: BLE001 -- recorded, never takes the headline number down
            siglip_block = {"error": f"{type(ex).__name__}: {ex}"}
    # ---- probe training (configs[3]) on the same GPUs: all ranks take part (NCCL gradient all-reduce at N > 1)
    probe_block = None
    if not args.lite and not args.no_probe:
        if model is not None:
            model.engine.close()
        del model
        torch.cuda.empty_cache()
        try:
            probe_block = probe_training_leg(world, rank, local, lib, peaks, comm_sms=args.probe_comm_sms,
                                             cpu_baseline=not args.no_cpu_baseline)
        except Exception as ex:  # noqa: BLE001 -- recorded, never takes the headline number down
            probe_block = {"error": f"{type(ex).__name__}: {ex}"}
        model = None

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    traffic, traffic_note = ncu_traffic()
    names = ["gemm_tcgen05", "gemv", "flash_attn", "decode_attn", "norm", "pool", "other"]
    cats = {n: {"launches": int(cat_n[i]), "ms_per_step": cat_ms[i] / prof_steps,
                "tflops": (cat_fl[i] / (cat_ms[i] * 1e-3) / 1e12) if cat_ms[i] > 0 and cat_fl[i] > 0 else None,
                "gbs": (cat_by[i] / (cat_ms[i] * 1e-3) / 1e9) if cat_ms[i] > 0 and cat_by[i] > 0 else None}
            for i, n in enumerate(names)}
    g = cats["gemm_tcgen05"]
    gemm_tf = g["tflops"] or 0.0
    roofline = {
        "kernel": "gemm_tcgen05_kernel (all ViT / projector / Llama linears of the step)",
        "bound": "tensor", "achieved": gemm_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
        "frac": gemm_tf / peaks["tf_sustained"], "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
        "share_of_step": g["ms_per_step"] / prof_ms_step, "launches_per_step": g["launches"] / prof_steps,
        "timing": f"CUDA events around every launch on its launch stream, in a separate pass of {prof_steps} step(s) right after "
                  f"the timed region (value is timed with this profiler off); profiled step {prof_ms_step:.1f} ms",
        "traffic": traffic, "traffic_note": traffic_note,
        "decode_attn_hbm": {"achieved_gbs": cats["decode_attn"]["gbs"], "peak_gbs": peaks["hbm_gbs"],
                            "frac": (cats["decode_attn"]["gbs"] or 0.0) / peaks["hbm_gbs"]},
    }
    step_tf = algorithmic_flops_per_action(cfg, T) * B / (ms_step * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": n_warm,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": workload_name(args.config, B, P0),
                   "global_batch": world * B, "weights": "random-init N(0,0.02) bf16, replicated per GPU",
                   "l2": "per-step working set (>40 GB of activations + 15 GB weights) exceeds the 126 MB L2; no flush needed",
                   "parallelism": f"dp{world} (observations sharded, no data-path collective)"},
        "step_tflops_per_gpu": step_tf, "step_frac_of_sustained_peak": step_tf / peaks["tf_sustained"],
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "OpenVLAForActionPrediction.predict_action_and_capture(host ids, pinned host frames, pooled_out=pinned)",
                "fresh_result_value": e2e_fresh},
        "gpu_launches": launches, "roofline": roofline, "kernel_breakdown": cats,
    }
    if bs1:
        line["bs1_latency"] = bs1
    if siglip_block:
        line["siglip_bs64"] = siglip_block
    if probe_block:
        line["probe_training"] = probe_block
    if world == 1 and not args.no_cpu_baseline:
        if model is not None:
            model.engine.close()
            del model
        torch.cuda.empty_cache()
        try:
            r = cpu_reference_run(args.config, P0, 1, 0, 90.0)
            line["cpu_baseline"] = {
                "value": r["actions_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                "sample": "1 observation (bs=1), full two-pass get_vla_action order, fp32, all host threads, "
                          f"{r['steps_timed']} timed pass(es) of {r['ms_per_action'] / 1e3:.1f} s",
            }
        except Exception as ex:  # noqa: BLE001 -- the baseline must never take the GPU number down
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {type(ex).__name__}: {ex}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
