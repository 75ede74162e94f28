"""Bring-up check of the tcgen05 GEMM on a B200 (run under gpurun). Each case runs in its own process with a
timeout so that a hung kernel cannot take the whole call down. Writes gpurun_out/gemm_check.jsonl.

    python tools/gpu_gemm_check.py            # all cases
    python tools/gpu_gemm_check.py --case i   # one case (child)
"""
import argparse
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# (name, M, N, K, mode, kind, bn, cg, flags)
CASES = [
    ("cg1_bn128_small", 128, 128, 64, 0, 0, 128, 1, {}),
    ("cg1_bn128_k256", 256, 256, 256, 0, 0, 128, 1, {}),
    ("cg1_bn256", 512, 512, 512, 0, 0, 256, 1, {}),
    ("cg1_bn64", 200, 192, 320, 0, 0, 64, 1, {"bias": 1}),
    ("cg1_ragged", 333, 4304, 1152, 0, 0, 128, 1, {"bias": 1, "gelu": 1}),
    ("cg1_resid_ls", 1000, 1024, 4096, 0, 0, 256, 1, {"bias": 1, "scale": 1, "resid": 1}),
    ("cg1_swiglu", 300, 2 * 1024, 512, 1, 0, 128, 1, {}),
    ("cg1_f32", 77, 32064, 512, 2, 0, 128, 1, {"round": 1}),
    ("cg2_bn256_small", 256, 256, 64, 0, 0, 256, 2, {}),
    ("cg2_bn256", 1024, 1024, 1024, 0, 0, 256, 2, {"bias": 1}),
    ("cg2_bn128", 700, 1152, 1152, 0, 0, 128, 2, {"bias": 1, "resid": 1}),
    ("cg2_swiglu", 520, 2 * 2048, 1024, 1, 0, 256, 2, {}),
    ("tf32_cg1", 300, 440, 1024, 2, 1, 128, 1, {"bias_f32": 1}),
    ("tf32_cg2", 4096, 440, 4096, 2, 1, 256, 2, {"bias_f32": 1}),
    ("big_cg1", 8192, 4096, 4096, 0, 0, 256, 1, {"time": 1}),
    ("big_cg2", 8192, 4096, 4096, 0, 0, 256, 2, {"time": 1}),
    ("big_cg2_bn128", 8192, 4096, 4096, 0, 0, 128, 2, {"time": 1}),
    ("huge_cg2", 73728, 4096, 4096, 0, 0, 256, 2, {"time": 1, "nocheck": 1}),
    ("huge_cg1", 73728, 4096, 4096, 0, 0, 256, 1, {"time": 1, "nocheck": 1}),
    ("auto_heur", 4096, 11008 * 2, 4096, 1, 0, 0, 0, {"time": 1}),
]


def run_case(i):
    import ctypes as C

    import torch

    from openvla_probe_b200 import _lib

    name, M, N, K, mode, kind, bn, cg, fl = CASES[i]
    lib = _lib.load()
    torch.manual_seed(i)
    dev = "cuda"
    dt = torch.bfloat16 if kind == 0 else torch.float32
    A = (torch.randn(M, K, device=dev) * 0.5).to(dt)
    W = (torch.randn(N, K, device=dev) * 0.05).to(dt)
    n_out = N // 2 if mode == 1 else N
    out = torch.full((M, n_out), 7.0, device=dev, dtype=torch.bfloat16 if mode != 2 else torch.float32)
    epi = _lib.GemmEpilogue()
    bias = scale = resid = bias32 = None
    if fl.get("bias"):
        bias = (torch.randn(N, device=dev) * 0.1).bfloat16()
        epi.bias_bf16 = bias.data_ptr()
    if fl.get("scale"):
        scale = (torch.randn(N, device=dev) * 0.5).bfloat16()
        epi.scale_bf16 = scale.data_ptr()
    if fl.get("resid"):
        resid = torch.randn(M, N, device=dev).bfloat16()
        epi.resid_bf16 = resid.data_ptr()
        epi.ld_resid = N
    if fl.get("bias_f32"):
        bias32 = torch.randn(N, device=dev)
        epi.bias_f32 = bias32.data_ptr()
    epi.gelu = int(fl.get("gelu", 0))
    epi.round_bf16 = int(fl.get("round", 0))

    def call():
        rc = lib.ovla_gemm(C.c_void_p(A.data_ptr()), C.c_longlong(K), C.c_void_p(W.data_ptr()), C.c_longlong(K),
                           M, N, K, mode, kind, C.c_void_p(out.data_ptr()), C.c_longlong(n_out), C.byref(epi),
                           bn, cg, None)
        _lib.check(rc)

    call()
    torch.cuda.synchronize()
    res = {"case": name, "M": M, "N": N, "K": K, "mode": mode, "kind": kind, "bn": bn, "cg": cg}
    if not fl.get("nocheck"):
        if kind == 1:
            torch.backends.cuda.matmul.allow_tf32 = False
        ref = A.float() @ W.float().t()
        if mode == 1:
            r3 = ref.view(M, N // 64, 2, 32)
            g, u = r3[:, :, 0, :].reshape(M, -1), r3[:, :, 1, :].reshape(M, -1)
            g, u = g.bfloat16().float(), u.bfloat16().float()
            ref = (torch.nn.functional.silu(g).bfloat16().float() * u)
        else:
            if bias is not None:
                ref = ref + bias.float()
            if bias32 is not None:
                ref = ref + bias32
            if mode == 0 or fl.get("round"):
                ref = ref.bfloat16().float()
            if fl.get("gelu"):
                ref = torch.nn.functional.gelu(ref).bfloat16().float()
            if scale is not None:
                ref = (ref * scale.float()).bfloat16().float()
            if resid is not None:
                ref = ref + resid.float()
        if mode != 2:
            ref = ref.bfloat16().float()
        got = out.float()
        err = (got - ref).abs()
        denom = ref.abs().max().item() + 1e-9
        res.update(max_abs=err.max().item(), rel_to_max=err.max().item() / denom,
                   mismatch_frac=(err > 0.02 * denom).float().mean().item(),
                   untouched=(got == 7.0).float().mean().item())
        tol = 2e-2 if kind == 0 else 5e-3
        res["ok"] = bool(res["rel_to_max"] < tol)
    if fl.get("time"):
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res.update(ms=ms, tflops=2.0 * M * N * K / ms / 1e9)
    print("RESULT " + json.dumps(res), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", type=int, default=-1)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    if args.case >= 0:
        run_case(args.case)
        return
    os.makedirs("gpurun_out", exist_ok=True)
    out = open("gpurun_out/gemm_check.jsonl", "w")
    for i, c in enumerate(CASES):
        if args.only and args.only not in c[0]:
            continue
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--case", str(i)], capture_output=True, text=True, timeout=120)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if line:
                rec = json.loads(line[-1][7:])
            else:
                rec = {"case": c[0], "ok": False, "rc": r.returncode, "stderr": r.stderr[-800:]}
        except subprocess.TimeoutExpired:
            rec = {"case": c[0], "ok": False, "timeout": True}
        rec["wall_s"] = round(time.time() - t0, 1)
        print(json.dumps(rec), flush=True)
        out.write(json.dumps(rec) + "\n")
        out.flush()


if __name__ == "__main__":
    main()
