"""Per-shape cost of the fused-RMSNorm epilogue parts on the four Llama prefill GEMMs at bs = 256 (M = 73728): each
GEMM timed with and without its producer (row sums of squares out) / consumer (1/rms row scale in) half, plus the
stand-alone RMSNorm kernel the fusion removes and the one sum-of-squares kernel it adds.  Run under gpurun."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvla_probe_b200 import _lib
from openvla_probe_b200.engine import rope_tables

lib = _lib.load()
P = lambda t: C.c_void_p(t.data_ptr())
M, D, I, H, T = int(os.environ.get("AB_M", 73728)), 4096, 11008, 32, 288
B = M // T
x = (torch.randn(M, D, device="cuda") * 0.5).bfloat16()
act = (torch.randn(M, I, device="cuda") * 0.5).bfloat16()
xres = (torch.randn(M, D, device="cuda")).bfloat16()
Wqkv = (torch.randn(3 * D, D, device="cuda") * 0.02).bfloat16()
Wo = (torch.randn(D, D, device="cuda") * 0.02).bfloat16()
Wgu = (torch.randn(2 * I, D, device="cuda") * 0.02).bfloat16()
Wd = (torch.randn(D, I, device="cuda") * 0.02).bfloat16()
gam = torch.ones(D, device="cuda", dtype=torch.bfloat16)
qkv = torch.empty(M, 3 * D, device="cuda", dtype=torch.bfloat16)
o_act = torch.empty(M, I, device="cuda", dtype=torch.bfloat16)
h = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
ss = torch.ones(M, D // 64, device="cuda")
Tmax = T + 8
kc = torch.zeros(B, H, Tmax, 128, device="cuda", dtype=torch.bfloat16)
vc = torch.zeros_like(kc)
cos, sin = rope_tables(128, 10000.0, Tmax)
cos, sin = cos.cuda(), sin.cuda()


def qkv_call(fused):
    if fused:
        _lib.check(lib.ovla_qkv_rope_gemm_rownorm(P(x), D, P(Wqkv), D, M, H, D, T, 0, P(cos), P(sin), P(qkv), 3 * D, P(kc), P(vc),
                                                  Tmax, P(ss), D // 64, D // 64, C.c_float(1e-5), 0, 0, None))
    else:
        _lib.check(lib.ovla_qkv_rope_gemm(P(x), D, P(Wqkv), D, M, H, D, T, 0, P(cos), P(sin), P(qkv), 3 * D, P(kc), P(vc), Tmax,
                                          0, 0, None))


def res_call(A, W, K, fused):
    epi = _lib.GemmEpilogue()
    epi.resid_bf16, epi.ld_resid = xres.data_ptr(), D
    if fused:
        epi.row_sumsq_out, epi.row_sumsq_ld = ss.data_ptr(), D // 64
    _lib.check(lib.ovla_gemm(P(A), K, P(W), K, M, D, K, 0, 0, P(xres), D, C.byref(epi), 0, 0, None))


def gu_call(fused):
    epi = _lib.GemmEpilogue()
    if fused:
        epi.row_sumsq_in, epi.row_sumsq_ld, epi.row_sumsq_parts, epi.norm_eps = ss.data_ptr(), D // 64, D // 64, 1e-5
    _lib.check(lib.ovla_gemm(P(x), D, P(Wgu), D, M, 2 * I, D, 1, 0, P(o_act), I, C.byref(epi), 0, 0, None))


CASES = {
    "qkv_rope": lambda f: qkv_call(f), "o_proj": lambda f: res_call(x, Wo, D, f), "gate_up": lambda f: gu_call(f),
    "down": lambda f: res_call(act, Wd, I, f),
}


def timeit(fn, reps=12):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def paired(fn, pairs=40):
    """plain / fused launches alternate under sustained load (the power cap moves the clock over seconds): medians of
    per-launch CUDA-event times."""
    import statistics
    ts = {0: [], 1: []}
    evs = []
    for i in range(pairs):
        for f in (0, 1):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(f); b.record()
            evs.append((f, a, b))
    torch.cuda.synchronize()
    for f, a, b in evs[8:]:
        ts[f].append(a.elapsed_time(b))
    return statistics.median(ts[0]), statistics.median(ts[1])


for _ in range(40):          # ~2 s of load first: measure at the sustained clock
    gu_call(0)
torch.cuda.synchronize()
for rnd in range(2):
    for name, fn in CASES.items():
        t0, t1 = paired(fn)
        print(json.dumps({"gemm": name, "M": M, "plain_us": round(t0 * 1e3, 1), "fused_us": round(t1 * 1e3, 1),
                          "delta_us": round((t1 - t0) * 1e3, 1)}), flush=True)
t = timeit(lambda: _lib.check(lib.ovla_rmsnorm(P(x), D, P(gam), C.c_float(1e-5), P(h), D, M, D, None)))
print(json.dumps({"kernel": "rmsnorm (removed, x2 per layer)", "us": round(t * 1e3, 1)}))
t = timeit(lambda: _lib.check(lib.ovla_row_sumsq(P(x), D, M, D, P(ss), D // 64, None)))
print(json.dumps({"kernel": "row_sumsq (added, x1 per step)", "us": round(t * 1e3, 1)}))
