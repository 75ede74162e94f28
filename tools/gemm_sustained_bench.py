"""Sustained (power-capped) throughput of the big prefill GEMM shapes vs rasterisation group size; run under gpurun.
Each group size runs in a child process (the knob is read once from the environment)."""
import ctypes as C, json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SHAPES = [("qkv", 73728, 12288, 4096, 0), ("o", 73728, 4096, 4096, 0), ("gate_up", 73728, 22016, 4096, 1), ("down", 73728, 4096, 11008, 0)]

def child():
    import torch
    from openvla_probe_b200 import _lib
    lib = _lib.load()
    bufs = {}
    for name, M, N, K, mode in SHAPES:
        A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
        W = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
        n_out = N // 2 if mode == 1 else N
        bufs[name] = (A, W, torch.empty(M, n_out, device="cuda", dtype=torch.bfloat16), n_out)
    epi = _lib.GemmEpilogue()
    def layer():
        for name, M, N, K, mode in SHAPES:
            A, W, o, n_out = bufs[name]
            _lib.check(lib.ovla_gemm(C.c_void_p(A.data_ptr()), C.c_longlong(K), C.c_void_p(W.data_ptr()), C.c_longlong(K), M, N, K, mode, 0,
                                     C.c_void_p(o.data_ptr()), C.c_longlong(n_out), C.byref(epi), 0, 0, None))
    for _ in range(8): layer()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 64            # ~2 s: long enough to sit under the power cap
    e0.record()
    for _ in range(reps): layer()
    e1.record(); torch.cuda.synchronize()
    fl = sum(2.0 * M * N * K for _, M, N, K, _ in SHAPES)
    ms = e0.elapsed_time(e1) / reps
    print("RESULT " + json.dumps({"ms_per_layer": round(ms, 3), "tflops": round(fl / ms / 1e9, 1)}))

if __name__ == "__main__":
    if sys.argv[1:] == ["child"]: child(); sys.exit(0)
    if sys.argv[1:] == ["l2"]:
        for hint in ("nn", "ln", "lf", "nf", "ll", "nn", "ln"):
            r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, OVLA_GEMM_L2=hint), capture_output=True, text=True, timeout=200)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            print(json.dumps({"l2_hint_a_w": hint, "res": json.loads(line[-1][7:]) if line else r.stderr[-300:]}), flush=True)
        sys.exit(0)
    for g in (8, 4, 16, 32, 8, 16):
        r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, OVLA_GEMM_GROUP=str(g)), capture_output=True, text=True, timeout=200)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        print(json.dumps({"group_m": g, "res": json.loads(line[-1][7:]) if line else r.stderr[-300:]}), flush=True)
