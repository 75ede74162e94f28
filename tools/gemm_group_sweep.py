"""Per-shape sustained TFLOP/s vs rasterisation group size (run under gpurun)."""
import ctypes as C, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SHAPES = [("qkv", 73728, 12288, 4096, 0), ("o", 73728, 4096, 4096, 0), ("gate_up", 73728, 22016, 4096, 1), ("down", 73728, 4096, 11008, 0),
          ("vit_fc1", 66816, 4096, 1024, 0), ("vit_fc2", 66816, 1024, 4096, 0), ("sig_fc1", 65536, 4304, 1152, 0), ("proj1", 65536, 8704, 2176, 0)]
def child():
    import torch
    from openvla_probe_b200 import _lib
    lib = _lib.load(); out = {}
    for name, M, N, K, mode in SHAPES:
        A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16(); W = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
        n_out = N // 2 if mode == 1 else N
        o = torch.empty(M, n_out, device="cuda", dtype=torch.bfloat16); epi = _lib.GemmEpilogue()
        def run():
            _lib.check(lib.ovla_gemm(C.c_void_p(A.data_ptr()), C.c_longlong(K), C.c_void_p(W.data_ptr()), C.c_longlong(K), M, N, K, mode, 0,
                                     C.c_void_p(o.data_ptr()), C.c_longlong(n_out), C.byref(epi), 0, 0, None))
        fl = 2.0 * M * N * K
        reps = max(20, int(0.8e15 / fl))
        for _ in range(reps // 4): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): run()
        e1.record(); torch.cuda.synchronize()
        out[name] = round(fl * reps / e0.elapsed_time(e1) / 1e9, 1)
        del A, W, o
    print("RESULT " + json.dumps(out))
if __name__ == "__main__":
    if len(sys.argv) > 1: child(); sys.exit(0)
    for g in (8, 12, 16, 24, 16, 8):
        r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, OVLA_GEMM_GROUP=str(g)), capture_output=True, text=True, timeout=200)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        print(json.dumps({"group_m": g, "tflops": json.loads(line[-1][7:]) if line else r.stderr[-300:]}), flush=True)
