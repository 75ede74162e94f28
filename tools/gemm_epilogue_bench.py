"""Per-shape throughput of the GEMMs whose epilogue matters (short K, GELU, LayerScale + residual); run under gpurun."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
#        name            M      N      K     bias gelu scale resid
SHAPES = [("dino_qkv",  66816, 3072,  1024,  1, 0, 0, 0), ("dino_proj", 66816, 1024, 1024, 1, 0, 1, 1),
          ("dino_fc1",  66816, 4096,  1024,  1, 1, 0, 0), ("dino_fc2",  66816, 1024, 4096, 1, 0, 1, 1),
          ("sig_qkv",   65536, 3456,  1152,  1, 0, 0, 0), ("sig_proj",  65536, 1152, 1152, 1, 0, 0, 1),
          ("sig_fc1",   65536, 4304,  1152,  1, 1, 0, 0), ("sig_fc2",   65536, 1152, 4304, 1, 0, 0, 1),
          ("proj_fc1",  65536, 8704,  2176,  1, 1, 0, 0), ("llama_o",   72448, 4096, 4096, 0, 0, 0, 1),
          ("llama_down", 72448, 4096, 11008, 0, 0, 0, 1), ("llama_gate_up_swiglu", 72448, 22016, 4096, 0, 0, 0, 0),
          ("dino_like_swiglu_k1024", 66816, 8192, 1024, 0, 0, 0, 0)]


def main():
    import torch
    from openvla_probe_b200 import _lib
    lib = _lib.load()
    P = lambda t: C.c_void_p(t.data_ptr())
    out = {}
    for name, M, N, K, bias, gelu, scale, resid in SHAPES:
        A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
        W = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
        b = (torch.randn(N, device="cuda") * 0.1).bfloat16()
        sc = (1 + 0.1 * torch.randn(N, device="cuda")).bfloat16()
        swiglu = "swiglu" in name
        n_out = N // 2 if swiglu else N
        X = [(torch.randn(M, n_out, device="cuda")).bfloat16() for _ in range(2)]
        epi = _lib.GemmEpilogue()
        if bias: epi.bias_bf16 = b.data_ptr()
        epi.gelu = gelu
        if scale: epi.scale_bf16 = sc.data_ptr()

        def run(x):
            if resid:
                epi.resid_bf16, epi.ld_resid = x.data_ptr(), N      # in place, like the residual stream
            _lib.check(lib.ovla_gemm(P(A), C.c_longlong(K), P(W), C.c_longlong(K), M, N, K, 1 if swiglu else 0, 0, P(x),
                                     C.c_longlong(n_out), C.byref(epi), 0, 0, None))
        for x in X: run(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            for x in X: run(x)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (2 * reps)
        out[name] = {"us": round(us, 1), "tflops": round(2.0 * M * N * K / us / 1e6, 1)}
        del A, W, X
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
