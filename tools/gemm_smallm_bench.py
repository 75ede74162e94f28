"""tcgen05 GEMM tile-config sweep for the bs=1 prefill shapes (M = 288 / 261): weights rotate over several copies so
that every launch streams them from HBM, as in the real 32-layer pass. Run under gpurun."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvla_probe_b200 import _lib
lib = _lib.load()
MM = int(sys.argv[1]) if len(sys.argv) > 1 else 0
SHAPES = [("qkv", 288, 12288, 4096, 0), ("o", 288, 4096, 4096, 0), ("gate_up", 288, 22016, 4096, 1), ("down", 288, 4096, 11008, 0),
          ("vit_qkv", 261, 3072, 1024, 0), ("vit_fc1", 261, 4096, 1024, 0), ("vit_fc2", 261, 1024, 4096, 0), ("sig_fc1", 256, 4304, 1152, 0)]
if MM:
    SHAPES = [(n, MM, N, K, mode) for n, _, N, K, mode in SHAPES[:4]] + [("lm_head", MM, 32064, 4096, 2)]
CFGS = [(0, 0), (64, 1), (128, 1), (256, 1), (256, 2)]   # (0, 0) = library heuristic (incl. split-K unless OVLA_SPLITK=0)
os.makedirs("gpurun_out", exist_ok=True)
f = open(f"gpurun_out/gemm_smallm_M{MM}.jsonl", "w")
for name, M, N, K, mode in SHAPES:
    n_copy = max(2, min(8, int(600e6 // (N * K * 2))))
    Ws = [(torch.randn(N, K, device="cuda") * 0.02).bfloat16() for _ in range(n_copy)]
    A = torch.randn(M, K, device="cuda").bfloat16()
    n_out = N // 2 if mode == 1 else N
    out = torch.empty(M, n_out, device="cuda", dtype=torch.float32 if mode == 2 else torch.bfloat16)
    row = {"shape": name, "M": M, "N": N, "K": K, "ideal_us": round(N * K * 2 / 6.53e6, 1)}
    for bn, cg in CFGS:
        epi = _lib.GemmEpilogue()
        def run(W):
            _lib.check(lib.ovla_gemm(C.c_void_p(A.data_ptr()), C.c_longlong(K), C.c_void_p(W.data_ptr()), C.c_longlong(K), M, N, K, mode, 0,
                                     C.c_void_p(out.data_ptr()), C.c_longlong(n_out), C.byref(epi), bn, cg, None))
        for W in Ws: run(W)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            for W in Ws: run(W)
        e1.record(); torch.cuda.synchronize()
        row[f"bn{bn}_cg{cg}_us"] = round(e0.elapsed_time(e1) * 1e3 / (5 * len(Ws)), 1)
    print(json.dumps(row), flush=True); f.write(json.dumps(row) + "\n")
