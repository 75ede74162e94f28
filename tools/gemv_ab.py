"""A/B of the two weight-streaming kernels (register-pipelined vs cp.async.bulk shared-memory ring) on the Llama decode
shapes, cold weights (8 distinct copies per shape), M in {1, 2, 4}; also checks that both return the same bits."""
import json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.gemv_microbench import SHAPES  # noqa: E402


def child(M):
    import ctypes as C, torch
    from openvla_probe_b200 import _lib
    lib = _lib.load()
    out = {}
    torch.manual_seed(0)
    for name, N, K, mode in SHAPES:
        Ws = [(torch.randn(N, K, device="cuda") * 0.02).bfloat16() for _ in range(8 if N * K * 2 < 150e6 else 4)]
        x = torch.randn(M, K, device="cuda").bfloat16()
        n_out = N // 2 if mode == 1 else N
        o = torch.empty(M, n_out, device="cuda", dtype=torch.float32 if mode == 2 else torch.bfloat16)
        epi = _lib.GemmEpilogue(); epi.round_bf16 = 1
        def run(W):
            _lib.check(lib.ovla_gemv(x.data_ptr(), K, W.data_ptr(), K, M, N, K, mode, o.data_ptr(), n_out, C.byref(epi), None))
        for W in Ws: run(W)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 6
        e0.record()
        for _ in range(reps):
            for W in Ws: run(W)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * len(Ws))
        run(Ws[0]); torch.cuda.synchronize()
        out[name] = {"us": round(us, 2), "gbs": round(N * K * 2 / us / 1e3, 1), "sum": float(o.double().sum()), "absmax": float(o.abs().max())}
    print("RESULT " + json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(int(sys.argv[2])); sys.exit(0)
    for M in (1, 2, 4):
        recs = {}
        for tma in ("0", "1"):
            env = dict(os.environ, OVLA_GEMV_TMA=tma)
            r = subprocess.run([sys.executable, __file__, "child", str(M)], env=env, capture_output=True, text=True, timeout=300)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            recs[tma] = json.loads(line[-1][7:]) if line else r.stderr[-600:]
        same = isinstance(recs["0"], dict) and isinstance(recs["1"], dict) and all(recs["0"][k]["sum"] == recs["1"][k]["sum"] and recs["0"][k]["absmax"] == recs["1"][k]["absmax"] for k in recs["0"])
        print(json.dumps({"M": M, "register_kernel": recs["0"], "tma_ring_kernel": recs["1"], "bit_identical_checksums": same}), flush=True)
