"""GEMV (M<=8 decode linear) bandwidth microbenchmark over tuning knobs; run under gpurun.
Each (threads, chunks, blocks-per-SM) variant runs in a child process (knobs are read once from the environment)."""
import itertools, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SHAPES = [("qkv", 12288, 4096, 0), ("o", 4096, 4096, 0), ("gate_up", 22016, 4096, 1), ("down", 4096, 11008, 0), ("lm_head", 32064, 4096, 2)]

def child(M):
    import ctypes as C, torch
    from openvla_probe_b200 import _lib
    lib = _lib.load()
    out = {}
    for name, N, K, mode in SHAPES:
        # 8 distinct weight copies so that consecutive launches never hit L2 (126 MB)
        Ws = [(torch.randn(N, K, device="cuda") * 0.02).bfloat16() for _ in range(8 if N * K * 2 < 150e6 else 4)]
        x = torch.randn(M, K, device="cuda").bfloat16()
        n_out = N // 2 if mode == 1 else N
        o = torch.empty(M, n_out, device="cuda", dtype=torch.float32 if mode == 2 else torch.bfloat16)
        epi = _lib.GemmEpilogue(); epi.round_bf16 = 1
        def run(W):
            _lib.check(lib.ovla_gemv(C.c_void_p(x.data_ptr()), C.c_longlong(K), C.c_void_p(W.data_ptr()), C.c_longlong(K), M, N, K, mode,
                                     C.c_void_p(o.data_ptr()), C.c_longlong(n_out), C.byref(epi), None))
        for W in Ws: run(W)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 6
        e0.record()
        for _ in range(reps):
            for W in Ws: run(W)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * len(Ws))
        out[name] = {"us": round(us, 2), "gbs": round(N * K * 2 / us / 1e3, 1)}
    print("RESULT " + json.dumps(out))

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(int(sys.argv[2])); sys.exit(0)
    os.makedirs("gpurun_out", exist_ok=True)
    f = open("gpurun_out/gemv_microbench.jsonl", "w")
    for M in (1,):
        for thr, ch, bps in itertools.product((128, 256), (2, 4, 8), (8, 16)):
            env = dict(os.environ, OVLA_GEMV_THREADS=str(thr), OVLA_GEMV_CH=str(ch), OVLA_GEMV_BPS=str(bps))
            r = subprocess.run([sys.executable, __file__, "child", str(M)], env=env, capture_output=True, text=True, timeout=120)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            rec = {"M": M, "threads": thr, "ch": ch, "bps": bps, "res": json.loads(line[-1][7:]) if line else r.stderr[-300:]}
            print(json.dumps(rec), flush=True); f.write(json.dumps(rec) + "\n"); f.flush()
