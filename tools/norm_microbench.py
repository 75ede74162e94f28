"""RMSNorm / LayerNorm bandwidth vs launch knobs (run under gpurun)."""
import ctypes as C, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
def child():
    import torch
    from openvla_probe_b200 import _lib
    lib = _lib.load(); out = {}
    for name, rows, D in (("rms_73728x4096", 73728, 4096), ("ln_66816x1024", 66816, 1024), ("ln_65536x1152", 65536, 1152)):
        xs = [torch.randn(rows, D, device="cuda").bfloat16() for _ in range(3)]
        w = torch.ones(D, device="cuda").bfloat16(); b = torch.zeros(D, device="cuda").bfloat16()
        o = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
        def run(x):
            if name.startswith("rms"):
                _lib.check(lib.ovla_rmsnorm(C.c_void_p(x.data_ptr()), C.c_longlong(D), C.c_void_p(w.data_ptr()), C.c_float(1e-6), C.c_void_p(o.data_ptr()), C.c_longlong(D), rows, D, None))
            else:
                _lib.check(lib.ovla_layernorm(C.c_void_p(x.data_ptr()), C.c_longlong(D), C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()), C.c_float(1e-6), C.c_void_p(o.data_ptr()), C.c_longlong(D), rows, D, None))
        for x in xs: run(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            for x in xs: run(x)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 30
        out[name] = {"us": round(us, 1), "gbs": round(4.0 * rows * D / us / 1e3, 1)}
    print("RESULT " + json.dumps(out))
if __name__ == "__main__":
    if len(sys.argv) > 1: child(); sys.exit(0)
    for thr, strm, split in ((128, 0, 0), (128, 0, 1), (128, 1, 0), (64, 0, 0), (128, 0, 1)):
        r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, OVLA_NORM_THREADS=str(thr), OVLA_NORM_STREAM=str(strm), OVLA_NORM_SPLIT=str(split)), capture_output=True, text=True, timeout=120)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        print(json.dumps({"threads": thr, "stream": strm, "split": split, "res": json.loads(line[-1][7:]) if line else r.stderr[-300:]}), flush=True)
