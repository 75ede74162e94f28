"""Tile rasterisation of the four big Llama prefill GEMMs (bs = 256: M = 73 728) vs DRAM traffic and sustained speed.

One process, knobs set through ovla_debug_gemm_raster (row-tiles per group G, column-tiles per super-group NC, L2
eviction hints of the A / W loads).  Two passes over the same config list:

    python tools/gemm_raster_sweep.py time  [out.jsonl]   # sustained TFLOP/s per shape and config (round robin, twice)
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
        --clock-control none --csv --log-file gpurun_out/raster_ncu.csv python tools/gemm_raster_sweep.py ncu
    python tools/gemm_raster_sweep.py join gpurun_out/raster_ncu.csv   # launch order -> (shape, config) table

o_proj / down_proj run with the in-place residual epilogue, gate_up with SwiGLU, as in the engine.
"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# name: (M, N, K, mode, in-place residual, gelu)
SHAPES_LLAMA = {"qkv": (73728, 12288, 4096, 0, False, False), "o": (73728, 4096, 4096, 0, True, False),
                "gate_up": (73728, 22016, 4096, 1, False, False), "down": (73728, 4096, 11008, 0, True, False)}
SHAPES_VIT = {"dino_qkv": (66816, 3072, 1024, 0, False, False), "dino_proj": (66816, 1024, 1024, 0, True, False),
              "dino_fc1": (66816, 4096, 1024, 0, False, True), "dino_fc2": (66816, 1024, 4096, 0, True, False),
              "sig_qkv": (65536, 3456, 1152, 0, False, False), "sig_fc1": (65536, 4304, 1152, 0, False, True),
              "sig_fc2": (65536, 1152, 4304, 0, True, False), "proj_fc1": (65536, 8704, 2176, 0, False, True),
              "proj_fc2": (65536, 4096, 8704, 0, False, True)}
# (G, NC, l2_a, l2_b, sync, serpentine); NC = 0: all column-tiles; hints 0 normal / 1 evict-first / 2 evict-last; sync = K blocks
# between two alignment points of the producers (0 = off, >= K/64: once per tile); -1 = the launcher's heuristic
AUTO = (-1, -1, -1, -1, -1, -1)
CONFIGS_LLAMA = {   # (G, NC, l2_a, l2_b, sync, serpentine)
    "qkv": [(16, 0, 0, 0, 0, 0), AUTO, (16, 0, 0, 0, 64, 1), (20, 0, 0, 0, 64, 1), (24, 0, 0, 0, 64, 1), (12, 0, 0, 0, 64, 1),
            (2, 16, 0, 0, 64, 0), (2, 16, 0, 0, 64, 1)],
    "o": [(8, 0, 0, 0, 0, 0), AUTO],
    "gate_up": [(16, 0, 0, 0, 0, 0), AUTO, (16, 0, 0, 0, 64, 1), (20, 0, 0, 0, 64, 1), (24, 0, 0, 0, 64, 1), (12, 0, 0, 0, 64, 1)],
    "down": [(16, 0, 0, 0, 0, 0), AUTO, (2, 8, 0, 0, 58, 1), (3, 0, 0, 0, 58, 1), (4, 0, 0, 0, 58, 1)],
}
VIT_GRID = [(8, 0, 0, 0, 0, 0), (16, 0, 0, 0, 0, 0), AUTO, (8, 0, 0, 0, 999, 0), (16, 0, 0, 0, 999, 0), (2, 0, 0, 0, 999, 0), (4, 0, 0, 0, 999, 0)]
CONFIGS_VIT = {n: VIT_GRID for n in SHAPES_VIT}
SHAPES, CONFIGS = dict(SHAPES_LLAMA), dict(CONFIGS_LLAMA)
if os.environ.get("RASTER_SET") == "vit":
    SHAPES, CONFIGS = dict(SHAPES_VIT), dict(CONFIGS_VIT)


def setup():
    import torch
    from openvla_probe_b200 import _lib
    lib = _lib.load()
    bufs = {}
    for name, (M, N, K, mode, resid, gelu) in SHAPES.items():
        A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
        W = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
        n_out = N // 2 if mode == 1 else N
        o = torch.zeros(M, n_out, device="cuda", dtype=torch.bfloat16)
        bias = torch.zeros(N, device="cuda", dtype=torch.bfloat16) if gelu else None
        bufs[name] = (A, W, o, n_out, bias)

    def run(name):
        M, N, K, mode, resid, gelu = SHAPES[name]
        A, W, o, n_out, bias = bufs[name]
        epi = _lib.GemmEpilogue()
        if resid:
            epi.resid_bf16 = o.data_ptr()
            epi.ld_resid = n_out
        if gelu:
            epi.bias_bf16 = bias.data_ptr()
            epi.gelu = 1
        _lib.check(lib.ovla_gemm(A.data_ptr(), K, W.data_ptr(), K, M, N, K, mode, 0, o.data_ptr(), n_out, C.byref(epi), 0, 0, None))

    return torch, lib, run, bufs


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "time"
    if what == "join":
        import csv
        rows = list(csv.reader(open(sys.argv[2], errors="replace")))
        hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
        h = rows[hi]
        idi, mi, vi, ki = h.index("ID"), h.index("Metric Name"), h.index("Metric Value"), h.index("Kernel Name")
        per = {}
        for r in rows[hi + 1:]:
            if len(r) <= vi or not r[idi].isdigit() or "gemm_tcgen05" not in r[ki]:
                continue
            per.setdefault(int(r[idi]), {})[r[mi]] = float(r[vi].replace(",", ""))
        order = [(n, c) for n in SHAPES for c in CONFIGS[n]]
        ids = sorted(per)
        assert len(ids) == len(order), (len(ids), len(order))
        for i, (n, c) in zip(ids, order):
            m = per[i]
            M, N, K, mode, resid, _ = SHAPES[n]
            n_out = N // 2 if mode == 1 else N
            algo = 2.0 * (M * K + N * K + M * n_out * (2 if resid else 1))
            rd, wr = m.get("dram__bytes_read.sum", 0), m.get("dram__bytes_write.sum", 0)
            unit_fix = 1.0
            print(json.dumps({"shape": n, "G": c[0], "NC": c[1], "l2_a_w": "anfl"[c[2] + 1] + "anfl"[c[3] + 1], "sync": c[4], "serp": c[5],
                              "dram_read_gb": round(rd * unit_fix / 1e9, 2), "dram_write_gb": round(wr / 1e9, 2),
                              "x_algorithmic": round((rd + wr) / algo, 2), "algorithmic_gb": round(algo / 1e9, 2),
                              "l2_hit_pct": round(m.get("lts__t_sector_hit_rate.pct", -1), 1),
                              "ncu_us": round(m.get("gpu__time_duration.sum", 0) / 1e3, 1)}))
        return
    torch, lib, run, bufs = setup()
    if what == "ncu":
        for name in SHAPES:
            for c in CONFIGS[name]:
                lib.ovla_debug_gemm_raster(*c)
                run(name)
        torch.cuda.synchronize()
        return
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 else None
    # the rasterisation must not change a single bit of the result
    for name in SHAPES:
        ref = None
        for c in CONFIGS[name]:
            lib.ovla_debug_gemm_raster(*c)
            bufs[name][2].zero_()
            run(name)
            torch.cuda.synchronize()
            if ref is None: ref = bufs[name][2].clone()
            else: assert torch.equal(ref, bufs[name][2]), (name, c)
        del ref
    print("bit-identical outputs over all configs", flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # warm the box into its power-capped state
    lib.ovla_debug_gemm_raster(*AUTO)
    for _ in range(40):
        for name in SHAPES: run(name)
    torch.cuda.synchronize()
    for rnd in range(2):
        for name, (M, N, K, mode, resid, _) in SHAPES.items():
            fl = 2.0 * M * N * K
            reps = max(8, int(0.5e-3 * 1.2e15 / fl * 1e3))   # ~0.5 s per config
            for c in CONFIGS[name]:
                lib.ovla_debug_gemm_raster(*c)
                for _ in range(3): run(name)
                e0.record()
                for _ in range(reps): run(name)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                line = json.dumps({"round": rnd, "shape": name, "G": c[0], "NC": c[1], "l2_a_w": "anfl"[c[2] + 1] + "anfl"[c[3] + 1], "sync": c[4], "serp": c[5],
                                   "us": round(ms * 1e3, 1), "tflops": round(fl / ms / 1e9, 1)})
                print(line, flush=True)
                if out: out.write(line + "\n"); out.flush()


if __name__ == "__main__":
    main()
