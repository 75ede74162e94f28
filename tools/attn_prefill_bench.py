"""Causal prefill attention (Llama-2-7B shape): mma.sync flash kernel vs the tcgen05 kernel (run under gpurun).
Both read the rotated queries from the fused qkv buffer and K/V from the cache layout [B, H, Tmax, 128]."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from openvla_probe_b200 import _lib
    lib = _lib.load()
    P = lambda t: C.c_void_p(t.data_ptr())
    H, hd, Tmax = 32, 128, 320
    D = H * hd
    for B, T in ((256, 283), (64, 283), (16, 283), (1, 283)):
        # 3 input sets (> L2 at B=256: 256*283*12288*2 B = 1.8 GB each)
        n_sets = 3 if B >= 64 else 1
        sets = []
        for _ in range(n_sets):
            qkv = (torch.randn(B * T, 3 * D, device="cuda") * 0.5).bfloat16()
            kc = (torch.randn(B, H, Tmax, hd, device="cuda") * 0.5).bfloat16()
            vc = (torch.randn(B, H, Tmax, hd, device="cuda") * 0.5).bfloat16()
            sets.append((qkv, kc, vc))
        out = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
        out2 = torch.empty_like(out)
        s12 = (C.c_longlong * 12)(3 * D * T, 3 * D, hd, H * Tmax * hd, hd, Tmax * hd, H * Tmax * hd, hd, Tmax * hd,
                                  D * T, D, hd)

        def run_old(s, o):
            _lib.check(lib.ovla_flash_attention(P(s[0]), P(s[1]), P(s[2]), P(o), s12, B, H, T, T, hd, 1, None))

        def run_tc(s, o):
            _lib.check(lib.ovla_prefill_attention_tc(P(s[0]), C.c_longlong(3 * D), P(s[1]), P(s[2]), P(o),
                                                     C.c_longlong(D), B, H, T, Tmax, None))
        run_old(sets[0], out); run_tc(sets[0], out2); torch.cuda.synchronize()
        diff = float((out.float() - out2.float()).abs().max())
        res = {"B": B, "T": T, "max_abs_diff_between_kernels": diff}
        flops = 4.0 * B * H * hd * (T * (T + 1) / 2)
        for name, fn in (("mma_sync", run_old), ("tcgen05", run_tc)):
            for s in sets: fn(s, out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5 if B >= 64 else 50
            e0.record()
            for _ in range(reps):
                for s in sets: fn(s, out)
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (reps * n_sets)
            res[name] = {"us": round(us, 1), "causal_tflops": round(flops / us / 1e6, 1)}
        print(json.dumps(res), flush=True)

    # ViT tower shapes: packed qkv, non-causal; DINOv2 ViT-L/14 (head_dim 64, 261 tokens), SigLIP SO400M (72, 256)
    for H, hd, T, B in ((16, 64, 261, 256), (16, 64, 261, 16), (16, 72, 256, 256), (16, 72, 256, 16)):
      D = H * hd
      if True:
          sets = [(torch.randn(B * T, 3 * D, device="cuda") * 0.5).bfloat16() for _ in range(3)]
          out = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
          out2 = torch.empty_like(out)
          s12 = (C.c_longlong * 12)(3 * D * T, 3 * D, hd, 3 * D * T, 3 * D, hd, 3 * D * T, 3 * D, hd, D * T, D, hd)

          def run_old(x, o):
              _lib.check(lib.ovla_flash_attention(P(x), C.c_void_p(x.data_ptr() + 2 * D), C.c_void_p(x.data_ptr() + 4 * D),
                                                  P(o), s12, B, H, T, T, hd, 0, None))

          def run_tc(x, o):
              _lib.check(lib.ovla_attention_tc_qkv(P(x), C.c_longlong(3 * D), P(o), C.c_longlong(D), B, H, T, hd, 0, None))
          run_old(sets[0], out); run_tc(sets[0], out2); torch.cuda.synchronize()
          res = {"shape": "dinov2" if hd == 64 else "siglip", "B": B, "T": T, "max_abs_diff_between_kernels": float((out.float() - out2.float()).abs().max())}
          flops = 4.0 * B * H * hd * T * T
          for name, fn in (("mma_sync", run_old), ("tcgen05", run_tc)):
              for x in sets: fn(x, out)
              torch.cuda.synchronize()
              e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
              reps = 10
              e0.record()
              for _ in range(reps):
                  for x in sets: fn(x, out)
              e1.record(); torch.cuda.synchronize()
              us = e0.elapsed_time(e1) * 1e3 / (reps * 3)
              res[name] = {"us": round(us, 1), "tflops": round(flops / us / 1e6, 1)}
          print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
