"""Large-shape check of both prefill attention kernels against torch SDPA in fp32 on the GPU (checker only)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvla_probe_b200 import _lib
lib = _lib.load()
P = lambda t: C.c_void_p(t.data_ptr())
H, hd, Tmax, T = 32, 128, 320, 283
D = H * hd
for B, scale in ((64, 0.5), (64, 1.0), (128, 0.5)):
    torch.manual_seed(B)
    qkv = (torch.randn(B * T, 3 * D, device="cuda") * scale).bfloat16()
    kc = (torch.randn(B, H, Tmax, hd, device="cuda") * scale).bfloat16()
    vc = (torch.randn(B, H, Tmax, hd, device="cuda") * scale).bfloat16()
    q = qkv.view(B, T, 3, H, hd)[:, :, 0].permute(0, 2, 1, 3).float()
    ref = torch.nn.functional.scaled_dot_product_attention(q, kc[:, :, :T].float(), vc[:, :, :T].float(), is_causal=True)
    ref = ref.permute(0, 2, 1, 3).reshape(B * T, D)
    s12 = (C.c_longlong * 12)(3 * D * T, 3 * D, hd, H * Tmax * hd, hd, Tmax * hd, H * Tmax * hd, hd, Tmax * hd, D * T, D, hd)
    res = {"B": B, "scale": scale}
    for rep in range(3):
        o1 = torch.zeros(B * T, D, device="cuda", dtype=torch.bfloat16)
        o2 = torch.zeros_like(o1)
        _lib.check(lib.ovla_flash_attention(P(qkv), P(kc), P(vc), P(o1), s12, B, H, T, T, hd, 1, None))
        _lib.check(lib.ovla_prefill_attention_tc(P(qkv), C.c_longlong(3 * D), P(kc), P(vc), P(o2), C.c_longlong(D), B, H, T, Tmax, None))
        torch.cuda.synchronize()
        for name, o in (("mma_sync", o1), ("tcgen05", o2)):
            d = (o.float() - ref).abs()
            idx = int(d.argmax())
            row, col = idx // D, idx % D
            res[f"{name}_{rep}"] = {"max": float(d.max()), "n_gt_0.03": int((d > 0.03).sum()), "b": row // T, "t": row % T, "h": col // hd, "c": col % hd,
                                    "ref": float(ref[row, col]), "got": float(o[row, col])}
    print(json.dumps(res), flush=True)
