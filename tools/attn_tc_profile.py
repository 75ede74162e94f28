"""A few launches of the tcgen05 prefill attention at the Llama bs=256 shape, for ncu (run under gpurun)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvla_probe_b200 import _lib
lib = _lib.load()
P = lambda t: C.c_void_p(t.data_ptr())
B, H, hd, Tmax, T = 256, 32, 128, 320, 283
D = H * hd
qkv = (torch.randn(B * T, 3 * D, device="cuda") * 0.5).bfloat16()
kc = (torch.randn(B, H, Tmax, hd, device="cuda") * 0.5).bfloat16()
vc = (torch.randn(B, H, Tmax, hd, device="cuda") * 0.5).bfloat16()
out = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    _lib.check(lib.ovla_prefill_attention_tc(P(qkv), C.c_longlong(3 * D), P(kc), P(vc), P(out), C.c_longlong(D), B, H, T, Tmax, None))
torch.cuda.synchronize()
print("ok")
