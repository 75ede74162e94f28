"""Where an epilogue warp of the bf16 GEMM spends its cycles per tile (debug build, prints from the kernel):
    python -m openvla_probe_b200.build -DOVLA_DBG_EPI_TIMELINE=1 --variant=epitl
    OVLA_B200_LIB=openvla_probe_b200/libovla_b200_epitl.so python tools/gemm_epilogue_timeline.py
One launch per shape after a warm-up launch; the kernel prints, for warp 4 and warp 11 of the first and last CTA, the
mean cycles per tile spent waiting for the accumulator, in tcgen05.ld, in the residual wait + read, in the epilogue math,
waiting for the staging slot, in the store (st.shared + fence + TMA issue) and in the release / loop overhead."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvla_probe_b200 import _lib
from gemm_epilogue_bench import SHAPES
lib = _lib.load()
P = lambda t: C.c_void_p(t.data_ptr())
for name, M, N, K, bias, gelu, scale, resid in SHAPES[:8]:
    A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    b = (torch.randn(N, device="cuda") * 0.1).bfloat16()
    sc = (1 + 0.1 * torch.randn(N, device="cuda")).bfloat16()
    x = torch.randn(M, N, device="cuda").bfloat16()
    epi = _lib.GemmEpilogue()
    if bias: epi.bias_bf16 = b.data_ptr()
    epi.gelu = gelu
    if scale: epi.scale_bf16 = sc.data_ptr()
    if resid: epi.resid_bf16, epi.ld_resid = x.data_ptr(), N
    print(f"== {name} M={M} N={N} K={K} (main loop of a 256x256 tile: {K // 64 * 512} tensor cycles)", flush=True)
    _lib.check(lib.ovla_gemm(P(A), K, P(W), K, M, N, K, 0, 0, P(x), N, C.byref(epi), 0, 0, None))
    torch.cuda.synchronize()
    del A, W, x
