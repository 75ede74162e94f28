"""Operator-level GPU parity: every C-ABI kernel against the CPU oracle's definition of the same op on seeded inputs,
including the awkward shapes of the real model (K = 588->592, N = 4304, head_dim 72, 261 tokens, ragged tails).

Tolerances (stated): bf16 outputs are compared with the oracle computed in fp32 from the same bf16 inputs and then
rounded to bf16 -- max error <= 2 bf16 ulps of the row scale (accumulation order differs, rounding points do not);
integer / index results (argmax, de-tokeniser table) are bit-exact; fp64 de-tokenise / un-normalise is bit-exact.
"""
import ctypes as C
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import openvla_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from openvla_probe_b200 import _lib

    return _lib, _lib.load()


def P(t):
    return C.c_void_p(t.data_ptr())


def bf(x):
    return x.to(torch.bfloat16)


def close_bf16(got, ref32, ulps=2.0):
    """|got - bf16(ref)| <= ulps * 2^-8 * max(|ref| row scale)."""
    got, ref32 = got.float().cpu(), ref32.float().cpu()
    scale = ref32.abs().amax(dim=-1, keepdim=True).clamp_min(1e-6)
    err = (got - ref32).abs() / scale
    return float(err.max()) <= ulps * 2.0 ** -8, float(err.max())


# ----------------------------------------------------------------------------------------------- GEMM epilogues
@pytest.mark.parametrize("M,N,K,bn,cg", [(512, 4304, 1152, 0, 0), (261 * 2, 1024, 4096, 128, 1), (256, 1152, 592, 64, 1),
                                         (1024, 1024, 1024, 256, 2), (130, 3456, 1152, 128, 2)])
def test_gemm_bias_gelu_scale_residual(L, M, N, K, bn, cg):
    _lib, lib = L
    g = torch.Generator().manual_seed(M + N)
    A = bf(torch.randn(M, K, generator=g) * 0.5)
    W = bf(torch.randn(N, K, generator=g) * 0.05)
    bias, scale, resid = bf(torch.randn(N, generator=g) * 0.1), bf(1 + 0.1 * torch.randn(N, generator=g)), bf(torch.randn(M, N, generator=g))
    for gelu, use_scale, use_res in [(0, 0, 0), (1, 0, 0), (0, 1, 1), (0, 0, 1)]:
        ref = F.linear(A.float(), W.float(), bias.float()).bfloat16().float()
        if gelu:
            ref = F.gelu(ref).bfloat16().float()
        if use_scale:
            ref = (ref * scale.float()).bfloat16().float()
        if use_res:
            ref = ref + resid.float()
        Ad, Wd, bd, sd, rd = A.cuda(), W.cuda(), bias.cuda(), scale.cuda(), resid.cuda()
        out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
        epi = _lib.GemmEpilogue()
        epi.bias_bf16, epi.gelu = bd.data_ptr(), gelu
        if use_scale:
            epi.scale_bf16 = sd.data_ptr()
        if use_res:
            epi.resid_bf16, epi.ld_resid = rd.data_ptr(), N
        _lib.check(lib.ovla_gemm(P(Ad), C.c_longlong(K), P(Wd), C.c_longlong(K), M, N, K, 0, 0, P(out), C.c_longlong(N),
                                 C.byref(epi), bn, cg, None))
        ok, e = close_bf16(out, ref, ulps=3.0)
        assert ok, (gelu, use_scale, use_res, e)


@pytest.mark.parametrize("M", [1, 3, 8, 77, 300])
def test_swiglu_and_fp32_logits_gemm_and_gemv_agree_with_oracle(L, M):
    _lib, lib = L
    g = torch.Generator().manual_seed(M)
    K, I, V = 256, 704, 32064
    x = bf(torch.randn(M, K, generator=g))
    Wg, Wu = bf(torch.randn(I, K, generator=g) * 0.06), bf(torch.randn(I, K, generator=g) * 0.06)
    ref = (F.silu(F.linear(x.float(), Wg.float()).bfloat16().float()).bfloat16().float()
           * F.linear(x.float(), Wu.float()).bfloat16().float())
    Wi = torch.stack([Wg.view(I // 32, 32, K), Wu.view(I // 32, 32, K)], 1).reshape(2 * I, K).contiguous()
    out = torch.empty(M, I, dtype=torch.bfloat16, device="cuda")
    epi = _lib.GemmEpilogue()
    xd, Wd = x.cuda(), Wi.cuda()
    if M <= 8:
        _lib.check(lib.ovla_gemv(P(xd), C.c_longlong(K), P(Wd), C.c_longlong(K), M, 2 * I, K, 1, P(out), C.c_longlong(I),
                                 C.byref(epi), None))
    else:
        _lib.check(lib.ovla_gemm(P(xd), C.c_longlong(K), P(Wd), C.c_longlong(K), M, 2 * I, K, 1, 0, P(out), C.c_longlong(I),
                                 C.byref(epi), 0, 0, None))
    ok, e = close_bf16(out, ref, ulps=3.0)
    assert ok, e
    # lm_head: fp32 storage of bf16-rounded logits (HF `.float()`), vocab 32064 = 501 * 64
    Wl = bf(torch.randn(V, K, generator=g) * 0.05)
    refl = F.linear(x.float(), Wl.float()).bfloat16().float()
    lg = torch.empty(M, V, dtype=torch.float32, device="cuda")
    epi = _lib.GemmEpilogue()
    epi.round_bf16 = 1
    Wld = Wl.cuda()
    if M <= 8:
        _lib.check(lib.ovla_gemv(P(xd), C.c_longlong(K), P(Wld), C.c_longlong(K), M, V, K, 2, P(lg), C.c_longlong(V),
                                 C.byref(epi), None))
    else:
        _lib.check(lib.ovla_gemm(P(xd), C.c_longlong(K), P(Wld), C.c_longlong(K), M, V, K, 2, 0, P(lg), C.c_longlong(V),
                                 C.byref(epi), 0, 0, None))
    assert torch.equal(lg.cpu().bfloat16().float(), lg.cpu())          # values are exactly representable in bf16
    ok, e = close_bf16(lg, refl, ulps=2.0)
    assert ok, e


def test_gemm_rejects_bad_shapes(L):
    _lib, lib = L
    a = torch.zeros(16, 24, dtype=torch.bfloat16, device="cuda")
    o = torch.zeros(16, 24, dtype=torch.bfloat16, device="cuda")
    epi = _lib.GemmEpilogue()
    assert lib.ovla_gemm(P(a), C.c_longlong(24), P(a), C.c_longlong(24), 16, 12, 24, 0, 0, P(o), C.c_longlong(12), C.byref(epi), 0, 0, None) != 0
    assert b"multiple of 8" in lib.ovla_last_error()
    assert lib.ovla_gemm(P(a), C.c_longlong(20), P(a), C.c_longlong(24), 16, 16, 20, 0, 0, P(o), C.c_longlong(16), C.byref(epi), 0, 0, None) != 0
    assert lib.ovla_gemm(P(a), C.c_longlong(24), P(a), C.c_longlong(24), 0, 16, 24, 0, 0, P(o), C.c_longlong(16), C.byref(epi), 0, 0, None) != 0


# ----------------------------------------------------------------------------------------------- norms
@pytest.mark.parametrize("rows,D", [(5, 128), (261, 1024), (300, 1152), (7, 4096), (1, 256)])
def test_layernorm_and_rmsnorm(L, rows, D):
    _lib, lib = L
    g = torch.Generator().manual_seed(D + rows)
    x = bf(torch.randn(rows, D, generator=g) * 3 + 0.5)
    w, b = bf(1 + 0.1 * torch.randn(D, generator=g)), bf(0.1 * torch.randn(D, generator=g))
    out = torch.empty(rows, D, dtype=torch.bfloat16, device="cuda")
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()          # keep device copies alive across the calls
    _lib.check(lib.ovla_layernorm(P(xd), C.c_longlong(D), P(wd), P(bd), C.c_float(1e-6), P(out),
                                  C.c_longlong(D), rows, D, None))
    ref = F.layer_norm(x.float(), (D,), w.float(), b.float(), eps=1e-6)
    ok, e = close_bf16(out, ref)
    assert ok, e
    _lib.check(lib.ovla_rmsnorm(P(xd), C.c_longlong(D), P(wd), C.c_float(1e-6), P(out), C.c_longlong(D), rows,
                                D, None))
    ref = O.rms_norm(x, w, 1e-6)                      # oracle in bf16: the reference's exact rounding points
    assert float((out.cpu().float() - ref.float()).abs().max()) <= 2.0 ** -7 * float(ref.float().abs().max())
    mism = (out.cpu() != ref).float().mean().item()
    assert mism < 0.01                                # identical up to fp32 reduction-order effects on rsqrt


# ----------------------------------------------------------------------------------------------- attention
@pytest.mark.parametrize("B,H,T,hd,causal", [(2, 2, 21, 64, 0), (1, 16, 261, 64, 0), (2, 16, 256, 72, 0), (1, 3, 16, 72, 0),
                                             (2, 4, 288, 128, 1), (3, 2, 25, 128, 1), (1, 2, 130, 128, 1)])
def test_flash_attention_vs_oracle(L, B, H, T, hd, causal):
    _lib, lib = L
    g = torch.Generator().manual_seed(T * hd)
    D = H * hd
    qkv = bf(torch.randn(B, T, 3, H, hd, generator=g))
    q, k, v = [qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3)]           # [B,H,T,hd]
    ref = O._sdpa(q.float(), k.float(), v.float(), causal=bool(causal)).permute(0, 2, 1, 3).reshape(B, T, D)
    buf = qkv.reshape(B * T, 3 * D).contiguous().cuda()
    out = torch.empty(B * T, D, dtype=torch.bfloat16, device="cuda")
    s = (C.c_longlong * 12)(3 * D * T, 3 * D, hd, 3 * D * T, 3 * D, hd, 3 * D * T, 3 * D, hd, D * T, D, hd)
    kq = C.c_void_p(buf.data_ptr() + 2 * D)
    vq = C.c_void_p(buf.data_ptr() + 4 * D)
    _lib.check(lib.ovla_flash_attention(P(buf), kq, vq, P(out), s, B, H, T, T, hd, causal, None))
    got = out.view(B, T, D).float().cpu()
    assert float((got - ref).abs().max()) <= 0.03 * float(ref.abs().max()) + 1e-3
    assert float((got - ref).norm() / ref.norm()) < 6e-3


@pytest.mark.parametrize("B,H,T,Tmax,qscale", [(2, 4, 288, 304, 1.0), (3, 2, 25, 64, 1.0), (1, 2, 130, 130, 1.0),
                                               (2, 3, 128, 128, 1.0), (1, 2, 283, 320, 1.0), (2, 2, 400, 401, 1.0),
                                               (2, 2, 283, 320, 6.0), (1, 1, 1, 16, 1.0), (1, 2, 700, 704, 3.0)])
def test_prefill_attention_tcgen05_vs_oracle(L, B, H, T, Tmax, qscale):
    """tcgen05 causal prefill attention (queries in the fused qkv buffer, K/V in the cache layout [B,H,Tmax,128])
    against the oracle SDPA, to the same tolerance as the mma.sync kernel; rows >= T of the cache hold other data.
    qscale > 1 makes the row maxima jump between key chunks, which exercises the lazy O / l rescale in TMEM."""
    _lib, lib = L
    hd = 128
    g = torch.Generator().manual_seed(T * 7 + H)
    D = H * hd
    qkv = torch.randn(B, T, 3, H, hd, generator=g)
    qkv[:, :, 0] *= qscale
    qkv = bf(qkv)
    q, k, v = [qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3)]           # [B,H,T,hd]
    ref = O._sdpa(q.float(), k.float(), v.float(), causal=True).permute(0, 2, 1, 3).reshape(B, T, D)
    buf = qkv.reshape(B * T, 3 * D).contiguous().cuda()
    kc = bf(torch.randn(B, H, Tmax, hd, generator=g) * 3)                    # finite junk beyond T
    vc = bf(torch.randn(B, H, Tmax, hd, generator=g) * 3)
    kc[:, :, :T] = k
    vc[:, :, :T] = v
    kcd, vcd = kc.cuda(), vc.cuda()
    out = torch.zeros(B * T, D, dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.ovla_prefill_attention_tc(P(buf), C.c_longlong(3 * D), P(kcd), P(vcd), P(out), C.c_longlong(D),
                                             B, H, T, Tmax, None))
    torch.cuda.synchronize()
    got = out.view(B, T, D).float().cpu()
    assert float((got - ref).abs().max()) <= 0.006 * float(ref.abs().max()) + 1e-3
    assert float((got - ref).norm() / ref.norm()) < 4e-3


@pytest.mark.parametrize("B,H,T,hd,causal,qscale", [(2, 2, 21, 64, 0, 1.0), (1, 16, 261, 64, 0, 1.0), (3, 4, 256, 64, 0, 1.0),
                                                    (2, 3, 261, 64, 0, 5.0), (2, 2, 130, 64, 1, 1.0),
                                                    (2, 2, 200, 128, 0, 1.0), (1, 2, 283, 128, 1, 4.0),
                                                    (2, 16, 256, 72, 0, 1.0), (1, 3, 16, 72, 0, 1.0), (2, 2, 261, 72, 0, 5.0),
                                                    (3, 5, 70, 72, 0, 1.0)])
def test_attention_tcgen05_packed_qkv_vs_oracle(L, B, H, T, hd, causal, qscale):
    """Same kernel on a packed [B*T, 3D] qkv buffer (the ViT towers' layout), head_dim 64 / 128, and 72 (SigLIP): there
    each operand is a 64-column slab plus a 16-column slab whose last 8 columns the TMA box zero-fills."""
    _lib, lib = L
    g = torch.Generator().manual_seed(T * hd + causal)
    D = H * hd
    qkv = torch.randn(B, T, 3, H, hd, generator=g)
    qkv[:, :, 0] *= qscale
    qkv = bf(qkv)
    q, k, v = [qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3)]
    ref = O._sdpa(q.float(), k.float(), v.float(), causal=bool(causal)).permute(0, 2, 1, 3).reshape(B, T, D)
    buf = qkv.reshape(B * T, 3 * D).contiguous().cuda()
    out = torch.zeros(B * T, D, dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.ovla_attention_tc_qkv(P(buf), C.c_longlong(3 * D), P(out), C.c_longlong(D), B, H, T, hd, causal, None))
    torch.cuda.synchronize()
    got = out.view(B, T, D).float().cpu()
    assert float((got - ref).abs().max()) <= 0.006 * float(ref.abs().max()) + 1e-3
    assert float((got - ref).norm() / ref.norm()) < 4e-3


def test_attention_tcgen05_many_waves_deterministic(L):
    """Llama prefill shape (32 heads, 283 tokens) with enough (batch, head) tiles for many waves of CTAs: every
    element within 2 bf16 ulps of the fp32 oracle and two launches bit-identical (barrier-phase races show up as rare
    run-to-run differences)."""
    _lib, lib = L
    B, H, T, Tmax, hd = 12, 32, 283, 320, 128
    D = H * hd
    g = torch.Generator().manual_seed(5)
    qkv = bf(torch.randn(B, T, 3, H, hd, generator=g))
    q, k, v = [qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3)]
    ref = O._sdpa(q.float(), k.float(), v.float(), causal=True).permute(0, 2, 1, 3).reshape(B * T, D)
    buf = qkv.reshape(B * T, 3 * D).contiguous().cuda()
    kc = torch.zeros(B, H, Tmax, hd, dtype=torch.bfloat16)
    vc = torch.zeros(B, H, Tmax, hd, dtype=torch.bfloat16)
    kc[:, :, :T] = k
    vc[:, :, :T] = v
    kcd, vcd = kc.cuda(), vc.cuda()
    outs = []
    for _ in range(3):
        out = torch.zeros(B * T, D, dtype=torch.bfloat16, device="cuda")
        _lib.check(lib.ovla_prefill_attention_tc(P(buf), C.c_longlong(3 * D), P(kcd), P(vcd), P(out), C.c_longlong(D),
                                                 B, H, T, Tmax, None))
        torch.cuda.synchronize()
        outs.append(out.cpu())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    err = (outs[0].float() - ref).abs()
    assert float(err.max()) <= 2 * 2.0 ** -8 * float(ref.abs().max())
    assert int((err > 2.0 ** -7 * ref.abs().clamp_min(1.0)).sum()) == 0


@pytest.mark.parametrize("hd,T,H,B", [(64, 261, 16, 40), (72, 256, 16, 40)])
def test_vit_attention_tcgen05_many_waves_deterministic(L, hd, T, H, B):
    """The two ViT shapes (DINOv2: head_dim 64, 261 tokens; SigLIP: head_dim 72, 256 tokens) with 640 (batch, head)
    sequences = many waves of the persistent CTAs: three launches bit-identical and within 2 bf16 ulps of the oracle."""
    _lib, lib = L
    D = H * hd
    g = torch.Generator().manual_seed(hd)
    qkv = bf(torch.randn(B, T, 3, H, hd, generator=g))
    q, k, v = [qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3)]
    ref = O._sdpa(q.float(), k.float(), v.float(), causal=False).permute(0, 2, 1, 3).reshape(B * T, D)
    buf = qkv.reshape(B * T, 3 * D).contiguous().cuda()
    outs = []
    for _ in range(3):
        out = torch.zeros(B * T, D, dtype=torch.bfloat16, device="cuda")
        _lib.check(lib.ovla_attention_tc_qkv(P(buf), C.c_longlong(3 * D), P(out), C.c_longlong(D), B, H, T, hd, 0, None))
        torch.cuda.synchronize()
        outs.append(out.cpu())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    err = (outs[0].float() - ref).abs()
    assert float(err.max()) <= 2 * 2.0 ** -8 * float(ref.abs().max())


@pytest.mark.parametrize("M,N,K", [(1100, 4096, 4096), (1100, 4096, 11008), (1024, 1152, 4304), (20 * 261, 1024, 4096)])
def test_gemm_inplace_residual_ragged_multi_tile_bit_exact(L, M, N, K):
    """The engine always calls the residual epilogue IN PLACE (out == resid: o_proj / down_proj / ViT proj / fc2).  At
    ragged M with more tiles than workers (M = 1100, N = 4096 as CTA pairs: 80 tile pairs on 74 workers, the second
    CTA of the last pair fully out of bounds) a worker runs several tiles and its residual staging slot is refilled by
    TMA while the previous chunk is still being read: round 1 ordered that refill with __syncwarp only and, rarely, a
    lane read the NEXT chunk's residual (run-to-run differences of whole 16-byte pieces; found by
    tools/determinism_bisect.py, fixed with a proxy fence).  Repeated in-place launches must equal the out-of-place
    result bit for bit for the heuristic tile and for every forced tile shape, and must not write outside [0, M)."""
    _lib, lib = L
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = bf(torch.randn(M, K, generator=g, device="cuda") * 0.5)
    W = bf(torch.randn(N, K, generator=g, device="cuda") * 0.03)
    X = bf(torch.randn(M, N, generator=g, device="cuda"))

    def call(out, resid, bn, cg):
        epi = _lib.GemmEpilogue()
        epi.resid_bf16, epi.ld_resid = resid.data_ptr(), N
        _lib.check(lib.ovla_gemm(P(A), C.c_longlong(K), P(W), C.c_longlong(K), M, N, K, 0, 0, P(out), C.c_longlong(N),
                                 C.byref(epi), bn, cg, None))

    ref = torch.empty_like(X)
    call(ref, X, 0, 0)
    torch.cuda.synchronize()
    want = (A.float() @ W.float().t()).bfloat16().float() + X.float()
    ok, e = close_bf16(ref, want, ulps=3.0)
    assert ok, e
    for bn, cg in [(0, 0), (256, 2), (128, 1), (128, 2), (64, 1)]:
        for rep in range(6):
            buf = torch.full((M + 64, N), 7.0, dtype=torch.bfloat16, device="cuda")
            buf[:M].copy_(X)
            call(buf[:M], buf[:M], bn, cg)
            torch.cuda.synchronize()
            n_bad = int((buf[:M] != ref).sum())
            assert n_bad == 0, (bn, cg, rep, n_bad)
            assert bool((buf[M:] == 7.0).all()), (bn, cg, rep, "stored outside [0, M)")


@pytest.mark.parametrize("M,N,K,mode", [(8192 + 77, 4096, 4096, 0), (9000, 1152, 4304, 0), (8448, 2816, 2048, 1),
                                          (8300, 4096, 5632, 0)])
def test_gemm_rasterisation_and_wave_alignment_do_not_change_results(L, M, N, K, mode):
    """Round 2: multi-wave GEMMs (M >= 8192) align the TMA producers of the persistent grid at tile / mid-tile boundaries
    through two global counters (gemm.cuh `GemmShape::sync`) and walk the tiles in row groups inside column
    super-groups, optionally serpentine.  None may change a bit of the output: every (group, super-group, alignment distance, order) -- forced
    through ovla_debug_gemm_raster -- must equal the un-aligned, single-super-group launch, for ragged M, an N that is
    not a multiple of the tile, the in-place residual epilogue and SwiGLU; the counters must be left clean so that
    launches of different grids follow each other on one stream; nothing may be stored outside [0, M)."""
    _lib, lib = L
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = bf(torch.randn(M, K, generator=g, device="cuda") * 0.5)
    W = bf(torch.randn(N, K, generator=g, device="cuda") * 0.03)
    n_out = N // 2 if mode == 1 else N
    X = bf(torch.randn(M, n_out, generator=g, device="cuda"))

    def call(buf, bn=0, cg=0):
        epi = _lib.GemmEpilogue()
        if mode == 0:
            epi.resid_bf16, epi.ld_resid = buf.data_ptr(), n_out
        _lib.check(lib.ovla_gemm(P(A), K, P(W), K, M, N, K, mode, 0, P(buf), n_out, C.byref(epi), bn, cg, None))

    try:
        lib.ovla_debug_gemm_raster(16, 0, 0, 0, 0, 0)    # round-1 rasterisation, no alignment
        ref = X.clone()
        call(ref)
        torch.cuda.synchronize()
        if mode == 0:
            want = (A.float() @ W.float().t()).bfloat16().float() + X.float()
            ok, e = close_bf16(ref, want, ulps=3.0)
            assert ok, e
        num_k = (K + 63) // 64
        cfgs = [(-1, -1, -1, -1, -1, -1), (2, 0, 0, 0, num_k, 1), (2, 3, 0, 0, 7, 1), (5, 2, 1, 2, 1, 0), (16, 1, 0, 0, num_k // 2, 1),
                (1, 0, 2, 0, 10 ** 6, 0), (3, 1000, 0, 0, 0, 1)]
        for ci, cfg in enumerate(cfgs):
            lib.ovla_debug_gemm_raster(*cfg)
            for bn, cg in [(0, 0), (128, 1)] if ci % 2 == 0 else [(0, 0)]:
                buf = torch.full((M + 64, n_out), 7.0, dtype=torch.bfloat16, device="cuda")
                buf[:M].copy_(X)
                call(buf[:M], bn, cg)
                torch.cuda.synchronize()
                assert int((buf[:M] != ref).sum()) == 0, (cfg, bn, cg)
                assert bool((buf[M:] == 7.0).all()), (cfg, bn, cg, "stored outside [0, M)")
    finally:
        lib.ovla_debug_gemm_raster(-1, -1, -1, -1, -1, -1)


def test_wave_aligned_gemms_on_two_streams_finish_and_agree(L):
    """Two wave-aligned GEMMs launched at the same time on two streams each want every SM (1 CTA per SM): the second
    grid becomes resident piecemeal while the first one drains, so its early CTAs sit at the first alignment point
    until the rest arrives.  Each stream has its own counter words and a waiting producer gives up after ~4 M cycles,
    so nothing can dead-lock; the results must equal the serial ones bit for bit and the whole thing must take
    milliseconds, not time-outs."""
    import time
    _lib, lib = L
    M, N, K = 8192 + 256, 4096, 4096
    g = torch.Generator(device="cuda").manual_seed(7)
    A = [bf(torch.randn(M, K, generator=g, device="cuda") * 0.5) for _ in range(2)]
    W = [bf(torch.randn(N, K, generator=g, device="cuda") * 0.03) for _ in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]

    def call(i, out, st):
        epi = _lib.GemmEpilogue()
        _lib.check(lib.ovla_gemm(P(A[i]), K, P(W[i]), K, M, N, K, 0, 0, P(out), N, C.byref(epi), 0, 0,
                                 C.c_void_p(st.cuda_stream) if st is not None else None))

    ref = [torch.empty(M, N, dtype=torch.bfloat16, device="cuda") for _ in range(2)]
    for i in range(2):
        call(i, ref[i], None)
    torch.cuda.synchronize()
    out = [torch.empty_like(r) for r in ref]
    for i in range(2):
        call(i, out[i], streams[i])       # first use of a stream allocates its counter words
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for rep in range(6):
        for i in range(2):
            call(i, out[i], streams[i])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    for i in range(2):
        assert torch.equal(out[i], ref[i]), i
    assert dt < 0.25, f"12 GEMMs of 0.28 TFLOP (~3 ms of work) took {dt:.3f} s: alignment waits are timing out"


@pytest.mark.parametrize("B,ctx", [(1, 1), (2, 37), (3, 290), (5, 64)])
def test_decode_rope_attention_and_cache_append(L, B, ctx):
    """Fused RoPE + KV append + 1-query attention == oracle llama attention step with a KV cache."""
    _lib, lib = L
    from openvla_probe_b200.engine import rope_tables

    H, hd, Tmax = 4, 128, 304
    D = H * hd
    g = torch.Generator().manual_seed(ctx)
    pos = ctx - 1
    kc = bf(torch.randn(B, H, Tmax, hd, generator=g))
    vc = bf(torch.randn(B, H, Tmax, hd, generator=g))
    qkv = bf(torch.randn(B, 3 * D, generator=g))
    cos, sin = rope_tables(hd, 10000.0, Tmax)
    d = O.VLADims(llm_dim=D, llm_heads=H)
    c_full, s_full = O.rope_cos_sin(d, torch.tensor([pos]), torch.bfloat16)
    q = qkv[:, :D].view(B, 1, H, hd).transpose(1, 2)
    k = qkv[:, D:2 * D].view(B, 1, H, hd).transpose(1, 2)
    v = qkv[:, 2 * D:].view(B, 1, H, hd).transpose(1, 2)
    q_r = q * c_full + O._rotate_half(q) * s_full
    k_r = k * c_full + O._rotate_half(k) * s_full
    K_all = torch.cat([kc[:, :, :pos], k_r], 2)
    V_all = torch.cat([vc[:, :, :pos], v], 2)
    ref = O._sdpa(q_r, K_all, V_all, causal=True, q_offset=pos).transpose(1, 2).reshape(B, D)
    kcd, vcd, qd, cd, sd = kc.cuda(), vc.cuda(), qkv.cuda(), cos.cuda(), sin.cuda()
    out = torch.empty(B, D, dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.ovla_decode_rope_attention(P(qd), C.c_longlong(3 * D), P(cd), P(sd), pos, P(kcd),
                                              P(vcd), B, H, hd, Tmax, P(out), C.c_longlong(D), None))
    assert torch.equal(kcd[:, :, pos].cpu(), k_r[:, :, 0])               # rotated key appended bit-exactly
    assert torch.equal(vcd[:, :, pos].cpu(), v[:, :, 0])
    assert torch.equal(kcd[:, :, :pos].cpu(), kc[:, :, :pos])            # older entries untouched
    got = out.float().cpu()
    assert float((got - ref.float()).abs().max()) <= 0.03 * float(ref.float().abs().max()) + 1e-3


def test_rope_kv_prefill_bit_exact(L):
    _lib, lib = L
    from openvla_probe_b200.engine import rope_tables

    B, T, H, hd, Tmax = 2, 19, 3, 128, 40
    D = H * hd
    g = torch.Generator().manual_seed(0)
    qkv = bf(torch.randn(B * T, 3 * D, generator=g))
    cos, sin = rope_tables(hd, 10000.0, Tmax)
    d = O.VLADims(llm_dim=D, llm_heads=H)
    c, s = O.rope_cos_sin(d, torch.arange(T), torch.bfloat16)
    x = qkv.view(B, T, 3, H, hd)
    q_ref = x[:, :, 0] * c[None, :, None] + O._rotate_half(x[:, :, 0]) * s[None, :, None]
    k_ref = x[:, :, 1] * c[None, :, None] + O._rotate_half(x[:, :, 1]) * s[None, :, None]
    buf = qkv.clone().cuda()
    kc = torch.zeros(B, H, Tmax, hd, dtype=torch.bfloat16, device="cuda")
    vc = torch.zeros_like(kc)
    cd, sd = cos.cuda(), sin.cuda()
    _lib.check(lib.ovla_rope_kv(P(buf), B, T, H, hd, 0, P(cd), P(sd), P(kc), P(vc), Tmax, None))
    assert torch.equal(buf.cpu().view(B, T, 3, H, hd)[:, :, 0], q_ref)
    assert torch.equal(kc.cpu()[:, :, :T].permute(0, 2, 1, 3), k_ref)
    assert torch.equal(vc.cpu()[:, :, :T].permute(0, 2, 1, 3), x[:, :, 2])
    assert lib.ovla_rope_kv(P(buf), B, T, H, hd, 30, P(cd), P(sd), P(kc), P(vc), Tmax, None) != 0  # overflow


# ----------------------------------------------------------------------------------------------- capture / argmax / detok
@pytest.mark.parametrize("B,T,D,n", [(1, 7, 256, 7), (3, 288, 4096, 287), (2, 33, 264, 1)])
def test_pool_tokens_mean_and_final(L, B, T, D, n):
    _lib, lib = L
    x = bf(torch.randn(B, T, D, generator=torch.Generator().manual_seed(T)))
    out = torch.empty(B, D, dtype=torch.float32, device="cuda")
    xd = x.cuda()
    for mode, ref in ((0, x[:, :n].float().mean(1)), (1, x[:, n - 1].float())):
        _lib.check(lib.ovla_pool_tokens(P(xd), C.c_longlong(T * D), C.c_longlong(D), B, n, D, mode, P(out),
                                        C.c_longlong(D), None))
        assert torch.allclose(out.cpu(), ref, rtol=0, atol=2e-6 * max(1.0, float(ref.abs().max())) * math.sqrt(n))
    assert lib.ovla_pool_tokens(P(xd), C.c_longlong(T * D), C.c_longlong(D), B, 0, D, 0, P(out), C.c_longlong(D), None) != 0


def test_argmax_bit_exact_ties_nan_inf(L):
    _lib, lib = L
    g = torch.Generator().manual_seed(0)
    V = 32064
    x = torch.randn(9, V, generator=g).bfloat16().float()             # bf16-rounded values -> many exact ties
    x[1, 31990] = x[1].max()
    x[1, 17] = x[1].max()                                             # tie: lowest index wins
    x[2, :] = 0.0                                                     # all equal -> index 0
    x[3, 32063] = float("inf")
    x[4, 5] = float("nan")
    x[4, 9] = float("inf")                                            # NaN counts as maximum
    x[5, :] = float("-inf")
    x[6, 32000:] = 100.0                                              # padding rows can win; first one reported
    out = torch.empty(9, dtype=torch.int64, device="cuda")
    xd = x.cuda()
    _lib.check(lib.ovla_argmax(P(xd), C.c_longlong(V), 9, V, P(out), None))
    assert out.cpu().tolist() == torch.argmax(x, dim=-1).tolist()
    assert lib.ovla_argmax(P(xd), C.c_longlong(V), 9, 0, P(out), None) != 0


def test_detokenize_unnormalize_bit_exact_all_ids(L):
    _lib, lib = L
    ids = np.arange(0, 32064, dtype=np.int64)
    ids = np.concatenate([ids, ids[:1]])                               # 32065 = 7 * 4580 + 5 -> pad to a multiple of 7
    ids = np.concatenate([ids, np.full(7 - len(ids) % 7, 31900, dtype=np.int64)])
    stats = O.default_stats()
    ref = O.unnormalize(O.detokenize(ids.reshape(-1, 7)), stats)
    _, centers = O.action_bins(256)
    q01, q99 = np.asarray(stats["q01"]), np.asarray(stats["q99"])
    mask = np.asarray(stats["mask"], dtype=np.uint8)
    ids_d, c_d, lo_d, hi_d, m_d = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (ids, centers, q01, q99, mask)]
    out = torch.empty(len(ids), dtype=torch.float64, device="cuda")
    _lib.check(lib.ovla_detokenize(P(ids_d), len(ids), 7, 32000, P(c_d), 255, P(lo_d), P(hi_d), P(m_d), P(out), None))
    assert np.array_equal(out.cpu().numpy().reshape(-1, 7), ref)       # float64, bit for bit


@pytest.mark.parametrize("B,T,H,bn,cg", [(2, 19, 2, 0, 0), (3, 288, 4, 256, 2), (1, 130, 2, 128, 1), (2, 70, 2, 256, 1)])
def test_fused_qkv_rope_gemm_equals_unfused_path(L, B, T, H, bn, cg):
    """The fused QKV + RoPE + KV-write epilogue is bit-identical to GEMM followed by the stand-alone RoPE kernel
    (which test_rope_kv_prefill_bit_exact pins bit-exactly to the oracle's bf16 rotary embedding)."""
    _lib, lib = L
    from openvla_probe_b200.engine import rope_tables

    hd, Tmax, K = 128, T + 9, 256
    D = H * hd
    g = torch.Generator().manual_seed(T)
    x = bf(torch.randn(B * T, K, generator=g)).cuda()
    W = bf(torch.randn(3 * D, K, generator=g) * 0.08).cuda()
    cos, sin = rope_tables(hd, 10000.0, Tmax)
    cd, sd = cos.cuda(), sin.cuda()
    # unfused: GEMM -> rope kernel
    ref = torch.empty(B * T, 3 * D, dtype=torch.bfloat16, device="cuda")
    epi = _lib.GemmEpilogue()
    _lib.check(lib.ovla_gemm(P(x), C.c_longlong(K), P(W), C.c_longlong(K), B * T, 3 * D, K, 0, 0, P(ref), C.c_longlong(3 * D),
                             C.byref(epi), bn, cg, None))
    kc0 = torch.zeros(B, H, Tmax, hd, dtype=torch.bfloat16, device="cuda")
    vc0 = torch.zeros_like(kc0)
    _lib.check(lib.ovla_rope_kv(P(ref), B, T, H, hd, 3, P(cd), P(sd), P(kc0), P(vc0), Tmax, None))
    # fused
    out = torch.zeros(B * T, 3 * D, dtype=torch.bfloat16, device="cuda")
    kc1, vc1 = torch.zeros_like(kc0), torch.zeros_like(kc0)
    _lib.check(lib.ovla_qkv_rope_gemm(P(x), C.c_longlong(K), P(W), C.c_longlong(K), B * T, H, K, T, 3, P(cd), P(sd), P(out),
                                      C.c_longlong(3 * D), P(kc1), P(vc1), Tmax, bn, cg, None))
    assert torch.equal(out[:, :D], ref[:, :D])
    assert torch.equal(kc1, kc0) and torch.equal(vc1, vc0)
    assert lib.ovla_qkv_rope_gemm(P(x), C.c_longlong(K), P(W), C.c_longlong(K), B * T, H, K, T, 20, P(cd), P(sd), P(out),
                                  C.c_longlong(3 * D), P(kc1), P(vc1), Tmax, bn, cg, None) != 0      # past KV capacity


def test_preprocess_frames_bit_exact_and_native_twin():
    """uint8 frames -> bf16 pixel_values on the device == the host transform (to_tensor, per-tower normalize, bf16 cast),
    bit for bit; the native-style OpenVLA.predict_action(image, instruction) gives the same action as the HF surface."""
    import dataclasses

    from openvla_probe_b200 import config as cfgmod
    from openvla_probe_b200.modeling_prismatic import from_state_dict
    from openvla_probe_b200.vlas import OpenVLA, PurePromptBuilder, hash_tokenizer

    od = O.tiny_dims()
    cfg = dataclasses.replace(cfgmod.tiny(), norm_stats={"synthetic": {"action": O.default_stats()}})
    model = from_state_dict(cfg, O.make_weights(od, seed=0), max_batch=2, max_prompt_len=40)
    rng = np.random.default_rng(1)                      # same frames as O.make_inputs(seed=1)
    img = rng.integers(0, 256, (2, od.image_size, od.image_size, 3), dtype=np.uint8)
    _, px_host = O.make_inputs(od, 2, prompt_len=9, seed=1)
    px_dev = model.preprocess_frames(torch.from_numpy(img))
    assert torch.equal(px_dev.cpu(), px_host)
    pb = PurePromptBuilder()
    pb.add_turn("human", "What action should the robot take to pick up the cup?")
    assert pb.get_prompt() == "In: What action should the robot take to pick up the cup?\nOut:"
    vla = OpenVLA(model)
    a_native = vla.predict_action(img[0], "Pick up the cup", unnorm_key="synthetic")
    ids = torch.tensor([hash_tokenizer(pb.get_prompt())])
    a_hf = model.predict_action(ids, unnorm_key="synthetic", pixel_values=px_host[:1])
    assert a_native.shape == (7,) and np.array_equal(a_native, a_hf)


@pytest.mark.parametrize("M,N,K,mode,expect_split", [(288, 4096, 11008, 0, True), (256, 4096, 11008, 0, True),
                                                     (261, 1024, 4096, 0, True), (128, 2048, 4096, 1, True),
                                                     (64, 2048, 8192, 2, True), (300, 12288, 4096, 0, False)])
def test_splitk_small_m_gemm(L, M, N, K, mode, expect_split):
    """Small-M shapes go through split-K (partial tiles + ordered reduce/epilogue kernel) under the library heuristic;
    same tolerance as the fused epilogue."""
    _lib, lib = L
    g = torch.Generator().manual_seed(N + K)
    A = bf(torch.randn(M, K, generator=g) * 0.5).cuda()
    W = bf(torch.randn(N, K, generator=g) * 0.03).cuda()
    epi = _lib.GemmEpilogue()
    ref = A.float() @ W.float().t()
    if mode == 0:
        bias, resid = bf(torch.randn(N, generator=g) * 0.1).cuda(), bf(torch.randn(M, N, generator=g)).cuda()
        epi.bias_bf16, epi.resid_bf16, epi.ld_resid = bias.data_ptr(), resid.data_ptr(), N
        ref = (ref + bias.float()).bfloat16().float() + resid.float()
        out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
        n_out = N
    elif mode == 1:
        r3 = ref.view(M, N // 64, 2, 32)
        gte, up = r3[:, :, 0].reshape(M, -1).bfloat16().float(), r3[:, :, 1].reshape(M, -1).bfloat16().float()
        ref = F.silu(gte).bfloat16().float() * up
        out = torch.empty(M, N // 2, dtype=torch.bfloat16, device="cuda")
        n_out = N // 2
    else:
        epi.round_bf16 = 1
        ref = ref.bfloat16().float()
        out = torch.empty(M, N, dtype=torch.float32, device="cuda")
        n_out = N
    before = lib.ovla_launch_count()
    _lib.check(lib.ovla_gemm(P(A), C.c_longlong(K), P(W), C.c_longlong(K), M, N, K, mode, 0, P(out), C.c_longlong(n_out),
                             C.byref(epi), 0, 0, None))
    # 2 launches = partial GEMM + ordered reduce/epilogue kernel (the split-K path); 1 = plain fused GEMM
    assert lib.ovla_launch_count() - before == (2 if expect_split else 1)
    ok, e = close_bf16(out, ref, ulps=3.0)
    assert ok, e


@pytest.mark.parametrize("B,H,W,scale", [(3, 256, 256, 0.9), (1, 224, 224, 0.9), (2, 300, 200, 0.5), (2, 224, 224, 1.0),
                                         (1, 97, 131, 0.9)])
def test_center_crop_frames_bit_exact(L, B, H, W, scale):
    """center_crop=True input branch (openvla_utils.py:155-175): the CUDA kernel and the oracle's float32 restatement of
    tf.image.crop_and_resize + convert_image_dtype agree on every byte."""
    _lib, lib = L
    rng = np.random.default_rng(H * W + B)
    img = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    img[0, : H // 3] = np.linspace(0, 255, W, dtype=np.uint8)[None, :, None]      # smooth ramp: many rounding boundaries
    ref = O.center_crop_frames(img, scale, 224)
    src = torch.from_numpy(img).cuda()
    out = torch.empty(B, 224, 224, 3, dtype=torch.uint8, device="cuda")
    _lib.check(lib.ovla_center_crop_frames(P(src), B, H, W, C.c_float(scale), P(out), 224, None))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), ref)
    if scale == 1.0 and (H, W) == (224, 224):
        assert np.array_equal(ref, img)                                           # full box at native size is the identity



# ----------------------------------------------------------------------------------------------- ragged prompts
def test_pool_tokens_ragged_rows_equal_uniform_calls():
    """ovla_pool_tokens_ragged: row b pools over n_rows - (P - lens[b]) rows, bit-identical to a uniform call of that
    length (pool over each sample's true length)."""
    import ctypes as C

    from openvla_probe_b200 import _lib

    lib = _lib.load()
    torch.manual_seed(0)
    B, T, D, P, n_rows = 5, 40, 512, 9, 38
    x = torch.randn(B, T, D, device="cuda").to(torch.bfloat16)
    lens = torch.tensor([9, 1, 4, 9, 7], dtype=torch.int32, device="cuda")
    for mode in (0, 1):
        out = torch.empty(B, D, device="cuda")
        _lib.check(lib.ovla_pool_tokens_ragged(x.data_ptr(), T * D, D, B, n_rows, D, mode, lens.data_ptr(), P,
                                               out.data_ptr(), D, _lib.stream_ptr()))
        for b in range(B):
            nb = n_rows - (P - int(lens[b]))
            one = torch.empty(1, D, device="cuda")
            _lib.check(lib.ovla_pool_tokens(x[b].data_ptr(), T * D, D, 1, nb, D, mode, one.data_ptr(), D, _lib.stream_ptr()))
            assert torch.equal(out[b], one[0]), (mode, b)
            ref = x[b, :nb].float().mean(0) if mode == 0 else x[b, nb - 1].float()
            assert torch.allclose(out[b], ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B", [3, 200])
def test_decode_attention_ragged_rows_equal_uniform_calls(B):
    """ovla_decode_rope_attention_ragged: row b works at pos - (P - lens[b]); output and the appended K/V rows are
    bit-identical to a batch-1 call at that position (B = 200 takes the low-register looped form of the kernel, B = 3
    the loads-up-front form: same bits)."""
    from openvla_probe_b200 import _lib

    lib = _lib.load()
    torch.manual_seed(1)
    H, hd, Tmax, P, pos = 4, 128, 64, 10, 40
    g = torch.Generator().manual_seed(2)
    lens = torch.randint(1, P + 1, (B,), generator=g, dtype=torch.int32)
    lens[0] = P
    qkv = torch.randn(B, 3 * H * hd, device="cuda").to(torch.bfloat16)
    kc0 = torch.randn(B, H, Tmax, hd, device="cuda").to(torch.bfloat16)
    vc0 = torch.randn(B, H, Tmax, hd, device="cuda").to(torch.bfloat16)
    ang = torch.arange(Tmax).float().view(-1, 1) * (1.0 / (10000 ** (torch.arange(0, hd, 2).float() / hd))).view(1, -1)
    cos, sin = ang.cos().to(torch.bfloat16).cuda(), ang.sin().to(torch.bfloat16).cuda()
    kc, vc = kc0.clone(), vc0.clone()
    out = torch.empty(B, H * hd, device="cuda", dtype=torch.bfloat16)
    lens_d = lens.cuda()
    _lib.check(lib.ovla_decode_rope_attention_ragged(qkv.data_ptr(), 3 * H * hd, cos.data_ptr(), sin.data_ptr(), pos,
                                                     lens_d.data_ptr(), P, kc.data_ptr(), vc.data_ptr(), B, H, hd, Tmax,
                                                     out.data_ptr(), H * hd, _lib.stream_ptr()))
    for b in list(range(min(B, 6))) + ([B - 1] if B > 6 else []):
        pb = pos - (P - int(lens[b]))
        k1, v1 = kc0[b:b + 1].clone(), vc0[b:b + 1].clone()
        o1 = torch.empty(1, H * hd, device="cuda", dtype=torch.bfloat16)
        _lib.check(lib.ovla_decode_rope_attention(qkv[b:b + 1].data_ptr(), 3 * H * hd, cos.data_ptr(), sin.data_ptr(), pb,
                                                  k1.data_ptr(), v1.data_ptr(), 1, H, hd, Tmax, o1.data_ptr(), H * hd,
                                                  _lib.stream_ptr()))
        assert torch.equal(out[b], o1[0]), b
        assert torch.equal(kc[b], k1[0]) and torch.equal(vc[b], v1[0]), b


# ----------------------------------------------------------------------------------------------- RMSNorm fused into the GEMMs
def _slot_sumsq(x32, N):
    """Reference of the partial-sum layout: slot 2 g + p = 128-column group g, 32-column chunks of parity p."""
    M = x32.shape[0]
    sq = (x32.double() ** 2).view(M, N // 128, 2, 2, 32)          # [group, pair, parity, 32]
    return sq.sum(dim=(2, 4)).reshape(M, N // 64)


def _rms_ref(x, gamma, eps):
    """LlamaRMSNorm on bf16 input in exact arithmetic (no intermediate rounding): the fp32-oracle view."""
    x64 = x.double()
    return gamma.double() * (x64 * torch.rsqrt((x64 ** 2).mean(-1, keepdim=True) + eps))


def test_row_sumsq_and_fold_norm_weight(L):
    _lib, lib = L
    g = torch.Generator().manual_seed(5)
    for rows, D in [(7, 256), (300, 4096), (33, 384)]:
        x = bf(torch.randn(rows, D, generator=g) * torch.rand(rows, 1, generator=g) * 30).cuda()
        ss = torch.full((rows, D // 64 + 4), -1.0, device="cuda")
        _lib.check(lib.ovla_row_sumsq(P(x), C.c_longlong(D), rows, D, P(ss), D // 64 + 4, None))
        ref = _slot_sumsq(x.float().cpu(), D)
        assert torch.allclose(ss[:, : D // 64].double().cpu(), ref, rtol=1e-5, atol=0)
        assert bool((ss[:, D // 64:] == -1).all())
    W = bf(torch.randn(704, 256, generator=g) * 0.05).cuda()
    gam = bf(1 + 0.2 * torch.randn(256, generator=g)).cuda()
    out = torch.empty_like(W)
    _lib.check(lib.ovla_fold_norm_weight(P(W), P(gam), P(out), C.c_longlong(704), 256, None))
    assert torch.equal(out, (W.float() * gam.float()).bfloat16())
    assert lib.ovla_row_sumsq(P(W), C.c_longlong(200), 4, 200, P(out), 8, None) != 0       # D % 128 != 0


@pytest.mark.parametrize("M,N,K,bn,cg", [(1100, 4096, 4096, 0, 0), (1100, 4096, 1024, 128, 1), (300, 256, 192, 0, 0),
                                         (700, 1024, 512, 256, 1), (2500, 512, 256, 128, 2), (1100, 4096, 4096, -1, 0)])
def test_gemm_residual_epilogue_writes_row_sums_of_squares(L, M, N, K, bn, cg):
    """Producer half of the fused RMSNorm: the in-place residual GEMM (o_proj / down_proj) also leaves, per row, the
    partial sums of squares of the bf16 values it stores -- equal (fp32 round-off) to the sums recomputed from its own
    output, in the fixed slot layout, for every tile shape; the output itself is bit-identical to the plain call.
    bn = -1: the direct-store epilogue (OVLA_GEMM_TMA_EPI=0 is read once per process, so it runs in a child)."""
    _lib, lib = L
    if bn == -1:
        import subprocess, sys, os
        code = ("import sys; sys.path.insert(0, 'tests'); sys.path.insert(0, '.'); import test_gpu_operators as t;"
                "from openvla_probe_b200 import _lib;"
                "t._sumsq_producer_case((_lib, _lib.load()), %d, %d, %d, 0, 0)" % (M, N, K))
        env = dict(os.environ, OVLA_GEMM_TMA_EPI="0")
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True,
                           cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        assert r.returncode == 0, r.stderr[-2000:]
        return
    _sumsq_producer_case(L, M, N, K, bn, cg)


def _sumsq_producer_case(L, M, N, K, bn, cg):
    _lib, lib = L
    g = torch.Generator().manual_seed(M + N + K)
    A = bf(torch.randn(M, K, generator=g) * 0.5).cuda()
    W = bf(torch.randn(N, K, generator=g) * 0.05).cuda()
    x0 = bf(torch.randn(M, N, generator=g) * (1 + 20 * torch.rand(M, 1, generator=g))).cuda()
    outs = []
    for with_ss in (0, 1):
        x = x0.clone()
        ss = torch.full((M, N // 64), float("nan"), device="cuda")
        epi = _lib.GemmEpilogue()
        epi.resid_bf16, epi.ld_resid = x.data_ptr(), N
        if with_ss:
            epi.row_sumsq_out, epi.row_sumsq_ld = ss.data_ptr(), N // 64
        _lib.check(lib.ovla_gemm(P(A), C.c_longlong(K), P(W), C.c_longlong(K), M, N, K, 0, 0, P(x), C.c_longlong(N),
                                 C.byref(epi), bn, cg, None))
        outs.append(x)
    assert torch.equal(outs[0], outs[1])
    ref = _slot_sumsq(outs[1].float().cpu(), N)
    got = ss.double().cpu()
    assert bool(torch.isfinite(got).all())
    assert torch.allclose(got, ref, rtol=2e-5, atol=0), float(((got - ref).abs() / ref).max())
    for _ in range(2):                                     # deterministic: plain stores, no atomics
        x = x0.clone()
        ss2 = torch.zeros_like(ss)
        epi.resid_bf16, epi.row_sumsq_out = x.data_ptr(), ss2.data_ptr()
        _lib.check(lib.ovla_gemm(P(A), C.c_longlong(K), P(W), C.c_longlong(K), M, N, K, 0, 0, P(x), C.c_longlong(N),
                                 C.byref(epi), bn, cg, None))
        assert torch.equal(ss2, ss)


@pytest.mark.parametrize("M,K,I,bn,cg", [(600, 256, 704, 0, 0), (1100, 4096, 1408, 0, 0), (300, 512, 256, 128, 1),
                                         (520, 1024, 512, 256, 2)])
def test_fused_rmsnorm_swiglu_gemm(L, M, K, I, bn, cg):
    """Consumer half: gate/up GEMM on the UN-normalised rows with the norm weight folded into W and 1/rms applied to
    the fp32 accumulators (row sums of squares from ovla_row_sumsq) against exact arithmetic; the fused result may be
    at most as far from it as 1.5x the un-fused device path (ovla_rmsnorm, then the plain SwiGLU GEMM) + 1 bf16 ulp of
    the row scale -- it skips two per-element roundings and adds one rounding of the weights."""
    _lib, lib = L
    eps = 1e-5
    g = torch.Generator().manual_seed(M + K)
    x = bf(torch.randn(M, K, generator=g) * (0.2 + 30 * torch.rand(M, 1, generator=g)))
    gam = bf(1 + 0.3 * torch.randn(K, generator=g))
    Wg, Wu = bf(torch.randn(I, K, generator=g) * 0.06), bf(torch.randn(I, K, generator=g) * 0.06)
    h = _rms_ref(x, gam, eps)
    ref = (F.silu(h @ Wg.double().t()) * (h @ Wu.double().t())).float()
    Wi = torch.stack([Wg.view(I // 32, 32, K), Wu.view(I // 32, 32, K)], 1).reshape(2 * I, K).contiguous().cuda()
    xd, gd = x.cuda(), gam.cuda()
    # un-fused device path
    hd_ = torch.empty_like(xd)
    _lib.check(lib.ovla_rmsnorm(P(xd), C.c_longlong(K), P(gd), C.c_float(eps), P(hd_), C.c_longlong(K), M, K, None))
    plain = torch.empty(M, I, dtype=torch.bfloat16, device="cuda")
    epi = _lib.GemmEpilogue()
    _lib.check(lib.ovla_gemm(P(hd_), C.c_longlong(K), P(Wi), C.c_longlong(K), M, 2 * I, K, 1, 0, P(plain), C.c_longlong(I),
                             C.byref(epi), bn, cg, None))
    # fused
    ss = torch.empty(M, K // 64, device="cuda")
    _lib.check(lib.ovla_row_sumsq(P(xd), C.c_longlong(K), M, K, P(ss), K // 64, None))
    Wn = torch.empty_like(Wi)
    _lib.check(lib.ovla_fold_norm_weight(P(Wi), P(gd), P(Wn), C.c_longlong(2 * I), K, None))
    fused = torch.empty(M, I, dtype=torch.bfloat16, device="cuda")
    epi = _lib.GemmEpilogue()
    epi.row_sumsq_in, epi.row_sumsq_ld, epi.row_sumsq_parts, epi.norm_eps = ss.data_ptr(), K // 64, K // 64, eps
    _lib.check(lib.ovla_gemm(P(xd), C.c_longlong(K), P(Wn), C.c_longlong(K), M, 2 * I, K, 1, 0, P(fused), C.c_longlong(I),
                             C.byref(epi), bn, cg, None))
    scale = ref.abs().amax(-1, keepdim=True).clamp_min(1e-6)
    e_f = float(((fused.float().cpu() - ref).abs() / scale).max())
    e_p = float(((plain.float().cpu() - ref).abs() / scale).max())
    assert e_f <= 1.5 * e_p + 2.0 ** -8, (e_f, e_p)
    rl_f = float((fused.float().cpu() - ref).norm() / ref.norm())
    rl_p = float((plain.float().cpu() - ref).norm() / ref.norm())
    assert rl_f <= 1.25 * rl_p + 1e-4, (rl_f, rl_p)
    # a norm input is refused by the epilogues that do not implement it
    bad = _lib.GemmEpilogue()
    bad.row_sumsq_in, bad.row_sumsq_ld, bad.row_sumsq_parts = ss.data_ptr(), K // 64, K // 64
    assert lib.ovla_gemm(P(xd), C.c_longlong(K), P(Wn), C.c_longlong(K), M, 2 * I, K, 0, 0, P(Wn), C.c_longlong(2 * I),
                         C.byref(bad), bn, cg, None) != 0


@pytest.mark.parametrize("B,T,H,bn,cg", [(3, 200, 2, 0, 0), (4, 288, 4, 256, 2), (2, 130, 2, 128, 1)])
def test_fused_rmsnorm_qkv_rope_gemm(L, B, T, H, bn, cg):
    """input_layernorm fused into the QKV + RoPE + KV-write GEMM against the un-fused device path (ovla_rmsnorm, then
    ovla_qkv_rope_gemm) and exact arithmetic: rotated q, and the k / v rows written to the cache."""
    _lib, lib = L
    from openvla_probe_b200.engine import rope_tables

    hd, Tmax, K, eps = 128, T + 5, 512, 1e-5
    D, M = H * hd, B * T
    g = torch.Generator().manual_seed(T + H)
    x = bf(torch.randn(M, K, generator=g) * (0.2 + 30 * torch.rand(M, 1, generator=g))).cuda()
    gam = bf(1 + 0.3 * torch.randn(K, generator=g)).cuda()
    W = bf(torch.randn(3 * D, K, generator=g) * 0.08).cuda()
    cos, sin = rope_tables(hd, 10000.0, Tmax)
    cd, sd = cos.cuda(), sin.cuda()
    h = torch.empty_like(x)
    _lib.check(lib.ovla_rmsnorm(P(x), C.c_longlong(K), P(gam), C.c_float(eps), P(h), C.c_longlong(K), M, K, None))
    q0 = torch.zeros(M, 3 * D, dtype=torch.bfloat16, device="cuda")
    kc0 = torch.zeros(B, H, Tmax, hd, dtype=torch.bfloat16, device="cuda")
    vc0 = torch.zeros_like(kc0)
    _lib.check(lib.ovla_qkv_rope_gemm(P(h), C.c_longlong(K), P(W), C.c_longlong(K), M, H, K, T, 0, P(cd), P(sd), P(q0),
                                      C.c_longlong(3 * D), P(kc0), P(vc0), Tmax, bn, cg, None))
    ss = torch.empty(M, K // 64, device="cuda")
    _lib.check(lib.ovla_row_sumsq(P(x), C.c_longlong(K), M, K, P(ss), K // 64, None))
    Wn = torch.empty_like(W)
    _lib.check(lib.ovla_fold_norm_weight(P(W), P(gam), P(Wn), C.c_longlong(3 * D), K, None))
    q1 = torch.zeros_like(q0)
    kc1, vc1 = torch.zeros_like(kc0), torch.zeros_like(kc0)
    _lib.check(lib.ovla_qkv_rope_gemm_rownorm(P(x), C.c_longlong(K), P(Wn), C.c_longlong(K), M, H, K, T, 0, P(cd), P(sd),
                                              P(q1), C.c_longlong(3 * D), P(kc1), P(vc1), Tmax, P(ss), K // 64, K // 64,
                                              C.c_float(eps), bn, cg, None))
    # exact reference: v needs no rotation; q, k through the bf16 cos / sin tables in fp64
    hx = _rms_ref(x.cpu(), gam.cpu(), eps)
    qkv = hx @ W.double().cpu().t()
    pos = torch.arange(T).repeat(B)
    c64, s64 = cos.double()[pos], sin.double()[pos]                      # [M, 64]

    def rot(t):                                                          # [M, H, 128]
        a, b = t[..., :64], t[..., 64:]
        return torch.cat([a * c64[:, None] - b * s64[:, None], b * c64[:, None] + a * s64[:, None]], -1)

    q_ref = rot(qkv[:, :D].view(M, H, hd)).reshape(M, D)
    k_ref = rot(qkv[:, D:2 * D].view(M, H, hd)).view(B, T, H, hd).permute(0, 2, 1, 3)
    v_ref = qkv[:, 2 * D:].view(B, T, H, hd).permute(0, 2, 1, 3)
    for name, f, p, r in [("q", q1[:, :D], q0[:, :D], q_ref), ("k", kc1[:, :, :T], kc0[:, :, :T], k_ref),
                          ("v", vc1[:, :, :T], vc0[:, :, :T], v_ref)]:
        f, p, r = f.double().cpu(), p.double().cpu(), r
        rl_f, rl_p = float((f - r).norm() / r.norm()), float((p - r).norm() / r.norm())
        assert rl_f <= 1.25 * rl_p + 1e-4, (name, rl_f, rl_p)
        scale = r.abs().amax(-1, keepdim=True).clamp_min(1e-6)
        assert float(((f - r).abs() / scale).max()) <= 1.5 * float(((p - r).abs() / scale).max()) + 2.0 ** -8, name
    assert bool((kc1[:, :, T:] == 0).all()) and bool((vc1[:, :, T:] == 0).all())
