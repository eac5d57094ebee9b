"""CTA-pair tcgen05 accumulation GEMM (csrc/tc_gemm.cu) through the C ABI against a plain fp32 torch.matmul of the
same bf16 operands (B200, `-m gpu`).

The kernel replaces the library GEMMs of the training backward (the sums autograd forms for the two bmm's of
gloria_loss.py:40,59); here it is checked alone, in every operand mode the backward uses:
  * row-major A from shared memory, transposed A from shared memory (caption side: A^T = X^T read M-major),
  * row-major A scaled in flight by g[m / m_div, k / k_div] and passed through tensor memory (image side),
  * transposed A scaled in flight in place in shared memory (caption side),
  * k-splits (red.global.add epilogue) and accumulation into an existing C,
  * ragged sizes: M, N, K that are not multiples of the 256 x 256 x 64 tile.
"""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(A, B, M, N, K, a_kmajor=1, ksplit=0, acc_into=None, g=None, g_sm=0, g_sk=0, m_div=1, k_div=1, force=0):
    from gloria_nlp_project_b200 import _lib
    lib = _lib.lib()
    out = acc_into if acc_into is not None else torch.full((M, N), float("nan"), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.gloria_b200_acc_gemm(A.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K, a_kmajor, ksplit,
                                  1 if acc_into is not None else 0, g.data_ptr() if g is not None else None,
                                  g_sm, g_sk, m_div, k_div, force, C.c_void_p(st))
    _lib.check(rc, "acc_gemm")
    torch.cuda.synchronize()
    return out


def _tol(K):
    return 4e-7 * max(K, 1024) ** 0.5 * 2      # fp32 accumulation in a different order than the checker's


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (256, 256, 512), (512, 768, 1024), (300, 128, 200), (1000, 384, 1112),
                                   (2048, 768, 4160)])
@pytest.mark.parametrize("mode", ["ss_ak", "ss_am", "ts", "sc"])
def test_plain(M, N, K, mode):
    gen = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    if mode in ("ss_am", "sc"):
        M = (M + 7) // 8 * 8
    A = torch.randn((M, K), device="cuda", generator=gen).to(torch.bfloat16)
    B = torch.randn((K, N), device="cuda", generator=gen).to(torch.bfloat16)
    ref = A.float() @ B.float()
    if mode in ("ss_am", "sc"):
        out = _gemm(A.t().contiguous(), B, M, N, K, a_kmajor=0, ksplit=1, force=1 if mode == "sc" else 0)
    else:
        out = _gemm(A, B, M, N, K, ksplit=1, force=1 if mode == "ts" else 0)
    assert torch.isfinite(out).all()
    e = _rel(out, ref)
    print(f"{mode} {M}x{N}x{K}: {e:.2e}")
    assert e < _tol(K)


def _weights(g, M, K, m_div, k_div, m_axis):
    """w[m, k] = g[m // m_div, k // k_div] (m_axis == 0) or g[k // k_div, m // m_div] (m_axis == 1)."""
    if m_axis == 0:
        return g.repeat_interleave(m_div, 0)[:M].repeat_interleave(k_div, 1)[:, :K]
    return g.repeat_interleave(k_div, 0)[:K].repeat_interleave(m_div, 1)[:, :M].t()


@pytest.mark.parametrize("M,N,K,m_div,k_div", [(512, 256, 512, 128, 64), (736, 768, 1040, 368, 104), (300, 128, 200, 50, 8),
                                               (1472, 768, 3328, 368, 104), (256 * 40 + 8, 768, 64 * 9, 368, 104)])
def test_scaled_through_tmem(M, N, K, m_div, k_div):
    """Image-side form: A = X^T rows [(j,s), (i,l)], weight g[j, i] = g[m // sp, k // lp]."""
    gen = torch.Generator(device="cuda").manual_seed(11 + M + K)
    A = torch.randn((M, K), device="cuda", generator=gen).to(torch.bfloat16)
    B = torch.randn((K, N), device="cuda", generator=gen).to(torch.bfloat16)
    nj, ni = (M + m_div - 1) // m_div, (K + k_div - 1) // k_div
    g = torch.randn((nj, ni), device="cuda", generator=gen)
    As = A * _weights(g, M, K, m_div, k_div, 0).to(torch.bfloat16)             # bf16 x bf16(g), one rounding: as the kernel
    ref = As.float() @ B.float()
    out = _gemm(A, B, M, N, K, ksplit=1, g=g, g_sm=ni, g_sk=1, m_div=m_div, k_div=k_div)
    e = _rel(out, ref)
    print(f"scaled (TMEM) {M}x{N}x{K}: {e:.2e}")
    assert e < _tol(K)


@pytest.mark.parametrize("M,N,K,m_div,k_div", [(512, 256, 512, 64, 128), (1040, 768, 736, 104, 368), (200, 128, 304, 8, 50),
                                               (3328, 768, 1472, 104, 368), (64 * 9 * 8, 768, 256 * 10 + 8, 104, 368)])
def test_scaled_in_place_transposed(M, N, K, m_div, k_div):
    """Caption-side form: A^T = X^T [(j,s), (i,l)] is what lies in memory, weight g[j, i] = g[k // sp, m // lp]."""
    gen = torch.Generator(device="cuda").manual_seed(13 + M + K)
    A = torch.randn((M, K), device="cuda", generator=gen).to(torch.bfloat16)
    B = torch.randn((K, N), device="cuda", generator=gen).to(torch.bfloat16)
    nj, ni = (K + k_div - 1) // k_div, (M + m_div - 1) // m_div
    g = torch.randn((nj, ni), device="cuda", generator=gen)
    As = A * _weights(g, M, K, m_div, k_div, 1).to(torch.bfloat16)
    ref = As.float() @ B.float()
    out = _gemm(A.t().contiguous(), B, M, N, K, a_kmajor=0, ksplit=1, g=g, g_sm=1, g_sk=ni, m_div=m_div, k_div=k_div)
    e = _rel(out, ref)
    print(f"scaled (in place) {M}x{N}x{K}: {e:.2e}")
    assert e < _tol(K)


@pytest.mark.parametrize("M,N,K", [(256 * 90, 256, 192), (256 * 40 + 8, 768, 64 * 9), (256 * 3, 768, 64 * 1200)])
@pytest.mark.parametrize("mode", ["ss_ak", "ss_am", "ts", "sc"])
def test_many_units_per_cluster(M, N, K, mode):
    """More work units than CTA pairs (every pipeline runs across unit boundaries), and a long K."""
    gen = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn((M, K), device="cuda", generator=gen).to(torch.bfloat16)
    B = torch.randn((K, N), device="cuda", generator=gen).to(torch.bfloat16)
    ref = A.float() @ B.float()
    if mode in ("ss_am", "sc"):
        out = _gemm(A.t().contiguous(), B, M, N, K, a_kmajor=0, ksplit=1, force=1 if mode == "sc" else 0)
    else:
        out = _gemm(A, B, M, N, K, ksplit=1, force=1 if mode == "ts" else 0)
    e = _rel(out, ref)
    print(f"many units {mode} {M}x{N}x{K}: {e:.2e}")
    assert e < _tol(K)


def test_repeatable_bitwise():
    """Static schedule, fixed accumulation order: two launches of the scaled GEMMs give bit-identical results."""
    M, N, K = 256 * 80, 768, 64 * 40
    gen = torch.Generator(device="cuda").manual_seed(3)
    A = torch.randn((M, K), device="cuda", generator=gen).to(torch.bfloat16)
    At = A.t().contiguous()
    B = torch.randn((K, N), device="cuda", generator=gen).to(torch.bfloat16)
    g = torch.randn(((M + 367) // 368, (K + 103) // 104), device="cuda", generator=gen)
    ni = g.shape[1]
    r0 = _gemm(A, B, M, N, K, ksplit=1, g=g, g_sm=ni, g_sk=1, m_div=368, k_div=104)
    gt = torch.randn(((K + 367) // 368, (M + 103) // 104), device="cuda", generator=gen)
    t0 = _gemm(At, B, M, N, K, a_kmajor=0, ksplit=1, g=gt, g_sm=1, g_sk=gt.shape[1], m_div=104, k_div=368)
    for _ in range(3):
        assert torch.equal(r0, _gemm(A, B, M, N, K, ksplit=1, g=g, g_sm=ni, g_sk=1, m_div=368, k_div=104))
        assert torch.equal(t0, _gemm(At, B, M, N, K, a_kmajor=0, ksplit=1, g=gt, g_sm=1, g_sk=gt.shape[1], m_div=104, k_div=368))


@pytest.mark.parametrize("mode", ["ss_ak", "ss_am"])
def test_ksplit_and_accumulate(mode):
    M, N, K = 520, 768, 64 * 300
    gen = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn((M, K), device="cuda", generator=gen).to(torch.bfloat16)
    B = torch.randn((K, N), device="cuda", generator=gen).to(torch.bfloat16)
    ref = A.float() @ B.float()
    Ain = A if mode == "ss_ak" else A.t().contiguous()
    ak = 1 if mode == "ss_ak" else 0
    out = _gemm(Ain, B, M, N, K, a_kmajor=ak, ksplit=3)
    assert _rel(out, ref) < _tol(K)
    out0 = _gemm(Ain, B, M, N, K, a_kmajor=ak, ksplit=0)
    assert _rel(out0, ref) < _tol(K)
    base = torch.randn((M, N), device="cuda", generator=gen)
    out2 = _gemm(Ain, B, M, N, K, a_kmajor=ak, ksplit=2, acc_into=base.clone())
    assert _rel(out2, ref + base) < _tol(K)
