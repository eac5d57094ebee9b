"""fp32 mode on the tensor cores (csrc/tc_f32.cu) -- split-precision tcgen05 GEMMs + streaming fp32 kernels -- against the
numpy oracle, against the CUDA-core fp32 kernels, and the split-precision GEMM alone against fp64 (B200, `-m gpu`).

Gates (BASELINE.json north_star, fp32 mode): logits / loss within 1e-5 relative; gradients within 1e-2 relative (held to
1e-3 here).  The bmm's this path replaces are gloria_loss.py:40,59 and their autograd.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import gloria_oracle as O
from oracle.make_golden import gen_inputs
from tests.util import cu, relerr

pytestmark = pytest.mark.gpu

LOGIT_TOL = 1e-5


@pytest.fixture(scope="module")
def gl():
    import gloria_nlp_project_b200 as g
    from gloria_nlp_project_b200 import gloria_loss
    g.set_precision("fp32")
    return gloria_loss


def _split3(x):
    p0 = x.to(torch.bfloat16)
    r = x - p0.float()
    p1 = r.to(torch.bfloat16)
    p2 = (r - p1.float()).to(torch.bfloat16)
    return torch.stack((p0, p1, p2)).contiguous()


def _planes_gemm(A, B, a_kmajor, nterms, ksplit=1):
    """A [nb, M, K] fp32, B [nb, K, N] fp32 -> C [nb, M, N] through gloria_b200_acc_gemm_planes."""
    from gloria_nlp_project_b200 import _lib
    lib = _lib.lib()
    nb, M, K = A.shape
    N = B.shape[2]
    Ap = _split3(A if a_kmajor else A.transpose(1, 2).contiguous())       # [3, nb, M, K] or [3, nb, K, M]
    Bp = _split3(B)
    out = torch.full((nb, M, N), float("nan"), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.gloria_b200_acc_gemm_planes(Ap.data_ptr(), Bp.data_ptr(), out.data_ptr(), M, N, K, 1 if a_kmajor else 0, nterms, nb,
                                         ksplit, 0, C.c_void_p(st))
    _lib.check(rc, "acc_gemm_planes")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("nb,M,N,K,a_kmajor", [(1, 512, 256, 768, True), (1, 1000, 384, 768, True), (3, 384, 768, 1088, True),
                                               (5, 320, 768, 384, False), (1, 640, 768, 2048, False), (2, 1088, 384, 768, True)])
def test_split_precision_gemm(nb, M, N, K, a_kmajor):
    """Six piece products reproduce an fp32 GEMM to fp32 accuracy (rms ~1e-6 of the mean magnitude); three to ~1e-5.
    Batches are stacked along the rows of the planes; rows past a batch's M and tiles past N are clipped."""
    g = torch.Generator(device="cuda").manual_seed(nb * 1000 + M + N + K)
    A = torch.randn((nb, M, K), device="cuda", generator=g)
    B = torch.randn((nb, K, N), device="cuda", generator=g)
    ref = A.double() @ B.double()
    scale = float(ref.abs().mean())
    for nterms, tol in ((6, 1.2e-5), (3, 1e-4)):
        out = _planes_gemm(A, B, a_kmajor, nterms)
        assert torch.isfinite(out).all()
        err = float((out.double() - ref).abs().max()) / scale
        print(f"nb={nb} {M}x{N}x{K} kmajor={a_kmajor} terms={nterms}: max err / mean|C| = {err:.2e}")
        assert err < tol * max(1.0, (K / 768) ** 0.5)


def test_split_precision_gemm_ksplit():
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn((1, 256, 8192), device="cuda", generator=g)
    B = torch.randn((1, 8192, 256), device="cuda", generator=g)
    ref = A.double() @ B.double()
    out = _planes_gemm(A, B, False, 3, ksplit=4)
    assert float((out.double() - ref).abs().max()) / float(ref.abs().mean()) < 3e-4


def _run(gl, img_l, txt_l, cl, **kw):
    img, txt = cu(img_l, True), cu(txt_l, True)
    l0, l1, _, _, _, maps = gl.local_loss(img, txt, cl, **kw)
    (l0 + l1).backward()
    sim, _, _, _ = gl.local_similarities(img.detach(), txt.detach(), cl)
    return float(l0), float(l1), sim, maps, img.grad, txt.grad


@pytest.mark.parametrize("B,seed,scale,budget", [(16, 3, 1.0, None), (16, 4, 0.05, 700 << 20), (48, 5, 1.0, None), (7, 6, 1.0, None)])
def test_vs_oracle_and_cuda_core_path(gl, B, seed, scale, budget, monkeypatch):
    """configs[0] (B = 16) and configs[1] (B = 48) shapes, ragged captions: the tensor-core fp32 path against the numpy
    oracle (fp64) and against the CUDA-core fp32 kernels; one case runs caption-chunked through a small workspace."""
    from gloria_nlp_project_b200 import ops, _lib
    assert _lib.lib().gloria_b200_f32tc_supported(768, 361, 97) == 0
    if budget is not None:
        monkeypatch.setattr(ops, "_F32_TC_WS_BUDGET", budget)
    img_l, txt_l, _, _, cl = gen_inputs(seed, B, 768, 19, 19, 97, scale=scale, dtype=np.float32)
    monkeypatch.setattr(ops, "_F32_TC", True)
    n0 = _lib.lib().gloria_b200_launch_count(1)
    l0, l1, sim, maps, d_img, d_txt = _run(gl, img_l, txt_l, cl)
    monkeypatch.setattr(ops, "_F32_TC", False)
    s0, s1, sim_s, maps_s, d_img_s, d_txt_s = _run(gl, img_l, txt_l, cl)
    o0, o1, _, _, _, omaps, ologits = O.local_loss(img_l.astype(np.float64), txt_l.astype(np.float64), cl)
    assert abs(l0 - o0) < LOGIT_TOL * abs(o0) and abs(l1 - o1) < LOGIT_TOL * abs(o1)
    e_tc, e_simt = relerr(sim * 10.0, ologits), relerr(sim_s * 10.0, ologits)
    print(f"B={B} scale={scale}: logits rel err tensor-core {e_tc:.2e}, CUDA-core {e_simt:.2e}")
    assert e_tc < LOGIT_TOL
    for i in (0, B // 2, B - 1):
        assert relerr(maps[i], omaps[i]) < (1e-4 if scale == 1.0 else 1e-5)
    od_img, od_txt = O.local_loss_bwd(img_l.astype(np.float64), txt_l.astype(np.float64), cl)
    g_tc, g_simt = max(relerr(d_img, od_img), relerr(d_txt, od_txt)), max(relerr(d_img_s, od_img), relerr(d_txt_s, od_txt))
    print(f"B={B} scale={scale}: gradient rel err tensor-core {g_tc:.2e}, CUDA-core {g_simt:.2e}")
    assert g_tc < 1e-3
    # padded word columns of d_words are exactly zero, as in the CUDA-core path
    for i, L in enumerate(cl):
        assert torch.all(d_txt[i, :, L:] == 0)


def test_regulariser_outputs_and_gradients(gl, monkeypatch):
    """word-mean attention + diagonal maps and the gradient through them (entropy / KL / no-attn regularisers,
    gloria_loss.py:108-139): tensor-core fp32 path against the oracle's closed form."""
    from gloria_nlp_project_b200 import ops
    monkeypatch.setattr(ops, "_F32_TC", True)
    kw = dict(attention_divergence_loss_weight=0.2, attention_entropy_loss_weight=0.1)
    img_l, txt_l, _, _, cl = gen_inputs(11, 6, 768, 19, 19, 97, scale=0.05, dtype=np.float32)
    img, txt = cu(img_l, True), cu(txt_l, True)
    out = gl.local_loss(img, txt, cl, **kw)
    (out[0] + out[1] + out[3] + out[4]).backward()
    monkeypatch.setattr(ops, "_F32_TC", False)
    img2, txt2 = cu(img_l, True), cu(txt_l, True)
    out2 = gl.local_loss(img2, txt2, cl, **kw)
    (out2[0] + out2[1] + out2[3] + out2[4]).backward()
    for a, b in zip(out[:5], out2[:5]):
        if a is not None and b is not None:
            assert abs(float(a) - float(b)) <= 2e-5 * max(1.0, abs(float(b)))
    assert relerr(img.grad, img2.grad) < 1e-4 and relerr(txt.grad, txt2.grad) < 1e-4


def test_word_window_and_max(gl, monkeypatch):
    """get_local_similarities' call shape: word offset 1, agg = max, fewer texts than images (gloria_model.py:171-207)."""
    from gloria_nlp_project_b200 import ops
    monkeypatch.setattr(ops, "_F32_TC", True)
    img_l, txt_l, _, _, cl = gen_inputs(13, 9, 768, 19, 19, 30, dtype=np.float32)
    lens = [max(1, c - 2) for c in cl[:4]]
    sim, _, _, _ = gl.local_similarities(cu(img_l), cu(txt_l[:4]), lens, 4.0, 5.0, "max", word_offset=1)
    ref = O.local_similarities(img_l.astype(np.float64), txt_l[:4].astype(np.float64), lens, 4.0, 5.0, "max", word_offset=1)
    assert relerr(sim, ref) < LOGIT_TOL


def test_image_blocks_forward_only(gl, monkeypatch):
    """Zero-shot call shape (many images, few short prompts, no gradient): the images are cut into blocks that fit the
    workspace budget; the result does not depend on the cut."""
    from gloria_nlp_project_b200 import ops
    monkeypatch.setattr(ops, "_F32_TC", True)
    img_l, txt_l, _, _, _ = gen_inputs(17, 300, 768, 19, 19, 18, dtype=np.float32)
    lens = [4, 16, 9, 12, 7]
    img, txt = cu(img_l), cu(txt_l[:5])
    with torch.no_grad():
        whole, _, _, _ = gl.local_similarities(img, txt, lens, 4.0, 5.0, "max", word_offset=1)
        monkeypatch.setattr(ops, "_F32_TC_WS_BUDGET", 160 << 20)
        cut, _, mean, _ = gl.local_similarities(img, txt, lens, 4.0, 5.0, "max", word_offset=1, want_mean_attn=True)
    assert torch.equal(whole, cut)
    assert torch.allclose(mean.sum(-1), torch.ones_like(mean[..., 0]), atol=1e-5)
    ref = O.local_similarities(img_l[:40].astype(np.float64), txt_l[:5].astype(np.float64), lens, 4.0, 5.0, "max", word_offset=1)
    assert relerr(cut[:40], ref) < LOGIT_TOL


def test_kept_state_can_be_differentiated_twice(gl, monkeypatch):
    """The training forward keeps its state for the backward, which only reads it: retain_graph works, and the gradients
    equal those of the recompute backward."""
    from gloria_nlp_project_b200 import ops
    monkeypatch.setattr(ops, "_F32_TC", True)
    img_l, txt_l, _, _, cl = gen_inputs(19, 5, 768, 19, 19, 40, dtype=np.float32)
    img, txt = cu(img_l, True), cu(txt_l, True)
    l0, l1, *_ = gl.local_loss(img, txt, cl)
    (l0 + l1).backward(retain_graph=True)
    g1i, g1t = img.grad.clone(), txt.grad.clone()
    img.grad = txt.grad = None
    (l0 + l1).backward()
    assert torch.equal(g1i, img.grad) and torch.equal(g1t, txt.grad)
    monkeypatch.setattr(ops, "_F32_TC_WS_BUDGET", 64 << 20)          # too small to keep a state: recompute backward
    img2, txt2 = cu(img_l, True), cu(txt_l, True)
    m0, m1, *_ = gl.local_loss(img2, txt2, cl)
    (m0 + m1).backward()
    assert relerr(img2.grad, g1i) < 1e-5 and relerr(txt2.grad, g1t) < 1e-5
