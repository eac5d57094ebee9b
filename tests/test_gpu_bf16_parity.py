"""bf16 tensor-core mode (tcgen05 / TMEM / TMA) vs the numpy oracle (B200, `-m gpu`).

Gates (BASELINE.json north_star): logits and loss within 2e-3 relative, gradients within 1e-2 relative (max-norm).
"""
import numpy as np
import pytest
import torch

from oracle import gloria_oracle as O
from oracle.make_golden import gen_inputs
from tests.util import cu, relerr

pytestmark = pytest.mark.gpu

LOGIT_TOL = 2e-3
GRAD_TOL = 1e-2


@pytest.fixture()
def gl():
    import gloria_nlp_project_b200 as g
    from gloria_nlp_project_b200 import gloria_loss
    g.set_precision("bf16")
    yield gloria_loss
    g.set_precision("auto")


def test_prepack_layouts(gl):
    """Cast/transposes into the TMA-legal bf16 layouts are exact re-arrangements (bit-exact vs torch casts)."""
    from gloria_nlp_project_b200 import ops
    img_l, txt_l, _, _, cl = gen_inputs(21, 3, 768, 19, 19, 97, cap_lens=[97, 40, 3], dtype=np.float32)
    ctx = cu(img_l).reshape(3, 768, 361)
    words = cu(txt_l)
    lens = torch.tensor(cl, dtype=torch.int32, device="cuda")
    ctx_t, ctx_n, words_t, wnorm = ops.tc_prepack(ctx, words, lens, 97, 0)
    assert ctx_t.shape == (3, 384, 768) and ctx_n.shape == (3, 768, 384) and words_t.shape == (3, 112, 768)
    ref = ctx.to(torch.bfloat16)
    assert torch.equal(ctx_n[:, :, :361], ref) and torch.all(ctx_n[:, :, 361:] == 0)
    assert torch.equal(ctx_t[:, :361], ref.transpose(1, 2)) and torch.all(ctx_t[:, 361:] == 0)
    for i, L in enumerate(cl):
        assert torch.equal(words_t[i, :L], words[i, :, :L].t().to(torch.bfloat16))
        assert torch.all(words_t[i, L:] == 0)
        assert torch.allclose(wnorm[i, :L], words[i, :, :L].norm(dim=0), rtol=1e-6)
    # word offset 1 (get_local_similarities, gloria_model.py:179)
    lens1 = torch.tensor([96, 40, 3], dtype=torch.int32, device="cuda")
    _, _, w1, _ = ops.tc_prepack(ctx, words, lens1, 96, 1)
    assert torch.equal(w1[1, :40], words[1, :, 1:41].t().to(torch.bfloat16))


@pytest.mark.parametrize("B,seed,scale,lens", [
    (3, 7, 1.0, [97, 41, 5]),
    (3, 7, 0.05, [97, 41, 5]),
    (5, 8, 1.0, [16, 9, 4, 2, 1]),
    (16, 3, 1.0, None),
    (48, 5, 0.05, None),
])
def test_forward_logits(gl, B, seed, scale, lens):
    img_l, txt_l, _, _, cl = gen_inputs(seed, B, 768, 19, 19, 97, cap_lens=lens, scale=scale, dtype=np.float32)
    sim, _, _, _ = gl.local_similarities(cu(img_l), cu(txt_l), cl)
    torch.cuda.synchronize()
    ref = O.local_similarities(img_l.astype(np.float64), txt_l.astype(np.float64), cl)
    err = relerr(sim * 10.0, ref * 10.0)
    print(f"bf16 forward B={B} scale={scale}: logits rel err {err:.3e}")
    assert err < LOGIT_TOL


def test_forward_zero_shot_shape(gl):
    """Rectangular 40 images x 7 prompts, word offset 1, max aggregation (gloria_model.py:171-207)."""
    img_l, txt_l, _, _, _ = gen_inputs(12, 40, 768, 19, 19, 97, dtype=np.float32)
    cl = [14, 9, 8, 6, 5, 4, 3]
    sim, _, _, _ = gl.local_similarities(cu(img_l), cu(txt_l[:7]), cl, 4.0, 5.0, "max", word_offset=1)
    ref = O.get_local_similarities(img_l.astype(np.float64), txt_l[:7].astype(np.float64), cl)
    assert relerr(sim, ref) < LOGIT_TOL


def test_loss_and_gradients(gl):
    """local_loss forward in bf16 on tensor cores; gradients vs the oracle within the 1e-2 gate."""
    B = 16
    img_l, txt_l, _, _, cl = gen_inputs(3, B, 768, 19, 19, 97, dtype=np.float32)
    img, txt = cu(img_l, True), cu(txt_l, True)
    l0, l1, _, _, _, maps = gl.local_loss(img, txt, cl)
    (l0 + l1).backward()
    o0, o1, _, _, _, omaps, _ = O.local_loss(img_l.astype(np.float64), txt_l.astype(np.float64), cl)
    assert relerr(l0, o0) < LOGIT_TOL and relerr(l1, o1) < LOGIT_TOL
    d_img, d_txt = O.local_loss_bwd(img_l.astype(np.float64), txt_l.astype(np.float64), cl)
    assert relerr(img.grad, d_img) < GRAD_TOL
    assert relerr(txt.grad, d_txt) < GRAD_TOL
    assert relerr(maps[3], omaps[3]) < 1e-3
