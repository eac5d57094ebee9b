"""Caption lengths that never leave the device, B200, `-m gpu`:

* local_loss / calc_loss with a `DeviceCapLens` give the numbers of the host-list path (loss, gradients, attention maps),
* such a step contains no host synchronisation: it is captured in a CUDA graph and replayed,
* a second backward over a consumed fused training state is caught (NaN), not silently wrong,
* the custom ops trace under torch.compile(fullgraph=True) (data-dependent state size declared in the fake).
"""
import numpy as np
import pytest
import torch

from tests.util import Holder, relerr

pytestmark = pytest.mark.gpu


def _inputs(B, seed, scale=0.05, ragged=True):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    img = (torch.randn((B, 768, 19, 19), device="cuda", generator=gen) * scale)
    txt = (torch.randn((B, 768, 97), device="cuda", generator=gen) * scale)
    rng = np.random.default_rng(seed)
    lens = [int(v) for v in rng.integers(5, 98, size=B)] if ragged else [97] * B
    lens[0] = 97
    for i, L in enumerate(lens):
        txt[i, :, L:] = 0
    return img, txt, lens


@pytest.mark.parametrize("mode,gtol", [("bf16", 2e-3), ("fp32", 1e-5)])
def test_device_cap_lens_equal_host_list(mode, gtol):
    import gloria_nlp_project_b200 as g
    from gloria_nlp_project_b200 import gloria_loss
    g.set_precision(mode)
    try:
        B = 24
        img0, txt0, lens = _inputs(B, 5)
        outs = []
        for cl in (lens, gloria_loss.DeviceCapLens(torch.tensor(lens, device="cuda"))):
            img, txt = img0.clone().requires_grad_(True), txt0.clone().requires_grad_(True)
            l0, l1, _, _, _, maps = gloria_loss.local_loss(img, txt, cl)
            (l0 + 0.7 * l1).backward()
            outs.append((float(l0), float(l1), img.grad.clone(), txt.grad.clone(), maps))
        a, b = outs
        assert isinstance(b[4], gloria_loss.LazyAttnMaps) and isinstance(a[4], list)
        # the device path pads every caption to the word axis (one launch); the host path buckets by length: same math,
        # different tiling of the same sums
        assert abs(a[0] - b[0]) <= gtol * abs(a[0]) and abs(a[1] - b[1]) <= gtol * abs(a[1])
        assert relerr(b[2], a[2]) < 5 * gtol and relerr(b[3], a[3]) < 5 * gtol
        for i in (0, 3, B - 1):
            assert b[4][i].shape == a[4][i].shape
            assert relerr(b[4][i], a[4][i]) < 1e-5
    finally:
        g.set_precision("auto")


def test_calc_loss_with_lazy_sentences_and_segmentation_labels():
    """Attention fine-tune config (gloria_model.py:132-150) fed by LazySentences: supervised-attention term from the padded
    maps (no sync) == the per-caption list form."""
    import gloria_nlp_project_b200 as g
    from gloria_nlp_project_b200 import gloria_loss
    from gloria_nlp_project_b200.gloria_model import GLoRIALossMixin

    class M(GLoRIALossMixin, Holder):
        pass

    g.set_precision("fp32")
    try:
        B = 8
        img0, txt0, lens = _inputs(B, 11)
        gg = torch.Generator(device="cuda").manual_seed(2)
        img_g = torch.randn((B, 768), device="cuda", generator=gg)
        txt_g = torch.randn((B, 768), device="cuda", generator=gg)
        labels = (torch.rand((B, 64, 64), device="cuda", generator=gg) > 0.7).float()
        words = [["[CLS]"] + ["w"] * (L - 1) + ["[SEP]"] for L in lens]            # cap_len = words without '[' + 1 = L

        class Sents(list):
            pass

        lazy = Sents(words)
        lazy.cap_lens = gloria_loss.DeviceCapLens(torch.tensor(lens, device="cuda"))
        res = []
        for sents in (words, lazy):
            m = M(local_loss_weight=0.0, global_loss_weight=0.0, segmentation_loss_weight=1.0)
            img, txt = img0.clone().requires_grad_(True), txt0.clone().requires_grad_(True)
            loss, maps = m.calc_loss(img, img_g, txt, txt_g, sents, segmentation_labels=labels)
            loss.backward()
            res.append((float(loss), img.grad.clone(), txt.grad.clone()))
        assert abs(res[0][0] - res[1][0]) < 1e-5 * abs(res[0][0])
        assert relerr(res[1][1], res[0][1]) < 1e-4 and relerr(res[1][2], res[0][2]) < 1e-4
    finally:
        g.set_precision("auto")


def test_step_is_cuda_graph_capturable():
    """BASELINE.json configs[1] shape (B = 48): local + global loss forward + backward with device-side caption lengths
    captured once in a CUDA graph; the replay reproduces the eager result (up to the order of the kernels' shared-memory
    float atomics, which differs from launch to launch)."""
    import gloria_nlp_project_b200 as g
    from gloria_nlp_project_b200 import gloria_loss
    g.set_precision("bf16")
    try:
        B = 48
        img, txt, lens = _inputs(B, 21)
        gg = torch.Generator(device="cuda").manual_seed(4)
        img_g = torch.randn((B, 768), device="cuda", generator=gg).requires_grad_(True)
        txt_g = torch.randn((B, 768), device="cuda", generator=gg).requires_grad_(True)
        img.requires_grad_(True)
        txt.requires_grad_(True)
        cl = gloria_loss.DeviceCapLens(torch.tensor(lens, device="cuda"))
        leaves = (img, txt, img_g, txt_g)

        def step():
            l0, l1, _, _, _, _ = gloria_loss.local_loss(img, txt, cl)
            g0, g1 = gloria_loss.global_loss(img_g, txt_g)
            loss = l0 + l1 + g0 + g1
            grads = torch.autograd.grad(loss, leaves)
            return loss, grads

        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                loss_e, grads_e = step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss_g, grads_g = step()
        for t in (loss_g, *grads_g):
            t.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert abs(float(loss_g) - float(loss_e)) < 1e-6 * abs(float(loss_e))
        for a, b in zip(grads_g, grads_e):
            assert torch.isfinite(a).all() and relerr(a, b) < 1e-3
    finally:
        g.set_precision("auto")


def test_c_abi_second_backward_over_consumed_state_is_poisoned():
    """Below the Python guard (tests/test_gpu_bf16_parity.py::test_fused_state_is_consumed_once): a caller of the C ABI
    that runs the training backward twice on one state gets NaN, not gradients scaled by dsim twice."""
    import ctypes as C
    import os
    from gloria_nlp_project_b200 import _lib, ops
    L = _lib.lib()
    B, lcap = 6, 97
    img, txt, lens = _inputs(B, 31, ragged=False)
    ctx = img.reshape(B, 768, 361).contiguous()
    dl = torch.tensor(lens, dtype=torch.int32, device="cuda")
    pk = ops.tc_prepack(ctx, txt, dl, lcap, 0)
    n = L.gloria_b200_tc_train_workspace(B, B, 768, 361, lcap)
    ws = torch.empty((n,), dtype=torch.uint8, device="cuda")
    sim = torch.empty((B, B), device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.gloria_b200_tc_local_sim_fwd_train(pk.ctx_h.data_ptr(), pk.ctx_t.data_ptr(), pk.words_h.data_ptr(),
                                                    pk.wnorm.data_ptr(), dl.data_ptr(), B, B, 768, 361, lcap, 4.0, 5.0, 0,
                                                    1e-8, sim.data_ptr(), ws.data_ptr(), n, st), "fwd_train")
    dsim = torch.randn((B, B), device="cuda") * 0.1
    outs = []
    for _ in range(2):
        d_ctx, d_words = torch.empty_like(ctx), torch.empty_like(txt)
        _lib.check(L.gloria_b200_tc_local_sim_bwd_train(pk.ctx_t.data_ptr(), pk.words_t.data_ptr(), dl.data_ptr(), B, B, 768,
                                                        361, 97, lcap, 0, dsim.data_ptr(), d_ctx.data_ptr(),
                                                        d_words.data_ptr(), ws.data_ptr(), n, st), "bwd_train")
        torch.cuda.synchronize()
        outs.append((d_ctx, d_words))
    assert torch.isfinite(outs[0][0]).all() and torch.isfinite(outs[0][1]).all()
    if os.environ.get("GLORIA_B200_BWD_GEMM", "cublas") in ("inflight", "0"):
        assert relerr(outs[1][0], outs[0][0]) < 1e-6            # nothing is mutated in this mode: simply repeatable
    else:
        assert torch.isnan(outs[1][0]).any() and torch.isnan(outs[1][1]).any()


def test_no_grad_forward_allocates_no_training_state():
    """A forward whose inputs do not require grad (evaluation without torch.no_grad) must not build the 41.7 GB-class
    training state: need_grad is the caller's statement, not `torch.is_grad_enabled()`."""
    import gloria_nlp_project_b200 as g
    from gloria_nlp_project_b200 import gloria_loss
    g.set_precision("bf16")
    try:
        img, txt, lens = _inputs(64, 41, ragged=False)
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        sim, _, _, _ = gloria_loss.local_similarities(img, txt, lens)
        torch.cuda.synchronize()
        peak = torch.cuda.max_memory_allocated() - base
        assert peak < 400 << 20, peak             # packed copies only (the training state at B = 64 is ~0.7 GB)
        assert torch.isfinite(sim).all()
    finally:
        g.set_precision("auto")


def test_ops_trace_under_torch_compile_fullgraph():
    from gloria_nlp_project_b200 import ops
    img, txt, lens = _inputs(6, 51, ragged=False)
    ctx = img.reshape(6, 768, 361).contiguous().requires_grad_(True)
    words = txt.clone().requires_grad_(True)
    dl = torch.tensor(lens, dtype=torch.int32, device="cuda")

    def f(c, w, l):
        sim, _, _, _ = ops.local_sim_fwd(c, w, l, 97, 0, 4.0, 5.0, 0, 1e-8, False, False, ops.MODE_BF16, True)
        losses, _, _ = ops.ce_bidir_fwd(sim, 10.0)
        return losses[0] + losses[1]

    ref = f(ctx, words, dl)
    gref = torch.autograd.grad(ref, (ctx, words))
    cf = torch.compile(f, fullgraph=True)
    out = cf(ctx, words, dl)
    gout = torch.autograd.grad(out, (ctx, words))
    assert abs(float(out) - float(ref)) < 1e-5 * abs(float(ref))
    assert relerr(gout[0], gref[0]) < 1e-3 and relerr(gout[1], gref[1]) < 1e-3      # (launch-to-launch atomics order)
