"""Pin the numpy oracle (oracle/gloria_oracle.py) to outputs of the REAL reference code.

tests/golden/*.npz were written by oracle/make_golden.py, which executes
/root/reference/gloria/loss/gloria_loss.py and the real GLoRIA model methods.  CPU only.
"""
import os

import numpy as np
import pytest

from oracle import gloria_oracle as O
from oracle.make_golden import checksum, gen_inputs


@pytest.fixture(scope="module")
def small(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "small_fp64.npz")))


def close(a, b, tol=1e-10):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)
    assert err < tol, err


def test_inputs_reproducible(small):
    img_l, txt_l, img_g, txt_g, cap_lens = gen_inputs(11, 5, 48, 4, 5, 11, cap_lens=[11, 9, 6, 2, 1])
    assert np.array_equal(img_l, small["img_l"]) and np.array_equal(txt_l, small["txt_l"])
    assert np.array_equal(img_g, small["img_g"]) and np.array_equal(txt_g, small["txt_g"])


def test_cosine_similarity(small):
    close(O.cosine_similarity(small["img_g"], small["txt_g"]), small["cos"])


def test_attention_fn(small):
    B = small["img_l"].shape[0]
    q = np.repeat(small["txt_l"][1:2, :, :9], B, axis=0)
    wc, at = O.attention_fn(q, small["img_l"], 4.0)
    close(wc, small["attn_wctx"])
    close(at, small["attn_map"])
    wc, at = O.attention_fn(q, small["img_l"], 4.0, no_attn_vec=small["nav"])
    close(wc, small["attn_wctx_nav"])
    close(at, small["attn_map_nav"])
    # attention rows: sum to 1 without the no-attn column, < 1 with it (SURVEY §4 property)
    assert np.allclose(small["attn_map"].sum((-1, -2)), 1.0)
    assert np.all(at.sum((-1, -2)) < 1.0)


@pytest.mark.parametrize("tag,kw", [
    ("sum", dict()),
    ("mean", dict(agg="mean", temp1=3.0, temp2=6.0, temp3=7.0)),
    ("reg", dict(no_attn_loss_weight=0.3, attention_divergence_loss_weight=0.2, attention_entropy_loss_weight=0.1)),
    ("ent_only", dict(attention_entropy_loss_weight=1.0, attention_divergence_loss_weight=0.5)),
])
def test_local_loss_forward(small, tag, kw):
    if tag == "reg":
        kw = dict(kw, no_attn_vec=small["nav"])
    cap_lens = small["cap_lens"].tolist()
    l0, l1, na, kl, ent, maps, logits = O.local_loss(small["img_l"], small["txt_l"], cap_lens, **kw)
    close(l0, small[f"local_{tag}_loss0"])
    close(l1, small[f"local_{tag}_loss1"])
    close(na, small[f"local_{tag}_no_attn_loss"])
    close(kl, small[f"local_{tag}_kl_loss"])
    close(ent, small[f"local_{tag}_entropy_loss"])
    for i, m in enumerate(maps):
        close(m, small[f"local_{tag}_att_{i}"])
    if tag == "sum":
        close(logits, small["local_sum_logits"])


@pytest.mark.parametrize("tag,kw", [("sum", dict()), ("mean", dict(agg="mean", temp1=3.0, temp2=6.0, temp3=7.0))])
def test_local_loss_backward_closed_form(small, tag, kw):
    """The hand-derived backward equals the reference's autograd."""
    d_img, d_txt = O.local_loss_bwd(small["img_l"], small["txt_l"], small["cap_lens"].tolist(), g0=1.0, g1=0.7, **kw)
    close(d_img, small[f"local_{tag}_d_img"], 1e-9)
    close(d_txt, small[f"local_{tag}_d_txt"], 1e-9)
    # padded word columns receive exactly zero gradient
    for i, L in enumerate(small["cap_lens"]):
        assert np.all(d_txt[i, :, L:] == 0)


@pytest.mark.parametrize("tag,use_nav,kw", [
    ("reg", True, dict(no_attn_loss_weight=0.3, attention_divergence_loss_weight=0.2, attention_entropy_loss_weight=0.1)),
    ("ent_only", False, dict(attention_entropy_loss_weight=1.0, attention_divergence_loss_weight=0.5)),
    ("sum", False, dict()),
])
def test_local_loss_full_backward_with_regularisers(small, tag, use_nav, kw):
    """Closed-form gradient INCLUDING the no-attn / symmetric-KL / entropy regularisers (and d no_attn_vec) equals the
    real reference's autograd (goldens local_reg_* / local_ent_only_* written by oracle/make_golden.py)."""
    nav = small["nav"] if use_nav else None
    d_img, d_txt, d_nav = O.local_loss_full_bwd(small["img_l"], small["txt_l"], small["cap_lens"].tolist(), g0=1.0, g1=0.7,
                                                no_attn_vec=nav, **kw)
    close(d_img, small[f"local_{tag}_d_img"], 1e-9)
    close(d_txt, small[f"local_{tag}_d_txt"], 1e-9)
    if use_nav:
        close(d_nav, small["local_reg_d_nav"], 1e-9)
    else:
        assert d_nav is None


def test_zero_word_vector(small):
    txt = small["txt_l"].copy()
    txt[2, :, 3] = 0.0
    cl = small["cap_lens"].tolist()
    l0, l1, *_ = O.local_loss(small["img_l"], txt, cl)
    close(l0, small["zero_word_loss0"])
    close(l1, small["zero_word_loss1"])
    d_img, d_txt = O.local_loss_bwd(small["img_l"], txt, cl)
    close(d_img, small["zero_word_d_img"], 1e-9)
    close(d_txt, small["zero_word_d_txt"], 1e-9)


def test_global_loss(small):
    l0, l1, _ = O.global_loss(small["img_g"], small["txt_g"])
    close(l0, small["global_loss0"])
    close(l1, small["global_loss1"])
    dc, dr = O.global_loss_bwd(small["img_g"], small["txt_g"], g0=1.0, g1=0.7)
    close(dc, small["d_img_g"], 1e-9)
    close(dr, small["d_txt_g"], 1e-9)


def test_model_methods(small):
    cl = small["cap_lens"].tolist()
    _, _, _, _, _, maps, _ = O.local_loss(small["img_l"], small["txt_l"], cl)
    for i, m in enumerate(maps):
        close(m, small[f"model_att_{i}"])
    close(O.get_local_similarities(small["img_l"], small["txt_l"][:3], small["zs_cap_lens"].tolist()),
          small["zs_local"], 1e-6)    # reference returns float32 via torch.Tensor
    close(O.get_global_similarities(small["img_g"], small["txt_g"][:3]), small["zs_global"], 1e-6)


def test_calc_loss_with_segmentation(small):
    cl = small["cap_lens"].tolist()
    l0, l1, _, _, _, maps, _ = O.local_loss(small["img_l"], small["txt_l"], cl)
    g0, g1, _ = O.global_loss(small["img_g"], small["txt_g"])
    seg = O.segmentation_attention_loss(maps, small["seg_labels"])
    close(l0 + l1 + g0 + g1 + 0.5 * seg, small["calc_loss"])
    close(seg, small["ft_loss"])


def test_attention_finetune_gradients(small):
    """imagenome_attn_finetune: only the supervised-attention term, gradient enters through the diagonal maps."""
    cl = small["cap_lens"].tolist()
    img, txt, seg = small["img_l"], small["txt_l"], small["seg_labels"].astype(np.float64)
    B, _, h, w = img.shape
    H, W = seg.shape[1:]
    _, _, _, _, _, maps, _ = O.local_loss(img, txt, cl)
    # d seg_loss / d att_maps, closed form: counts of label pixels per low-res cell
    iy = np.minimum((np.arange(H) * (h / H)).astype(int), h - 1)
    ix = np.minimum((np.arange(W) * (w / W)).astype(int), w - 1)
    cnt_all = np.zeros((h, w))
    np.add.at(cnt_all, (iy[:, None].repeat(W, 1), ix[None].repeat(H, 0)), 1.0)
    d_maps = []
    for i, m in enumerate(maps):
        mm = m[0].mean(0)
        cnt_lab = np.zeros((h, w))
        np.add.at(cnt_lab, (iy[:, None].repeat(W, 1), ix[None].repeat(H, 0)), seg[i])
        num, den = (cnt_lab * mm).sum(), (cnt_all * mm).sum()
        d_mm = -(cnt_lab / num - cnt_all / den) / B
        d_maps.append(np.broadcast_to(d_mm / m.shape[1], m.shape[1:]).copy()[None])
    d_img, d_txt = O.local_loss_bwd(img, txt, cl, g0=0.0, g1=0.0, d_att_maps=d_maps)
    close(d_img, small["ft_d_img_l"], 1e-8)
    close(d_txt, small["ft_d_txt_l"], 1e-8)


@pytest.mark.parametrize("tag,scale", [("unit", 1.0), ("small", 0.05)])
def test_full_size_dims(golden_dir, tag, scale):
    g = dict(np.load(os.path.join(golden_dir, f"full_{tag}.npz")))
    img_l, txt_l, img_g, txt_g, cap_lens = gen_inputs(7, 3, 768, 19, 19, 97, cap_lens=[97, 41, 5], scale=scale)
    close(checksum(img_l, txt_l, img_g, txt_g), g["input_checksum"], 1e-12)
    l0, l1, _, _, _, maps, logits = O.local_loss(img_l, txt_l, cap_lens)
    close(logits, g["f64_logits"], 1e-9)
    close(l0, g["f64_loss0"], 1e-9)
    close(l1, g["f64_loss1"], 1e-9)
    close(maps[1], g["f64_att_1"], 1e-9)
    d_img, d_txt = O.local_loss_bwd(img_l, txt_l, cap_lens)
    close(d_img[:, ::16, ::3, ::3], g["f64_d_img_sub"], 1e-8)
    close(d_txt[:, ::16, ::4], g["f64_d_txt_sub"], 1e-8)
    close(checksum(d_img), g["f64_d_img_cs"], 1e-7)
    # the fp32 reference itself sits within ~1e-5 of fp64 on logits: this is the floor for the fp32-mode gate
    rel = np.max(np.abs(g["f32_logits"] - g["f64_logits"])) / np.max(np.abs(g["f64_logits"]))
    assert rel < 2e-5, rel
    # oracle run in float32 agrees with the float32 reference within fp32 round-off
    l0f, l1f, _, _, _, _, logits_f = O.local_loss(img_l.astype(np.float32), txt_l.astype(np.float32), cap_lens)
    close(logits_f, g["f32_logits"], 2e-5)
    zs = O.get_local_similarities(img_l, txt_l, [c - 1 for c in cap_lens if c > 1] + [3])
    close(zs, g["zs_local_f32"], 2e-5)


# ------------------------------------------------------------------------------------------------------------
# the torch-CPU restatement used as bench.py's multi-threaded CPU baseline is pinned to the same vectors
# ------------------------------------------------------------------------------------------------------------
def test_torch_port_matches_reference(small):
    import torch
    from oracle import gloria_oracle_torch as T
    cl = small["cap_lens"].tolist()
    ti = torch.tensor(small["img_l"], requires_grad=True)
    tw = torch.tensor(small["txt_l"], requires_grad=True)
    l0, l1, maps, logits = T.local_loss(ti, tw, cl)
    (l0 + 0.7 * l1).backward()
    close(l0.item(), small["local_sum_loss0"])
    close(l1.item(), small["local_sum_loss1"])
    close(logits.detach().numpy(), small["local_sum_logits"])
    close(ti.grad.numpy(), small["local_sum_d_img"], 1e-9)
    close(tw.grad.numpy(), small["local_sum_d_txt"], 1e-9)
    for i, m in enumerate(maps):
        close(m.detach().numpy(), small[f"local_sum_att_{i}"])
    tg = torch.tensor(small["img_g"], requires_grad=True)
    tt = torch.tensor(small["txt_g"], requires_grad=True)
    g0, g1 = T.global_loss(tg, tt)
    (g0 + 0.7 * g1).backward()
    close(g0.item(), small["global_loss0"])
    close(g1.item(), small["global_loss1"])
    close(tg.grad.numpy(), small["d_img_g"], 1e-9)
    close(T.cosine_similarity(torch.tensor(small["img_g"]), torch.tensor(small["txt_g"])).numpy(), small["cos"])
    B = small["img_l"].shape[0]
    q = torch.tensor(small["txt_l"][1:2, :, :9]).repeat(B, 1, 1)
    wc, at = T.attention_fn(q, torch.tensor(small["img_l"]), 4.0)
    close(wc.numpy(), small["attn_wctx"])
    close(at.numpy(), small["attn_map"])


# ---------------------------------------------------------------------------------------------
# zero-shot driver (gloria/gloria.py:186-275); golden written by oracle/make_golden_zero_shot.py
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def zs(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "zero_shot_fp64.npz")))


def test_zero_shot_get_similarities(zs):
    sl = slice(2, 5)                                   # prompts of the second class
    cl = [int(v) for v in zs["cap_lens"][sl]]
    for kind in ("both", "local", "global"):
        got = O.get_similarities(zs["img_l"], zs["img_g"], zs["txt_l"][sl], zs["txt_g"][sl], cl, kind)
        # the reference rounds the global matrix to float32 (`torch.Tensor(...)`, gloria_model.py:169)
        close(got, zs[f"sim_{kind}_class1"], 1e-10 if kind == "local" else 1e-6)


def test_zero_shot_classification(zs):
    cl = [int(v) for v in zs["cap_lens"]]
    got = O.zero_shot_classification(zs["img_l"], zs["img_g"], zs["txt_l"], zs["txt_g"], cl, zs["class_sizes"])
    close(got, zs["class_similarities"], 1e-5)        # z-scores of values carrying that float32 rounding
    one = O.zero_shot_classification(zs["img_l"][:1], zs["img_g"][:1], zs["txt_l"], zs["txt_g"], cl, zs["class_sizes"])
    close(one, zs["class_similarities_one_image"], 1e-6)     # a single image is not normalised (gloria.py:268)


# ---------------------------------------------------------------------------------------------
# word-piece aggregation (text_model.py:32-90); golden written by oracle/make_golden_text.py
# ---------------------------------------------------------------------------------------------
def test_aggregate_tokens(golden_dir):
    g = np.load(os.path.join(golden_dir, "aggregate_tokens.npz"))
    idxtoword = {i: str(w) for i, w in enumerate(g["vocab"])}
    agg, sents = O.aggregate_tokens(g["embeddings"], g["caption_ids"], idxtoword)
    close(agg, g["agg"], 1e-14)
    assert [list(s) for s in sents] == [[str(w) for w in s] for s in g["sentences"]]


def test_aggregate_tokens_host_sentences(golden_dir):
    """The drop-in builds the word strings on the host from one read-back of the ids (no tensors): same strings, and the
    caption lengths the loss derives from them (gloria_model.py:107-109)."""
    from gloria_nlp_project_b200.text_model import _sentences, cap_lens_from_sents
    g = np.load(os.path.join(golden_dir, "aggregate_tokens.npz"))
    idxtoword = {i: str(w) for i, w in enumerate(g["vocab"])}
    sents = _sentences(g["caption_ids"].tolist(), idxtoword)
    assert sents == [[str(w) for w in s] for s in g["sentences"]]
    assert cap_lens_from_sents(sents)[:3] == [6, 2, 4]
