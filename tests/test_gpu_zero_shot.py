"""Zero-shot driver drop-in (gloria_nlp_project_b200/zero_shot.py) vs golden vectors of the real reference driver and
vs the oracle at the full feature sizes (B200, `-m gpu`)."""
import os

import numpy as np
import pytest
import torch

from oracle import gloria_oracle as O
from tests.util import Holder, relerr

pytestmark = pytest.mark.gpu


class FakeModel(Holder):
    """Table look-up encoders (the encoders are out of scope) around the drop-in similarity path: `imgs` = image
    indices, caption_ids[:, 0] = prompt index -- the same stand-in oracle/make_golden_zero_shot.py gives the reference."""

    def __init__(self, img_l, img_g, txt_l, txt_g):
        super().__init__()
        self.t = [torch.tensor(a, dtype=torch.float32, device="cuda") for a in (img_l, img_g, txt_l, txt_g)]
        self.image_calls = 0

    def image_encoder_forward(self, imgs):
        self.image_calls += 1
        return self.t[0][imgs], self.t[1][imgs]

    def text_encoder_forward(self, ids, mask, tt):
        return self.t[2][ids[:, 0]], self.t[3][ids[:, 0]], None


def class_texts(class_sizes, cap_lens, n_tok):
    out, o = {}, 0
    for k, n in enumerate(class_sizes):
        ids = torch.zeros((int(n), n_tok), dtype=torch.long, device="cuda")
        ids[:, 0] = torch.arange(o, o + int(n))
        out[f"class{k}"] = dict(caption_ids=ids, attention_mask=torch.ones_like(ids), token_type_ids=torch.zeros_like(ids),
                                cap_lens=[int(v) for v in cap_lens[o:o + int(n)]])
        o += int(n)
    return out


def test_zero_shot_driver_golden_fp32(golden_dir):
    """fp32 kernels (D = 48 is below the tensor-core tile) against what the reference's own driver returned."""
    import gloria_nlp_project_b200 as g
    from gloria_nlp_project_b200 import zero_shot
    zs = dict(np.load(os.path.join(golden_dir, "zero_shot_fp64.npz")))
    g.set_precision("fp32")
    try:
        m = FakeModel(zs["img_l"], zs["img_g"], zs["txt_l"], zs["txt_g"])
        texts = class_texts(zs["class_sizes"], zs["cap_lens"], zs["txt_l"].shape[2])
        imgs = torch.arange(zs["img_l"].shape[0], device="cuda")
        df = zero_shot.zero_shot_classification(m, imgs, texts)
        assert m.image_calls == 1                                      # the reference encodes once per class
        assert list(df.columns) == list(texts) and df.shape == zs["class_similarities"].shape
        assert relerr(df.to_numpy(), zs["class_similarities"]) < 1e-4  # z-scores of fp32 similarities
        one = zero_shot.zero_shot_classification(m, imgs[:1], texts)
        assert relerr(one.to_numpy(), zs["class_similarities_one_image"]) < 1e-5
        for kind in ("both", "local", "global"):
            got = zero_shot.get_similarities(m, imgs, texts["class1"], similarity_type=kind)
            assert isinstance(got, np.ndarray) and relerr(got, zs[f"sim_{kind}_class1"]) < 1e-5
        with pytest.raises(RuntimeError):
            zero_shot.get_similarities(m, imgs, texts["class1"], similarity_type="cosine")
        with pytest.raises(RuntimeError):
            zero_shot.get_similarities(m, imgs, ["raw text"])
    finally:
        g.set_precision("auto")


def test_zero_shot_driver_tensor_core():
    """5 classes x 5 short prompts against 60 images at D = 768 / 19 x 19 in bf16 mode (packed-prompt kernel) vs the
    oracle; class scores are z-scores, so the tolerance is absolute on O(1) values."""
    import gloria_nlp_project_b200 as g
    from gloria_nlp_project_b200 import zero_shot
    rng = np.random.default_rng(61)
    n_img, sizes, n_tok = 60, [5, 5, 5, 5, 5], 18
    img_l = rng.standard_normal((n_img, 768, 19, 19)).astype(np.float32) * 0.05
    img_g = rng.standard_normal((n_img, 768)).astype(np.float32)
    txt_l = rng.standard_normal((25, 768, n_tok)).astype(np.float32) * 0.05
    txt_g = rng.standard_normal((25, 768)).astype(np.float32)
    cap_lens = [int(v) for v in rng.integers(3, 17, size=25)]
    ref = O.zero_shot_classification(img_l.astype(np.float64), img_g.astype(np.float64), txt_l.astype(np.float64),
                                     txt_g.astype(np.float64), cap_lens, sizes)
    g.set_precision("bf16")
    try:
        m = FakeModel(img_l, img_g, txt_l, txt_g)
        df = zero_shot.zero_shot_classification(m, torch.arange(n_img, device="cuda"), class_texts(sizes, cap_lens, n_tok))
    finally:
        g.set_precision("auto")
    assert float(np.abs(df.to_numpy() - ref).max()) < 2e-2
    assert (df.to_numpy().argmax(1) == ref.argmax(1)).mean() > 0.95
