"""Golden vectors for the zero-shot driver (gloria/gloria.py:186-275), produced by the REAL reference functions
`get_similarities` and `zero_shot_classification` (test infrastructure; run in the build container only).

    python oracle/make_golden_zero_shot.py        -> tests/golden/zero_shot_fp64.npz

The reference driver needs a model with encoders; here the encoders are replaced by look-ups into seeded embedding
tables (the encoders are out of scope), while the similarity methods are the real
GLoRIA.get_local_similarities / get_global_similarities (gloria_model.py:164-207).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.make_golden import OUT, load_reference_model_class, make_model  # noqa: E402


def fake_model(GLoRIA, img_l, img_g, txt_l, txt_g):
    """Real similarity methods, table look-up encoders: `imgs` = image indices, caption_ids[:, 0] = prompt index."""
    m = make_model(GLoRIA)
    m.image_encoder_forward = lambda imgs: (torch.tensor(img_l)[imgs], torch.tensor(img_g)[imgs])
    m.text_encoder_forward = lambda ids, mask, tt: (torch.tensor(txt_l)[ids[:, 0]], torch.tensor(txt_g)[ids[:, 0]], None)
    return m


def class_texts(class_prompts, cap_lens, n_tok):
    """cls_txt_mapping in the reference's format (dict of processed-text dicts)."""
    out = {}
    for name, idx in class_prompts.items():
        ids = torch.zeros((len(idx), n_tok), dtype=torch.long)
        ids[:, 0] = torch.tensor(idx)
        out[name] = dict(caption_ids=ids, attention_mask=torch.ones_like(ids), token_type_ids=torch.zeros_like(ids),
                         cap_lens=[cap_lens[i] for i in idx])
    return out


def inputs(seed=21, n_img=6, n_txt=7, D=48, H=4, W=5, n_tok=13):
    rng = np.random.default_rng(seed)
    img_l = rng.standard_normal((n_img, D, H, W))
    img_g = rng.standard_normal((n_img, D))
    txt_l = rng.standard_normal((n_txt, D, n_tok))
    txt_g = rng.standard_normal((n_txt, D))
    cap_lens = [int(v) for v in rng.integers(2, n_tok - 1, size=n_txt)]      # words after [CLS]; len + 1 <= n_tok
    class_prompts = {"Atelectasis": [0, 1], "Cardiomegaly": [2, 3, 4], "Edema": [5, 6]}
    return img_l, img_g, txt_l, txt_g, cap_lens, class_prompts


def main():
    GLoRIA = load_reference_model_class()
    ref = sys.modules["gloria.gloria"]
    img_l, img_g, txt_l, txt_g, cap_lens, class_prompts = inputs()
    model = fake_model(GLoRIA, img_l, img_g, txt_l, txt_g)
    imgs = torch.arange(img_l.shape[0])
    texts = class_texts(class_prompts, cap_lens, txt_l.shape[2])
    out = dict(img_l=img_l, img_g=img_g, txt_l=txt_l, txt_g=txt_g, cap_lens=np.array(cap_lens),
               class_names=np.array(list(class_prompts)), class_sizes=np.array([len(v) for v in class_prompts.values()]))
    df = ref.zero_shot_classification(model, imgs, texts)
    out["class_similarities"] = df.to_numpy()
    assert list(df.columns) == list(class_prompts)
    for kind in ("both", "local", "global"):
        out[f"sim_{kind}_class1"] = ref.get_similarities(model, imgs, texts["Cardiomegaly"], similarity_type=kind)
    # a single image: the reference skips the normalisation (gloria.py:268)
    out["class_similarities_one_image"] = ref.zero_shot_classification(model, imgs[:1], texts).to_numpy()
    np.savez_compressed(os.path.join(OUT, "zero_shot_fp64.npz"), **out)
    print("written", os.path.join(OUT, "zero_shot_fp64.npz"), {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
