"""Drop-in for the zero-shot driver of the reference (gloria/gloria.py:186-275): `get_similarities` and
`zero_shot_classification` with the reference's signatures and return types.

What changes underneath: the reference calls `get_similarities` once per class, which
re-encodes all images for every class and moves two [N_img, n_prompts] matrices to the host per class.  Here the images
and all classes' prompts are encoded once, every prompt is scored in ONE launch of the packed-prompt kernel
(gloria_b200_tc_local_sim_fwd_packed: up to 8 short prompts share a word tile) plus one global-cosine kernel, and the
"(local + global) / 2 -> max over the class's prompts -> z-score over the images" tail stays on the device; one
[N_img, n_classes] matrix goes to the host.  The encoders themselves are the caller's (out of scope).
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch

from . import gloria_loss, ops

__all__ = ["get_similarities", "zero_shot_classification", "class_similarities"]

_KINDS = ("global", "local", "both")


def _device_similarities(gloria_model, img_emb_l, img_emb_g, text_emb_l, text_emb_g, cap_lens, kind):
    """[N_img, N_txt] on the device: gloria_model.py:164-207 (word slice [1 : L+1], temps 4 / 5, max over words)."""
    out = {}
    with torch.no_grad():
        if kind in ("local", "both"):
            sim, _, _, _ = gloria_loss.local_similarities(
                img_emb_l.detach(), text_emb_l.detach(), cap_lens, 4.0, 5.0, "max",
                no_attn_vec=getattr(gloria_model, "no_attn_vec", None), word_offset=1)
            out["local"] = sim
        if kind in ("global", "both"):
            cosm, _, _ = ops.global_sim_fwd(img_emb_g.detach().float(), text_emb_g.detach().float(), 1e-30)
            out["global"] = cosm
    if kind == "both":
        return (out["local"] + out["global"]) / 2                     # gloria.py:230
    return out[kind]


def _check_inputs(imgs, txts, similarity_type):
    if similarity_type not in _KINDS:                                  # gloria.py:204-216: same errors, same order
        raise RuntimeError("similarity type should be one of ['global', 'local', 'both']")
    if type(txts) == str or type(txts) == list:
        raise RuntimeError("Text input not processed - please use gloria_model.process_text")
    if type(imgs) == str or type(imgs) == list:
        raise RuntimeError("Image input not processed - please use gloria_model.process_img")


def _encode_text(gloria_model, txts):
    with torch.no_grad():
        text_emb_l, text_emb_g, _ = gloria_model.text_encoder_forward(
            txts["caption_ids"], txts["attention_mask"], txts["token_type_ids"])
    return text_emb_l, text_emb_g


def get_similarities(gloria_model, imgs, txts, similarity_type="both"):
    """gloria.py:186-237 -> numpy [N_img, N_txt]."""
    _check_inputs(imgs, txts, similarity_type)
    with torch.no_grad():
        img_emb_l, img_emb_g = gloria_model.image_encoder_forward(imgs)
    text_emb_l, text_emb_g = _encode_text(gloria_model, txts)
    sim = _device_similarities(gloria_model, img_emb_l, img_emb_g, text_emb_l, text_emb_g, txts["cap_lens"],
                               similarity_type)
    return sim.detach().cpu().numpy()


def class_similarities(gloria_model, img_emb_l, img_emb_g, text_emb_l, text_emb_g, cap_lens: Sequence[int],
                       class_sizes: Sequence[int]) -> torch.Tensor:
    """The arithmetic of gloria.py:255-269 on embeddings, on the device: [N_img, n_classes] (prompts of class k are the
    rows sum(class_sizes[:k]) ... of the text embeddings)."""
    sim = _device_similarities(gloria_model, img_emb_l, img_emb_g, text_emb_l, text_emb_g, list(cap_lens), "both")
    cols = [part.max(dim=1).values for part in torch.split(sim, [int(n) for n in class_sizes], dim=1)]   # :262
    cs = torch.stack(cols, dim=1)
    if cs.shape[0] > 1:                                                # :268; utils.normalize: numpy std, ddof = 0
        cs = (cs - cs.mean(dim=0)) / cs.std(dim=0, unbiased=False)
    return cs


def zero_shot_classification(gloria_model, imgs, cls_txt_mapping: Dict[str, dict]):
    """gloria.py:240-275 -> pandas DataFrame [N_img, n_classes] with the class names as columns."""
    import pandas as pd
    names = list(cls_txt_mapping.keys())
    for txt in cls_txt_mapping.values():
        _check_inputs(imgs, txt, "both")
    with torch.no_grad():
        img_emb_l, img_emb_g = gloria_model.image_encoder_forward(imgs)          # once, not once per class
    merged = {k: torch.cat([cls_txt_mapping[n][k] for n in names], 0)
              for k in ("caption_ids", "attention_mask", "token_type_ids")}
    cap_lens = [int(v) for n in names for v in cls_txt_mapping[n]["cap_lens"]]
    sizes = [int(cls_txt_mapping[n]["caption_ids"].shape[0]) for n in names]
    text_emb_l, text_emb_g = _encode_text(gloria_model, merged)
    cs = class_similarities(gloria_model, img_emb_l, img_emb_g, text_emb_l, text_emb_g, cap_lens, sizes)
    return pd.DataFrame(cs.detach().cpu().numpy(), columns=names)
