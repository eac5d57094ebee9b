"""Drop-in replacements for the loss / similarity methods of the reference GLoRIA module
(gloria/models/gloria_model.py:105-211): `_calc_local_loss`, `_calc_global_loss`, `calc_loss`,
`get_global_similarities`, `get_local_similarities`, `get_attn_maps`.

Use either as a mixin placed before the reference class, or patch an existing class / instance:

    from gloria_nlp_project_b200.gloria_model import patch_gloria
    patch_gloria(GLoRIA)            # class-level monkey patch; encoders etc. stay the reference's

The methods read the same attributes the reference's __init__ sets (gloria_model.py:60-75): temp1/2/3,
local_loss_weight, global_loss_weight, segmentation_loss_weight, no_attn_vec, no_attn_loss_weight,
attention_divergence_loss_weight, attention_entropy_loss_weight.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import gloria_loss, ops

__all__ = ["GLoRIALossMixin", "patch_gloria", "cap_lens_from_sents"]


def cap_lens_from_sents(sents):
    """gloria_model.py:107-109: number of words not starting with '[' plus one ([CLS] kept, [SEP] dropped).
    Sentences that come from this package's `text_model.aggregate_tokens` carry the same numbers as a device tensor
    (`sents.cap_lens`, computed by the word-boundary kernel): they are used as they are, and neither the strings nor
    a host copy of the lengths is ever built on the training path."""
    dev = getattr(sents, "cap_lens", None)
    if isinstance(dev, gloria_loss.DeviceCapLens):
        return dev
    return [len([w for w in sent if not w.startswith("[")]) + 1 for sent in sents]


class GLoRIALossMixin:
    def _calc_local_loss(self, img_emb_l, text_emb_l, sents):
        """gloria_model.py:105-123"""
        cap_lens = cap_lens_from_sents(sents)
        return gloria_loss.local_loss(
            img_emb_l,
            text_emb_l,
            cap_lens,
            temp1=self.temp1,
            temp2=self.temp2,
            temp3=self.temp3,
            no_attn_vec=getattr(self, "no_attn_vec", None),
            no_attn_loss_weight=getattr(self, "no_attn_loss_weight", None),
            attention_divergence_loss_weight=getattr(self, "attention_divergence_loss_weight", None),
            attention_entropy_loss_weight=getattr(self, "attention_entropy_loss_weight", None),
        )

    def _calc_global_loss(self, img_emb_g, text_emb_g):
        """gloria_model.py:125-127"""
        return gloria_loss.global_loss(img_emb_g, text_emb_g, temp3=self.temp3)

    def _regularised(self):
        return any(getattr(self, n, None) is not None for n in
                   ("no_attn_loss_weight", "attention_divergence_loss_weight", "attention_entropy_loss_weight"))

    def _diagonal_maps(self, img_emb_l, text_emb_l, sents):
        return gloria_loss.diagonal_attention_maps(img_emb_l, text_emb_l, cap_lens_from_sents(sents), temp1=self.temp1,
                                                   no_attn_vec=getattr(self, "no_attn_vec", None))

    def calc_loss(self, img_emb_l, img_emb_g, text_emb_l, text_emb_g, sents, segmentation_labels=None):
        """gloria_model.py:132-150 -> (loss, attn_maps).

        The reference always runs local_loss over all B^2 pairs; when the contrastive local term has weight 0 and no
        regulariser is configured (the attention fine-tune config) only its diagonal attention maps are used, so only
        the B diagonal pairs are computed here."""
        loss = 0
        no_attn_loss = kl_loss = entropy_loss = 0
        if self.local_loss_weight != 0 or self._regularised():
            l_loss0, l_loss1, no_attn_loss, kl_loss, entropy_loss, attn_maps = self._calc_local_loss(
                img_emb_l, text_emb_l, sents)
            if self.local_loss_weight != 0:
                loss += (l_loss0 + l_loss1) * self.local_loss_weight
        else:
            attn_maps = self._diagonal_maps(img_emb_l, text_emb_l, sents)
        if self.global_loss_weight != 0:
            g_loss0, g_loss1 = self._calc_global_loss(img_emb_g, text_emb_g)
            loss += (g_loss0 + g_loss1) * self.global_loss_weight
        if segmentation_labels is not None and getattr(self, "segmentation_loss_weight", None):
            # supervised attention (gloria_model.py:143-147): word-mean of each diagonal map, nearest upsample to
            # the label resolution, normalise to sum 1, -log of the mass inside the label -- in per-cell form
            loss += gloria_loss.supervised_attention_loss(attn_maps, segmentation_labels) * self.segmentation_loss_weight
        loss += no_attn_loss + kl_loss + entropy_loss
        return loss, attn_maps

    def get_global_similarities(self, img_emb_g, text_emb_g):
        """gloria_model.py:164-169: cosine similarity matrix returned as a CPU float32 tensor (the reference goes
        through sklearn on the host; zero rows give 0)."""
        with torch.no_grad():
            cosm, _, _ = ops.global_sim_fwd(img_emb_g.detach().float(), text_emb_g.detach().float(), 1e-30)
        return cosm.cpu()

    def get_local_similarities(self, img_emb_l, text_emb_l, cap_lens):
        """gloria_model.py:171-207: words [1 : L+1], temp1 = 4.0, temp2 = 5.0, max over words, CPU result."""
        with torch.no_grad():
            sim, _, _, _ = gloria_loss.local_similarities(
                img_emb_l.detach(), text_emb_l.detach(), cap_lens, 4.0, 5.0, "max",
                no_attn_vec=getattr(self, "no_attn_vec", None), word_offset=1)
        return sim.cpu()

    def get_attn_maps(self, img_emb_l, text_emb_l, sents):
        """gloria_model.py:209-211 (the reference runs the whole local_loss for these B maps)"""
        return self._diagonal_maps(img_emb_l, text_emb_l, sents)


_METHODS = ["_calc_local_loss", "_calc_global_loss", "_regularised", "_diagonal_maps", "calc_loss",
            "get_global_similarities", "get_local_similarities", "get_attn_maps"]


def patch_gloria(target):
    """Monkey-patch a reference GLoRIA class (or instance) so its loss path runs on the B200 kernels."""
    import types
    for name in _METHODS:
        fn = getattr(GLoRIALossMixin, name)
        if isinstance(target, type):
            setattr(target, name, fn)
        else:
            setattr(target, name, types.MethodType(fn, target))
    if not isinstance(target, type):
        target.local_loss = gloria_loss.local_loss
        target.global_loss = gloria_loss.global_loss
    return target
