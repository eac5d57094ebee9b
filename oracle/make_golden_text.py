"""Golden vectors for the word-piece aggregation, produced by the REAL reference method
BertEncoder.aggregate_tokens (gloria/models/text_model.py:32-90), called unbound on a stand-in `self` that carries
only `idxtoword` (the method reads nothing else; the BERT weights are not needed).  Build container only.

    python oracle/make_golden_text.py        -> tests/golden/aggregate_tokens.npz
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.make_golden import OUT, load_reference_model_class  # noqa: E402

VOCAB = ["[PAD]", "[CLS]", "[SEP]", "[UNK]", "the", "heart", "is", "enlarged", "cardio", "##meg", "##aly", "no", "pleural",
         "eff", "##usion", "pneumo", "##thorax", "small", "left", "##-sided", "atel", "##ect", "##asis", ".", "##s"]


def captions():
    """Token-id rows (T = 16) covering: plain words, multi-piece words, a leading "##" piece, [SEP] right after [CLS],
    a caption without [SEP] (truncated), a full-length caption, pieces right before [SEP]."""
    w = {t: i for i, t in enumerate(VOCAB)}
    rows = [
        ["[CLS]", "the", "heart", "is", "enlarged", ".", "[SEP]"],
        ["[CLS]", "cardio", "##meg", "##aly", "[SEP]"],
        ["[CLS]", "no", "pleural", "eff", "##usion", "##s", "[SEP]"],
        ["##meg", "##aly", "is", "small", "[SEP]"],
        ["[CLS]", "[SEP]"],
        ["[CLS]", "small", "left", "##-sided", "pneumo", "##thorax", "no", "atel", "##ect", "##asis", "the", "heart", "is",
         "enlarged", "cardio", "##meg"],                                     # no [SEP]: the open word is dropped
        ["[CLS]", "atel", "##ect", "##asis", ".", "no", "eff", "##usion", ".", "the", "heart", "is", "small", ".", "no", "[SEP]"],
        ["[CLS]", "pneumo", "##thorax", "[SEP]", "the", "heart"],            # tokens after [SEP] are ignored
    ]
    T = 16
    ids = np.zeros((len(rows), T), dtype=np.int64)
    for i, r in enumerate(rows):
        ids[i, :len(r)] = [w[t] for t in r]
    return ids


def main():
    load_reference_model_class()
    from gloria.models.text_model import BertEncoder
    ids = captions()
    rng = np.random.default_rng(31)
    emb = rng.standard_normal((ids.shape[0], 4, ids.shape[1], 24))
    stand_in = types.SimpleNamespace(idxtoword={i: t for i, t in enumerate(VOCAB)})
    agg, sents = BertEncoder.aggregate_tokens(stand_in, torch.tensor(emb), torch.tensor(ids))
    np.savez_compressed(os.path.join(OUT, "aggregate_tokens.npz"), vocab=np.array(VOCAB), caption_ids=ids, embeddings=emb,
                        agg=agg.numpy(), sentences=np.array(sents))
    print("written", os.path.join(OUT, "aggregate_tokens.npz"), agg.shape, [len([w for w in s if w != "[PAD]"]) for s in sents])


if __name__ == "__main__":
    main()
