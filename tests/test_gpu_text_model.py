"""Word-piece aggregation on the device (gloria_nlp_project_b200/text_model.py) vs golden vectors of the real
BertEncoder.aggregate_tokens and vs the oracle on random token streams (B200, `-m gpu`).  Sums of at most a few
word pieces in fp32: results are bit-exact up to the summation order (the kernel adds the pieces in token order, as
torch.stack(...).sum(0) does for <= 8 rows), so the fp32 gate is 1e-6; fp16 / bf16 round once."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import gloria_oracle as O
from tests.util import relerr

pytestmark = pytest.mark.gpu


def test_aggregate_tokens_golden(golden_dir):
    from gloria_nlp_project_b200 import text_model
    g = np.load(os.path.join(golden_dir, "aggregate_tokens.npz"))
    idxtoword = {i: str(w) for i, w in enumerate(g["vocab"])}
    enc = types.SimpleNamespace(idxtoword=idxtoword)
    text_model.patch_bert_encoder(enc)
    emb = torch.tensor(g["embeddings"], dtype=torch.float32, device="cuda", requires_grad=True)
    ids = torch.tensor(g["caption_ids"], device="cuda")
    agg, sents = enc.aggregate_tokens(emb, ids)
    assert agg.shape == emb.shape
    assert relerr(agg, g["agg"]) < 1e-6
    assert sents == [[str(w) for w in s] for s in g["sentences"]]
    assert text_model.cap_lens_from_sents(sents).tolist() == \
        [len([w for w in s if not str(w).startswith("[")]) + 1 for s in g["sentences"]]
    # backward = gather: compare with autograd through a dense torch restatement of the same sums
    wgt = torch.randn(agg.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    (agg * wgt).sum().backward()
    ref_in = torch.tensor(g["embeddings"], dtype=torch.float64, requires_grad=True)
    ref_out, _ = torch_aggregate(ref_in, g["caption_ids"], idxtoword)
    (ref_out * wgt.double().cpu()).sum().backward()
    assert relerr(emb.grad, ref_in.grad) < 1e-6


def torch_aggregate(emb, ids, idxtoword):
    """Dense torch restatement (differentiable) of the oracle's sums, for the gradient check."""
    B, layers, T, D = emb.shape
    rows = []
    for b in range(B):
        words, bank = [], []
        for t in range(T):
            w = idxtoword[int(ids[b, t])]
            if w == "[SEP]":
                words.append(bank); words.append([t]); bank = None
                break
            if not w.startswith("##"):
                if bank:
                    words.append(bank)
                bank = [t]
            else:
                bank = (bank or []) + [t]
        out = [emb[b, :, k].sum(1) if len(k) else emb[b, :, 0] * 0 for k in words]
        out += [emb[b, :, 0] * 0] * (T - len(out))
        rows.append(torch.stack(out, 1))
    return torch.stack(rows, 0), None


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-6), (torch.float16, 2e-3), (torch.bfloat16, 1.6e-2)])
def test_aggregate_tokens_random_streams(dtype, tol):
    """B = 48 captions x 97 tokens x 4 layers x 768 (the pretraining shapes): random word / piece streams with a [SEP]
    at a random position (some captions have none), against the oracle."""
    from gloria_nlp_project_b200 import text_model
    rng = np.random.default_rng(9)
    vocab = ["[PAD]", "[CLS]", "[SEP]"] + [f"w{i}" for i in range(40)] + [f"##p{i}" for i in range(20)]
    idxtoword = dict(enumerate(vocab))
    B, layers, T, D = 48, 4, 97, 768
    ids = np.zeros((B, T), dtype=np.int64)
    for b in range(B):
        n = int(rng.integers(3, T + 1))
        body = rng.integers(3, len(vocab), size=n)
        ids[b, :n] = body
        ids[b, 0] = 1
        if b % 7 != 3 and n < T:
            ids[b, n - 1] = 2                                        # [SEP]; every 7th caption is left without one
    emb = rng.standard_normal((B, layers, T, D)).astype(np.float32)
    emb_t = torch.tensor(emb, device="cuda").to(dtype)
    table = text_model.VocabTable(idxtoword)
    agg, sents = text_model.aggregate_tokens(emb_t, torch.tensor(ids), table)        # ids on the host, as the loader has them
    ref, ref_sents = O.aggregate_tokens(emb_t.float().cpu().numpy().astype(np.float64), ids, idxtoword)
    assert relerr(agg.float(), ref) < tol
    # caption lengths straight from the word-boundary kernel (no strings involved) == gloria_model.py:107-109 on the strings
    dev_lens = text_model.cap_lens_from_sents(sents)
    assert dev_lens.tensor.is_cuda and sents._built is None            # nothing was materialised on the host so far
    assert dev_lens.tolist() == [len([w for w in s if not w.startswith("[")]) + 1 for s in ref_sents]
    assert sents == ref_sents
