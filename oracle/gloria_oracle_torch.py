"""Multi-threaded CPU restatement of the GLoRIA loss path in torch (ATen on the host cores, autograd backward).

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference (strongbeamsprout/gloria-nlp-project) computes this path with
ATen ops and autograd; this file restates that op sequence (caption loop, two bmm, two softmaxes, three reductions,
two cross entropies -- gloria/loss/gloria_loss.py:11-63, 66-88, 99-170) so that `bench.py`'s `cpu_baseline` /
`--impl reference` legs can time "the reference's own CPU loss" on the GPU box, where /root/reference does not
exist.  It is never imported by the product package.

Parity status: PINNED -- tests/test_oracle_golden.py checks it against tests/golden/*.npz (outputs of the real
reference functions, see oracle/make_golden.py) next to the numpy oracle.
"""
from __future__ import annotations

import torch

__all__ = ["cosine_similarity", "attention_fn", "local_similarities", "local_loss", "global_loss", "loss_step"]


def cosine_similarity(x1, x2, dim=1, eps=1e-8):
    """gloria_loss.py:11-16"""
    num = (x1 * x2).sum(dim)
    den = (x1.norm(2, dim) * x2.norm(2, dim)).clamp(min=eps)
    return (num / den).squeeze()


def attention_fn(query, context, temp1):
    """gloria_loss.py:19-63 (no_attn_vec branch omitted: the timed configs do not use it).
    query [B, D, L], context [B, D, H, W] -> weighted context [B, D, L], attention [B, L, H, W]."""
    B, D, L = query.shape
    H, W = context.shape[2], context.shape[3]
    S = H * W
    ctx = context.reshape(B, D, S)
    ctx_t = ctx.transpose(1, 2).contiguous()                         # :35  (the reference copies it every call)
    scores = torch.bmm(ctx_t, query)                                 # :40  [B, S, L]
    p = torch.softmax(scores.reshape(B * S, L), dim=-1).reshape(B, S, L)      # :42-44 softmax over words
    a = torch.softmax(p.transpose(1, 2).contiguous().reshape(B * L, S) * temp1, dim=-1)   # :46-52 over regions
    a_t = a.reshape(B, L, S).transpose(1, 2).contiguous()            # :53-55 [B, S, L]
    wctx = torch.bmm(ctx, a_t)                                       # :59  [B, D, L]
    return wctx, a.reshape(B, L, H, W)


def _ce_arange(logits):
    n = logits.shape[0]
    return torch.nn.functional.cross_entropy(logits, torch.arange(n))


def local_similarities(img_features, words_emb, cap_lens, temp1=4.0, temp2=5.0, agg="sum"):
    """The [B_img, B_cap] matrix of gloria_loss.py:116-162 (before the temp3 scale) and the diagonal maps."""
    B = img_features.shape[0]
    sims, maps = [], []
    for i in range(words_emb.shape[0]):                              # :116
        L = int(cap_lens[i])
        word = words_emb[i, :, :L].unsqueeze(0).contiguous().repeat(B, 1, 1)    # :119-123
        wctx, attn = attention_fn(word, img_features, temp1)         # :126
        if i < B:
            maps.append(attn[i].unsqueeze(0).contiguous())           # :141-143
        w2 = word.transpose(1, 2).contiguous().reshape(B * L, -1)    # :144-148
        c2 = wctx.transpose(1, 2).contiguous().reshape(B * L, -1)
        r = cosine_similarity(w2, c2).reshape(B, L)                  # :150-151
        e = (r * temp2).exp()                                        # :153
        r = e.sum(1, keepdim=True) if agg == "sum" else e.mean(1, keepdim=True)   # :154-157
        sims.append(r.log())                                         # :158
    return torch.cat(sims, 1), maps                                  # :162


def local_loss(img_features, words_emb, cap_lens, temp1=4.0, temp2=5.0, temp3=10.0, agg="sum"):
    """gloria_loss.py:99-170 without the optional regularisers -> (loss0, loss1, att_maps, logits)."""
    sim, maps = local_similarities(img_features, words_emb, cap_lens, temp1, temp2, agg)
    logits = sim * temp3                                             # :164
    return _ce_arange(logits), _ce_arange(logits.t()), maps, logits  # :169-170


def global_loss(cnn_code, rnn_code, eps=1e-8, temp3=10.0):
    """gloria_loss.py:66-88"""
    n1 = cnn_code.norm(2, dim=1, keepdim=True)
    n2 = rnn_code.norm(2, dim=1, keepdim=True)
    scores0 = cnn_code @ rnn_code.t() / (n1 @ n2.t()).clamp(min=eps) * temp3
    return _ce_arange(scores0), _ce_arange(scores0.t())


def loss_step(img_l, txt_l, img_g, txt_g, cap_lens, backward=True):
    """One fwd(+bwd) pass of local + global loss on CPU tensors; returns the loss value."""
    leaves = [t.detach().clone().requires_grad_(backward) for t in (img_l, txt_l, img_g, txt_g)]
    l0, l1, _, _ = local_loss(leaves[0], leaves[1], cap_lens)
    g0, g1 = global_loss(leaves[2], leaves[3])
    loss = l0 + l1 + g0 + g1
    if backward:
        loss.backward()
    return float(loss.detach()), [t.grad for t in leaves]
