"""CPU oracle for the GLoRIA local/global similarity + loss path.

TEST INFRASTRUCTURE ONLY.  This module is a plain-numpy restatement of the
reference algorithm (strongbeamsprout/gloria-nlp-project); it is the checker for
the CUDA path and is never shipped or measured as the product.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.

Parity status: PINNED.  `oracle/make_golden.py` executes the *real* reference
functions (imported from /root/reference) on seeded inputs and stores their
outputs under `tests/golden/`; `tests/test_oracle_golden.py` checks every
function below against those vectors (the reference itself ships no tests or
golden vectors).

All functions take/return numpy arrays and work in the dtype of their inputs
(float64 for tight checks, float32 for tolerance checks).  Reference citations
are `file:line` relative to /root/reference.
"""

from __future__ import annotations

import numpy as np

__all__ = [
    "cosine_similarity",
    "attention_fn",
    "cross_entropy_arange",
    "global_loss",
    "global_loss_bwd",
    "local_similarities",
    "local_loss",
    "local_sim_pair_bwd",
    "local_loss_bwd",
    "local_loss_full_bwd",
    "get_local_similarities",
    "get_global_similarities",
    "segmentation_attention_loss",
    "kl_divergence",
    "entropy",
]


# --------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------
def _softmax(x, axis):
    m = np.max(x, axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / np.sum(e, axis=axis, keepdims=True)


def _logsumexp(x, axis):
    m = np.max(x, axis=axis, keepdims=True)
    return np.squeeze(m, axis) + np.log(np.sum(np.exp(x - m), axis=axis))


def cosine_similarity(x1, x2, dim=1, eps=1e-8):
    """gloria/loss/gloria_loss.py:11-16  (sum(x1*x2) / clamp(|x1||x2|, eps)).squeeze()"""
    w12 = np.sum(x1 * x2, axis=dim)
    w1 = np.sqrt(np.sum(x1 * x1, axis=dim))
    w2 = np.sqrt(np.sum(x2 * x2, axis=dim))
    return np.squeeze(w12 / np.maximum(w1 * w2, eps))


def attention_fn(query, context, temp1, no_attn_vec=None):
    """gloria/loss/gloria_loss.py:19-63.

    query   [B, D, L]; context [B, D, H, W]
    returns weightedContext [B, D, L], attn [B, L, H, W]
    The attention is a double softmax: over the L words for every region
    (:42-43), then x temp1 and over the regions for every word (:51-52).
    """
    B, D, L = query.shape
    ih, iw = context.shape[2], context.shape[3]
    S = ih * iw
    ctx = context.reshape(B, D, S)                                  # :30
    if no_attn_vec is not None:                                     # :31-34
        v = np.broadcast_to(no_attn_vec.reshape(1, D, 1), (B, D, 1))
        ctx = np.concatenate([v, ctx], axis=2)
        S += 1
    scores = np.einsum("bds,bdl->bsl", ctx, query)                  # :35-40  bmm(contextT, query)
    p = _softmax(scores, axis=2)                                    # :42-43  softmax over words
    a = _softmax(np.swapaxes(p, 1, 2) * temp1, axis=2)              # :46-53  [B, L, S] softmax over regions
    wctx = np.einsum("bds,bls->bdl", ctx, a)                        # :55-59  bmm(context, attnT)
    if no_attn_vec is not None:                                     # :60-61
        a = a[:, :, 1:]
    return wctx, a.reshape(B, L, ih, iw)                            # :63


def cross_entropy_arange(logits):
    """nn.CrossEntropyLoss()(logits, arange(n)) with mean reduction (gloria_loss.py:86-87,169-170)."""
    n = logits.shape[0]
    lse = _logsumexp(logits, axis=1)
    return np.mean(lse - logits[np.arange(n), np.arange(n)])


def _cross_entropy_arange_grad(logits):
    """d mean-CE / d logits for labels = arange."""
    n = logits.shape[0]
    g = _softmax(logits, axis=1)
    g[np.arange(n), np.arange(n)] -= 1.0
    return g / n


def global_loss(cnn_code, rnn_code, eps=1e-8, temp3=10.0):
    """gloria/loss/gloria_loss.py:66-88.  Returns (loss0, loss1, scores0)."""
    n1 = np.sqrt(np.sum(cnn_code * cnn_code, axis=1, keepdims=True))   # :75
    n2 = np.sqrt(np.sum(rnn_code * rnn_code, axis=1, keepdims=True))   # :76
    scores0 = cnn_code @ rnn_code.T                                     # :78
    norm0 = n1 @ n2.T                                                   # :79
    scores0 = scores0 / np.maximum(norm0, eps) * temp3                  # :80
    loss0 = cross_entropy_arange(scores0)                               # :86
    loss1 = cross_entropy_arange(scores0.T)                             # :87
    return loss0, loss1, scores0


def global_loss_bwd(cnn_code, rnn_code, eps=1e-8, temp3=10.0, g0=1.0, g1=1.0):
    """Closed-form gradient of g0*loss0 + g1*loss1 of `global_loss` w.r.t. both codes."""
    n1 = np.sqrt(np.sum(cnn_code * cnn_code, axis=1, keepdims=True))
    n2 = np.sqrt(np.sum(rnn_code * rnn_code, axis=1, keepdims=True))
    dots = cnn_code @ rnn_code.T
    norm0 = n1 @ n2.T
    den = np.maximum(norm0, eps)
    scores0 = dots / den * temp3
    dsc = g0 * _cross_entropy_arange_grad(scores0) + g1 * _cross_entropy_arange_grad(scores0.T).T
    ddots = dsc * temp3 / den
    live = (norm0 >= eps).astype(cnn_code.dtype)          # clamp passes gradient only when un-clamped
    dnorm0 = -dsc * temp3 * dots / (den * den) * live
    dn1 = dnorm0 @ n2                                      # [B,1]
    dn2 = dnorm0.T @ n1
    with np.errstate(divide="ignore", invalid="ignore"):
        u1 = np.where(n1 > 0, cnn_code / n1, 0.0)          # torch.norm backward is 0 at the zero vector
        u2 = np.where(n2 > 0, rnn_code / n2, 0.0)
    dc = ddots @ rnn_code + dn1 * u1
    dr = ddots.T @ cnn_code + dn2 * u2
    return dc, dr


def kl_divergence(attn1, attn2):
    """gloria/loss/gloria_loss.py:91-92"""
    return np.sum(attn1 * np.log(attn1 / attn2), axis=-1)


def entropy(attn):
    """gloria/loss/gloria_loss.py:95-96"""
    return -np.sum(attn * np.log(attn), axis=-1)


# --------------------------------------------------------------------------------------
# local similarity: forward
# --------------------------------------------------------------------------------------
def _pair_block(ctx, w, temp1):
    """Everything `attention_fn` + `cosine_similarity` compute for one caption against all images.

    ctx [B, D, S] (no-attn column already prepended), w [D, L].
    Returns dict of per-image intermediates.
    """
    scores = np.matmul(np.swapaxes(ctx, 1, 2), w)            # S_[s,l] = sum_d ctx[d,s] w[d,l]
    p = _softmax(scores, axis=2)                            # softmax over words
    a = _softmax(np.swapaxes(p, 1, 2) * temp1, axis=2)      # [B, L, S]
    c = np.matmul(ctx, np.swapaxes(a, 1, 2))                 # weightedContext [B, D, L]
    dot = np.sum(w[None] * c, axis=1)
    nw = np.sqrt(np.sum(w * w, axis=0))[None, :]            # [1, L]
    nc = np.sqrt(np.sum(c * c, axis=1))                     # [B, L]
    return dict(scores=scores, p=p, a=a, c=c, dot=dot, nw=nw, nc=nc)


def local_similarities(img_features, words_emb, cap_lens, temp1=4.0, temp2=5.0, agg="sum",
                       no_attn_vec=None, word_offset=0, eps=1e-8, want_attn=False):
    """The B_img x B_cap matrix that gloria_loss.py:116-162 builds (before the temp3 scale at :164).

    word_offset=0, agg in {"sum","mean"} -> training (`local_loss`, :122,153-158);
    word_offset=1, agg="max" -> `GLoRIA.get_local_similarities` (gloria_model.py:179,198-201).
    Returns sim [B_img, B_cap]; with want_attn also a list (per caption) of A [B_img, L, S(+1)].
    """
    Bi, D = img_features.shape[:2]
    ctx = img_features.reshape(Bi, D, -1)
    if no_attn_vec is not None:
        v = np.broadcast_to(no_attn_vec.reshape(1, D, 1), (Bi, D, 1))
        ctx = np.concatenate([v, ctx], axis=2)
    Bc = words_emb.shape[0]
    sim = np.zeros((Bi, Bc), dtype=img_features.dtype)
    attns = []
    for i in range(Bc):
        L = int(cap_lens[i])
        w = words_emb[i, :, word_offset:word_offset + L]
        blk = _pair_block(ctx, w, temp1)
        r = blk["dot"] / np.maximum(blk["nw"] * blk["nc"], eps)       # cosine_similarity :11-16,150
        e = np.exp(r * temp2)                                          # :153
        if agg == "sum":
            sim[:, i] = np.log(np.sum(e, axis=1))                      # :155,158
        elif agg == "mean":
            sim[:, i] = np.log(np.mean(e, axis=1))                     # :157,158
        elif agg == "max":
            sim[:, i] = np.log(np.max(e, axis=1))                      # gloria_model.py:199-201
        else:
            raise ValueError(agg)
        if want_attn:
            attns.append(blk["a"])
    return (sim, attns) if want_attn else sim


def local_loss(img_features, words_emb, cap_lens, temp1=4.0, temp2=5.0, temp3=10.0, agg="sum",
               no_attn_vec=None, no_attn_loss_weight=None, attention_divergence_loss_weight=None,
               attention_entropy_loss_weight=None):
    """gloria/loss/gloria_loss.py:99-201.  Same 6-tuple as the reference, plus the logits as a 7th item."""
    Bi = img_features.shape[0]
    ih, iw = img_features.shape[2], img_features.shape[3]
    sim, attns = local_similarities(img_features, words_emb, cap_lens, temp1, temp2, agg,
                                    no_attn_vec=no_attn_vec, want_attn=True)
    has_v = no_attn_vec is not None
    att_maps = []
    no_attn_scores, flat, ents = [], [], []
    for i, a in enumerate(attns):
        a_ret = a[:, :, 1:] if has_v else a                            # :60-61
        if no_attn_loss_weight is not None:                            # :129-130
            no_attn_scores.append(np.log(1 - a_ret.sum(-1).mean(-1))[:, None])
        if attention_divergence_loss_weight is not None or attention_entropy_loss_weight is not None:
            f = a_ret.mean(1)                                          # :132
            if has_v:
                f = np.concatenate([1 - f.sum(-1, keepdims=True), f], -1)   # :133-135
            if attention_entropy_loss_weight is not None:
                ents.append(entropy(f)[:, None])                       # :137
            if attention_divergence_loss_weight is not None:
                flat.append(f[:, None])                                # :139
        att_maps.append(a_ret[i].reshape(1, -1, ih, iw))               # :141-143
    similarities = sim * temp3                                         # :164
    loss0 = cross_entropy_arange(similarities)                         # :169
    loss1 = cross_entropy_arange(similarities.T)                       # :170
    mask = np.eye(Bi, dtype=bool)                                      # :172
    if no_attn_loss_weight is not None:                                # :173-177
        nas = np.concatenate(no_attn_scores, 1)
        no_attn_loss = no_attn_loss_weight * nas[mask].mean()
    else:
        no_attn_loss = 0
    if attention_divergence_loss_weight is not None:                   # :180-192
        fl = np.concatenate(flat, 1)                                   # [B_img, B_cap, S]
        kls = []
        for i in range(Bi):
            fa = fl[i]
            cur = np.broadcast_to(fa[i], fa.shape)
            kls.append(((kl_divergence(cur, fa) + kl_divergence(fa, cur)) / 2)[:, None])
        kls = np.concatenate(kls, 1)
        kl_loss = attention_divergence_loss_weight * (-kls[~mask].mean())
    else:
        kl_loss = 0
    if attention_entropy_loss_weight is not None:                      # :195-197 (weight NOT applied, as in ref)
        entropy_loss = np.concatenate(ents, 1).mean()
    else:
        entropy_loss = 0
    return loss0, loss1, no_attn_loss, kl_loss, entropy_loss, att_maps, similarities


# --------------------------------------------------------------------------------------
# local similarity: closed-form backward (checked vs reference autograd in tests)
# --------------------------------------------------------------------------------------
def local_sim_pair_bwd(ctx, w, temp1, temp2, g, agg="sum", d_attn_ext=None, eps=1e-8):
    """Backward of sim[:, i] (one caption against all images) and of an optional extra
    gradient `d_attn_ext [B, L, S]` flowing into the attention maps.

    ctx [B, D, S], w [D, L], g [B] = dLoss/dsim[:, i].
    Returns (d_ctx [B, D, S], d_w [D, L]).
    """
    blk = _pair_block(ctx, w, temp1)
    scores, p, a, c, dot, nw, nc = (blk[k] for k in ("scores", "p", "a", "c", "dot", "nw", "nc"))
    den = np.maximum(nw * nc, eps)
    r = dot / den
    if agg in ("sum", "mean"):
        q = _softmax(r * temp2, axis=1)
    elif agg == "max":
        q = np.zeros_like(r)
        q[np.arange(r.shape[0]), np.argmax(r, axis=1)] = 1.0
    else:
        raise ValueError(agg)
    dr = g[:, None] * temp2 * q                                   # [B, L]
    ddot = dr / den
    live = (nw * nc >= eps).astype(w.dtype)
    dden = -dr * dot / (den * den) * live
    dnc = dden * nw
    dnw = dden * nc
    with np.errstate(divide="ignore", invalid="ignore"):
        beta = np.where(nc > 0, dnc / nc, 0.0)                    # torch.norm backward: 0 at zero vector
        gamma = np.where(nw > 0, dnw / nw, 0.0)
    dC = ddot[:, None, :] * w[None] + beta[:, None, :] * c        # [B, D, L]
    dw = np.sum(ddot[:, None, :] * c + gamma[:, None, :] * w[None], axis=0)
    dA = np.matmul(np.swapaxes(dC, 1, 2), ctx)                   # [B, L, S]
    if d_attn_ext is not None:
        dA = dA + d_attn_ext
    dctx = np.matmul(dC, a)                                      # [B, D, S]
    dZ = a * (dA - np.sum(a * dA, axis=2, keepdims=True))         # softmax #2 backward (logits temp1*P)
    dP = temp1 * np.swapaxes(dZ, 1, 2)                            # [B, S, L]
    dS = p * (dP - np.sum(p * dP, axis=2, keepdims=True))         # softmax #1 backward
    dctx += np.matmul(w[None], np.swapaxes(dS, 1, 2))
    dw += np.sum(np.matmul(ctx, dS), axis=0)
    return dctx, dw


def local_loss_bwd(img_features, words_emb, cap_lens, temp1=4.0, temp2=5.0, temp3=10.0, agg="sum",
                   g0=1.0, g1=1.0, d_att_maps=None):
    """Gradient of g0*loss0 + g1*loss1 (+ sum_i <d_att_maps[i], att_maps[i]>) w.r.t. img_features, words_emb."""
    Bi, D = img_features.shape[:2]
    ctx = img_features.reshape(Bi, D, -1)
    sim = local_similarities(img_features, words_emb, cap_lens, temp1, temp2, agg)
    logits = sim * temp3
    dsim = temp3 * (g0 * _cross_entropy_arange_grad(logits) + g1 * _cross_entropy_arange_grad(logits.T).T)
    dctx = np.zeros_like(ctx)
    dwords = np.zeros_like(words_emb)
    for i in range(words_emb.shape[0]):
        L = int(cap_lens[i])
        ext = None
        if d_att_maps is not None and d_att_maps[i] is not None:
            ext = np.zeros((Bi, L, ctx.shape[2]), dtype=ctx.dtype)
            ext[i] = d_att_maps[i].reshape(L, -1)
        dc, dw = local_sim_pair_bwd(ctx, words_emb[i, :, :L], temp1, temp2, dsim[:, i], agg, ext)
        dctx += dc
        dwords[i, :, :L] = dw
    return dctx.reshape(img_features.shape), dwords


def local_loss_full_bwd(img_features, words_emb, cap_lens, temp1=4.0, temp2=5.0, temp3=10.0, agg="sum", g0=1.0, g1=1.0,
                        no_attn_vec=None, no_attn_loss_weight=None, attention_divergence_loss_weight=None,
                        attention_entropy_loss_weight=None):
    """Gradient of  g0*loss0 + g1*loss1 + no_attn_loss + kl_loss + entropy_loss  (the five scalars of
    gloria_loss.py:99-201, each as the reference returns it) w.r.t. img_features, words_emb and no_attn_vec.

    The regularisers (:129-139, 173-197) are functions of the attention maps only, so their gradient enters the closed
    form as the extra `d_attn_ext` of `local_sim_pair_bwd`; with `no_attn_vec` the context carries the learned column
    (:31-34) and its gradient is the sum of column 0 of d_ctx over the images.  Pinned to the reference's autograd on
    the `local_reg_*` / `local_ent_only_*` goldens (tests/test_oracle_golden.py).
    Returns (d_img, d_words, d_no_attn_vec or None).
    """
    Bi, D = img_features.shape[:2]
    Bc = words_emb.shape[0]
    dt = img_features.dtype
    has_v = no_attn_vec is not None
    ctx = img_features.reshape(Bi, D, -1)
    if has_v:
        ctx = np.concatenate([np.broadcast_to(no_attn_vec.reshape(1, D, 1), (Bi, D, 1)), ctx], axis=2)
    sim, attns = local_similarities(img_features, words_emb, cap_lens, temp1, temp2, agg, no_attn_vec=no_attn_vec,
                                    want_attn=True)
    logits = sim * temp3
    dsim = temp3 * (g0 * _cross_entropy_arange_grad(logits) + g1 * _cross_entropy_arange_grad(logits.T).T)
    s0 = 1 if has_v else 0
    S = ctx.shape[2] - s0
    # word-mean attention of every pair and what the regularisers see of it (:132-135)
    fret = np.stack([a[:, :, s0:].mean(1) for a in attns], 1)                  # [Bi, Bc, S]
    f = np.concatenate([1 - fret.sum(-1, keepdims=True), fret], -1) if has_v else fret
    df = np.zeros_like(f)                                                      # dLoss / d f
    if attention_entropy_loss_weight is not None:                              # mean over all pairs of -sum f log f
        df += -(np.log(f) + 1.0) / (Bi * Bc)
    if attention_divergence_loss_weight is not None:                           # -w * mean_{c != i} sym_kl(f[i,i], f[i,c])
        wkl = -attention_divergence_loss_weight / (Bi * (Bc - 1))
        for i in range(Bi):
            pcur = f[i, i][None]                                               # [1, K]
            q = f[i]                                                           # [Bc, K]
            lr = np.log(pcur) - np.log(q)
            dq = 0.5 * (-lr - (pcur - q) / q)
            dp = 0.5 * (lr + (pcur - q) / pcur)
            off = np.ones(Bc, dtype=bool)
            off[i] = False
            df[i, off] += wkl * dq[off]
            df[i, i] += wkl * dp[off].sum(0)
    dfret = (df[:, :, 1:] - df[:, :, :1]) if has_v else df                     # through f = [1 - sum, fret]
    dctx = np.zeros_like(ctx)
    dwords = np.zeros_like(words_emb)
    for i in range(Bc):
        L = int(cap_lens[i])
        ext = np.zeros((Bi, L, ctx.shape[2]), dtype=dt)
        ext[:, :, s0:] = dfret[:, i][:, None, :] / L                           # f = mean over the caption's words
        if no_attn_loss_weight is not None and i < Bi:                         # w * mean_i log(1 - mean_l sum_s a_ret) (:129-130,173-177)
            m = attns[i][i, :, s0:].sum(-1).mean()
            ext[i, :, s0:] += no_attn_loss_weight / Bi * (-1.0 / L) / (1.0 - m)
        dc, dw = local_sim_pair_bwd(ctx, words_emb[i, :, :L], temp1, temp2, dsim[:, i], agg, ext)
        dctx += dc
        dwords[i, :, :L] = dw
    d_nav = dctx[:, :, 0].sum(0) if has_v else None
    return dctx[:, :, s0:].reshape(img_features.shape), dwords, d_nav


# --------------------------------------------------------------------------------------
# GLoRIA model methods on the path
# --------------------------------------------------------------------------------------
def get_local_similarities(img_emb_l, text_emb_l, cap_lens, no_attn_vec=None):
    """gloria/models/gloria_model.py:171-207: words [1:L+1], temp1=4.0, temp2=5.0, max over words."""
    return local_similarities(img_emb_l, text_emb_l, cap_lens, 4.0, 5.0, "max",
                              no_attn_vec=no_attn_vec, word_offset=1)


def get_global_similarities(img_emb_g, text_emb_g):
    """gloria/models/gloria_model.py:164-169 -> sklearn.metrics.pairwise.cosine_similarity
    (scikit-learn 0.24.1 pinned in requirements.txt): rows L2-normalised (zero rows left as zero), then X @ Y.T."""
    def _normalize(x):
        n = np.sqrt(np.sum(x * x, axis=1, keepdims=True))
        n = np.where(n == 0, 1.0, n)
        return x / n
    return _normalize(img_emb_g) @ _normalize(text_emb_g).T


def segmentation_attention_loss(att_maps, segmentation_labels):
    """Supervised-attention term of GLoRIA.calc_loss, gloria_model.py:143-147 (weight not applied).

    att_maps: list of [1, L_i, h, w]; segmentation_labels [B, H, W] (bool / 0-1).
    F.interpolate default mode is 'nearest': src index = floor(dst * in / out).
    """
    B, H, W = segmentation_labels.shape
    mean_maps = np.concatenate([m.mean(1) for m in att_maps], 0)         # [B, h, w]
    h, w = mean_maps.shape[1:]
    iy = np.minimum((np.arange(H) * (h / H)).astype(np.int64), h - 1)
    ix = np.minimum((np.arange(W) * (w / W)).astype(np.int64), w - 1)
    up = mean_maps[:, iy][:, :, ix]                                      # [B, H, W]
    up = up / up.sum(-1, keepdims=True).sum(-2, keepdims=True)
    lab = segmentation_labels.astype(mean_maps.dtype)
    return np.mean(-np.log((lab * up).sum(-1).sum(-1)))


# ---------------------------------------------------------------------------------------------
# zero-shot driver (gloria/gloria.py:186-275)
# ---------------------------------------------------------------------------------------------
def get_similarities(img_emb_l, img_emb_g, text_emb_l, text_emb_g, cap_lens, similarity_type="both"):
    """gloria.py:224-237 on embeddings: local (gloria_model.py:171-207), global (:164-169) or their mean."""
    g = get_global_similarities(img_emb_g, text_emb_g)
    loc = get_local_similarities(img_emb_l, text_emb_l, cap_lens)
    if similarity_type == "global":
        return g
    if similarity_type == "local":
        return loc
    if similarity_type == "both":
        return (loc + g) / 2
    raise RuntimeError("similarity type should be one of ['global', 'local', 'both']")


def zero_shot_classification(img_emb_l, img_emb_g, text_emb_l, text_emb_g, cap_lens, class_sizes):
    """gloria.py:240-275: per class, "both" similarities of every image with the class prompts, max over the prompts
    (:262), stacked to [N_img, n_classes]; z-scored over the images (utils.py:12-15, numpy std, ddof = 0) when there is
    more than one image (:268).  Prompts of class k are rows sum(class_sizes[:k]) ... of the text embeddings."""
    cols, o = [], 0
    for n in class_sizes:
        n = int(n)
        s = get_similarities(img_emb_l, img_emb_g, text_emb_l[o:o + n], text_emb_g[o:o + n], list(cap_lens[o:o + n]))
        cols.append(s.max(axis=1))
        o += n
    cs = np.stack(cols, axis=1)
    if cs.shape[0] > 1:
        cs = (cs - cs.mean(axis=0)) / cs.std(axis=0)
    return cs


# ---------------------------------------------------------------------------------------------
# word-piece aggregation (gloria/models/text_model.py:32-90)
# ---------------------------------------------------------------------------------------------
def aggregate_tokens(embeddings, caption_ids, idxtoword):
    """BertEncoder.aggregate_tokens: embeddings [B, layers, T, D], caption_ids [B, T] ->
    (agg [B, layers, T, D], sentences).  Word pieces ("##...") are summed into the word they continue; "[SEP]" closes
    the open word, is kept as a word of its own and ends the caption (:50-58); without "[SEP]" the word still open at
    the end is never emitted; rows beyond the caption's words are zero, sentences are padded with "[PAD]" (:77-82)."""
    B, layers, T, D = embeddings.shape
    out = np.zeros_like(embeddings)
    sentences = []
    for b in range(B):
        words, bank, bank_words, n = [], [], [], 0
        for t in range(T):
            word = idxtoword[int(caption_ids[b, t])]
            if word == "[SEP]":
                out[b, :, n] = np.sum([embeddings[b, :, k] for k in bank], axis=0) if bank else 0
                words.append("".join(bank_words))
                n += 1
                out[b, :, n] = embeddings[b, :, t]
                words.append(word)
                n += 1
                break
            if not word.startswith("##"):
                if len(bank_words) == 0:
                    bank, bank_words = [t], [word]
                else:
                    out[b, :, n] = np.sum([embeddings[b, :, k] for k in bank], axis=0)
                    words.append("".join(bank_words))
                    n += 1
                    bank, bank_words = [t], [word]
            else:
                bank.append(t)
                bank_words.append(word[2:])
        sentences.append(words + ["[PAD]"] * (T - n))
    return out, sentences
