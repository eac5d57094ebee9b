"""Drop-in replacement for the reference module gloria/loss/gloria_loss.py.

Same function names, positional/keyword signatures, defaults and return arities as the reference
(cosine_similarity :11, attention_fn :19, global_loss :66, kl_divergence :91, entropy :95, local_loss :99);
the arithmetic runs in the sm_100a kernels of libgloria_b200.so through the custom ops in `ops.py`.
CUDA tensors only -- there is no CPU fallback.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple, Union

import torch

from . import _config, ops

__all__ = ["cosine_similarity", "attention_fn", "global_loss", "kl_divergence", "entropy", "local_loss",
           "local_similarities", "diagonal_attention_maps", "supervised_attention_loss", "plan_length_buckets",
           "DeviceCapLens", "LazyAttnMaps"]


def _mode(*tensors: torch.Tensor) -> int:
    p = _config.get_precision()
    if p == "auto":
        low = any(t is not None and t.dtype in (torch.float16, torch.bfloat16) for t in tensors)
        p = "bf16" if (low or torch.is_autocast_enabled()) else "fp32"
    return ops.MODE_BF16 if p == "bf16" else ops.MODE_FP32


class DeviceCapLens:
    """Caption lengths that exist on the device only (e.g. from `text_model.aggregate_tokens`, which derives them from
    the token ids in its word-boundary kernel).  Passing one as `cap_lens` keeps the whole loss step free of host
    round trips: no `.tolist()`, no H2D copy, one launch padded to the word axis of `words_emb` (the kernels clamp
    each caption's length to [0, Lw - word_offset] themselves; no length buckets), and the attention maps are sliced
    per caption only when somebody reads them (`LazyAttnMaps`).  A plain CUDA int tensor is NOT treated this way: the
    reference reads `cap_lens[i]` as Python ints, and a tensor argument keeps that meaning."""

    def __init__(self, lens: torch.Tensor):
        self.tensor = lens.reshape(-1).to(torch.int32).contiguous()
        self._host = None

    def __len__(self):
        return self.tensor.numel()

    def tolist(self) -> List[int]:
        if self._host is None:
            self._host = [int(v) for v in self.tensor.tolist()]          # the one sync, only on demand
        return self._host


class LazyAttnMaps(Sequence):
    """`att_maps` of local_loss (gloria_loss.py:141-143: a list of [1, L_i, H, W] maps of the diagonal pairs) over the
    padded [B, Lcap, S] tensor the kernel wrote; element i is sliced to its caption's length when it is read (the
    lengths are then fetched from the device once).  `stacked` gives the padded tensor without any sync."""

    def __init__(self, diag: torch.Tensor, lens: "DeviceCapLens", s0: int, ih: int, iw: int):
        self.stacked, self._lens, self._s0, self._hw = diag, lens, s0, (ih, iw)

    def __len__(self):
        return self.stacked.shape[0]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        L = min(self._lens.tolist()[i], self.stacked.shape[1])
        return self.stacked[i:i + 1, :L, self._s0:].reshape(1, L, *self._hw)


class AttnMapList(list):
    """The reference's plain list of [1, L_i, H, W] maps (host-side caption lengths), which also remembers the padded
    [B, Lcap, S] tensor it views and the lengths on the device: consumers that reduce over all maps
    (`supervised_attention_loss`) then run a handful of batched launches instead of B small ones."""
    stacked = None
    lens_tensor = None
    _s0 = 0
    _hw = (0, 0)


def _att_maps(diag, lens, dev_lens_obj, s0, ih, iw, n, dev_lens=None):
    if lens is None:
        return LazyAttnMaps(diag, dev_lens_obj, s0, ih, iw)
    out = AttnMapList(diag[i:i + 1, :lens[i], s0:].reshape(1, lens[i], ih, iw) for i in range(n))
    if dev_lens is None:
        dev_lens = ops.upload_ints(lens[:n], diag.device)
    out.stacked, out.lens_tensor, out._s0, out._hw = diag, dev_lens, s0, (ih, iw)
    return out


def _cap_lens(cap_lens: Union[Sequence[int], torch.Tensor, "DeviceCapLens"], n: int, word_off: int, Lw: int,
              device: torch.device) -> Tuple[torch.Tensor, Optional[List[int]]]:
    """The reference indexes cap_lens[i] as Python ints (list, or tensor -> implicit sync); do the same once.
    A `DeviceCapLens` skips the host side altogether: returns (device int32 tensor, None)."""
    if isinstance(cap_lens, DeviceCapLens):
        if len(cap_lens) < n:
            raise RuntimeError(f"cap_lens has {len(cap_lens)} entries for {n} captions")
        t = cap_lens.tensor if cap_lens.tensor.device == device else cap_lens.tensor.to(device)
        return (t if len(cap_lens) == n else t[:n].contiguous()), None
    if isinstance(cap_lens, torch.Tensor):
        lens = [int(v) for v in cap_lens.reshape(-1).tolist()]
    else:
        lens = [int(v) for v in cap_lens]
    if len(lens) < n:
        raise RuntimeError(f"cap_lens has {len(lens)} entries for {n} captions")
    lens = lens[:n]
    if min(lens) < 1 or max(lens) + word_off > Lw:
        raise RuntimeError(f"cap_lens out of range: need 1 <= len and len + {word_off} <= {Lw}, got "
                           f"[{min(lens)}, {max(lens)}]")
    return ops.upload_ints(lens, device), lens


def _context(img_features: torch.Tensor, no_attn_vec: Optional[torch.Tensor]) -> torch.Tensor:
    """[B, D, H, W] -> [B, D, S] fp32, with the learned no-attention column prepended (gloria_loss.py:30-34)."""
    B, D = img_features.shape[0], img_features.shape[1]
    ctx = img_features.reshape(B, D, -1).float()
    if no_attn_vec is not None:
        v = no_attn_vec.float().expand(B, no_attn_vec.shape[0]).unsqueeze(-1)
        ctx = torch.cat([v, ctx], 2)
    return ctx


# Length buckets (bf16 tensor-core mode).  The kernels are compiled per padded caption length (multiples of 16 words)
# and a launch pads every caption to its longest one, so a batch with cap_lens ~ U{5..97} does 112 / 58 of the
# necessary work in one launch.  Captions are therefore grouped by padded length and each group gets its own launch
# (and its own fused-training state); groups are merged while that costs less than a launch's fixed overhead.
BUCKET_OVERHEAD_US = 300.0        # host + launch cost of one more group, independent of the batch
BUCKET_OVERHEAD_US_PER_IMAGE = 2.4    # prepack + gradient accumulation of one more group, per image
BUCKET_US_PER_IMAGE_WORD = 3.4e-3     # kernel time per image and padded caption word (98.7 ms / 512 / (512 * 112))


def plan_length_buckets(lens: Sequence[int], n_images: int, word_offset: int = 0) -> List[Tuple[List[int], int]]:
    """Group caption indices by padded length -> [(indices, lcap)], longest group first; one group = no bucketing.
    Pure host logic (tested on CPU)."""
    def lpad(n):
        return (n + 15) // 16 * 16
    groups = {}
    for i, n in enumerate(lens):
        groups.setdefault(lpad(int(n)), []).append(i)
    keys = sorted(groups, reverse=True)                       # padded lengths, descending
    merge_cost_words = (BUCKET_OVERHEAD_US + BUCKET_OVERHEAD_US_PER_IMAGE * n_images) / \
        (BUCKET_US_PER_IMAGE_WORD * max(n_images, 1))
    # greedy: repeatedly fold the group whose padding to its longer neighbour wastes the fewest caption-words
    while len(keys) > 1:
        waste = [(len(groups[keys[k + 1]]) * (keys[k] - keys[k + 1]), k) for k in range(len(keys) - 1)]
        w, k = min(waste)
        if w >= merge_cost_words:
            break
        groups[keys[k]] = groups[keys[k]] + groups.pop(keys[k + 1])
        keys.pop(k + 1)
    return [(sorted(groups[k]), max(int(lens[i]) for i in groups[k])) for k in keys]


def local_similarities(img_features, words_emb, cap_lens, temp1=4.0, temp2=5.0, agg="sum", no_attn_vec=None,
                       word_offset=0, eps=1e-8, want_attn_maps=False, want_mean_attn=False):
    """The [B_img, B_cap] matrix the caption loop of gloria_loss.py:116-162 builds (before the temp3 scale).

    Returns (sim, attn_diag or None, attn_mean or None, lens):
      attn_diag [B_cap, Lcap, S(+1)]  attention of the diagonal pairs (zero rows beyond each caption's length),
      attn_mean [B_img, B_cap, S(+1)] word-mean attention of every pair.
    """
    if not img_features.is_cuda:
        raise RuntimeError("gloria_b200 runs on CUDA tensors only (sm_100a); there is no CPU fallback")
    Bc, _, Lw = words_emb.shape
    dev_lens, lens = _cap_lens(cap_lens, Bc, word_offset, Lw, img_features.device)
    ctx = _context(img_features, no_attn_vec)
    words = words_emb.float()
    mode = _mode(img_features, words_emb)
    need_grad = torch.is_grad_enabled() and (img_features.requires_grad or words_emb.requires_grad
                                             or (no_attn_vec is not None and no_attn_vec.requires_grad))
    lcap = max(lens) if lens is not None else Lw - word_offset       # device-only lengths: pad to the word axis
    buckets = plan_length_buckets(lens, ctx.shape[0], word_offset) \
        if (mode == ops.MODE_BF16 and not want_mean_attn and lens is not None) else []
    if len(buckets) > 1:
        parts, order = [], []
        with ops.shared_ctx_pack(ctx):                               # the images are packed once for all groups
            for idx, lcap_b in buckets:
                sel = ops.upload_ints(idx, ctx.device)               # (through the kernel parameter buffer: no H2D copy)
                sim_b, _, _, _ = ops.local_sim_fwd(ctx, words.index_select(0, sel), dev_lens.index_select(0, sel),
                                                   lcap_b, word_offset, float(temp1), float(temp2), ops.AGG[agg],
                                                   float(eps), False, False, mode, need_grad)
                parts.append(sim_b)
                order += idx
        inv = [0] * Bc
        for k, i in enumerate(order):
            inv[i] = k
        sim = torch.cat(parts, 1).index_select(1, ops.upload_ints(inv, ctx.device))
        diag = None
        if want_attn_maps:                                           # B diagonal pairs through the exact fp32 kernels
            diag = ops.diag_attn_fwd(ctx, words, dev_lens, max(lens), word_offset, float(temp1))
        return sim, diag, None, lens
    sim, diag, mean, _ = ops.local_sim_fwd(ctx, words, dev_lens, lcap, word_offset, float(temp1), float(temp2),
                                           ops.AGG[agg], float(eps), bool(want_attn_maps), bool(want_mean_attn),
                                           mode, need_grad)
    return sim, (diag if want_attn_maps else None), (mean if want_mean_attn else None), lens


def diagonal_attention_maps(img_features, words_emb, cap_lens, temp1=4.0, no_attn_vec=None, word_offset=0):
    """att_maps of the diagonal pairs only -- the list local_loss returns as its 6th item (gloria_loss.py:141-143,
    `attn[i:i+1]` of caption i), computed for B pairs instead of B^2.  Differentiable."""
    if not img_features.is_cuda:
        raise RuntimeError("gloria_b200 runs on CUDA tensors only (sm_100a); there is no CPU fallback")
    Bc, _, Lw = words_emb.shape
    ih, iw = img_features.shape[2], img_features.shape[3]
    dev_lens, lens = _cap_lens(cap_lens, Bc, word_offset, Lw, img_features.device)
    ctx = _context(img_features, no_attn_vec)
    lcap = max(lens) if lens is not None else Lw - word_offset
    diag = ops.diag_attn_fwd(ctx, words_emb.float(), dev_lens, lcap, word_offset, float(temp1))
    s0 = 1 if no_attn_vec is not None else 0
    return _att_maps(diag, lens, cap_lens, s0, ih, iw, Bc, dev_lens)


_CELL_INDEX = {}


def _cell_index(ih, iw, oh, ow, device):
    """Source cell of every output pixel under nn.functional.interpolate's default (nearest) rule, taken from
    interpolate itself so that the index arithmetic is identical by construction."""
    key = (ih, iw, oh, ow, str(device))
    if key not in _CELL_INDEX:
        cells = torch.arange(ih * iw, dtype=torch.float32, device=device).view(1, 1, ih, iw)
        _CELL_INDEX[key] = torch.nn.functional.interpolate(cells, size=(oh, ow)).view(-1).long()
    return _CELL_INDEX[key]


def supervised_attention_loss(att_maps, segmentation_labels):
    """The supervised-attention term of GLoRIA.calc_loss (gloria_model.py:143-147) without the upsampled maps:
    nearest upsampling repeats each grid cell over a fixed set of pixels, so  sum(label * up) / sum(up)  is
    sum_c map[c] * (#labelled pixels of cell c) / sum_c map[c] * (#pixels of cell c).  Returns the batch mean of
    -log of that ratio (the caller applies segmentation_loss_weight)."""
    lens_t = att_maps._lens.tensor if isinstance(att_maps, LazyAttnMaps) else getattr(att_maps, "lens_tensor", None)
    if lens_t is not None and getattr(att_maps, "stacked", None) is not None:
        # word-mean of each map from the padded tensor (rows beyond a caption's length are zero): sum / length, no sync
        ih, iw = att_maps._hw
        d = att_maps.stacked[:, :, att_maps._s0:]
        n = lens_t[:d.shape[0]].clamp(1, d.shape[1]).to(d.dtype)
        mean_maps = (d.sum(1) / n[:, None]).reshape(d.shape[0], ih, iw)
    else:
        mean_maps = torch.cat([m.mean(1) for m in att_maps], 0)              # [B, h, w]  (:144)
    B, ih, iw = mean_maps.shape
    oh, ow = segmentation_labels.shape[1:]
    cell = _cell_index(ih, iw, oh, ow, mean_maps.device)
    lab = segmentation_labels.reshape(B, -1).to(mean_maps.dtype)
    inside = torch.zeros((B, ih * iw), dtype=mean_maps.dtype, device=mean_maps.device).index_add_(1, cell, lab)
    total = torch.bincount(cell, minlength=ih * iw).to(mean_maps.dtype)
    flat = mean_maps.reshape(B, -1)
    return -torch.log((flat * inside).sum(-1) / (flat * total).sum(-1)).mean()


# --------------------------------------------------------------------------------------------------------------
def cosine_similarity(x1, x2, dim=1, eps=1e-8):
    """gloria_loss.py:11-16 -- cosine along `dim` with the product of norms clamped at eps, then .squeeze()."""
    if not x1.is_cuda:
        raise RuntimeError("gloria_b200 runs on CUDA tensors only (sm_100a); there is no CPU fallback")
    if x1.shape != x2.shape:
        x1, x2 = torch.broadcast_tensors(x1, x2)
    a = x1.movedim(dim, -1)
    b = x2.movedim(dim, -1)
    lead = a.shape[:-1]
    out, _ = ops.row_cosine_fwd(a.reshape(-1, a.shape[-1]).float(), b.reshape(-1, b.shape[-1]).float(), float(eps))
    return out.reshape(lead).squeeze()


def attention_fn(query, context, temp1, no_attn_vec=None):
    """gloria_loss.py:19-63: query [B, D, L], context [B, D, H, W] -> (weightedContext [B, D, L], attn [B, L, H, W]).
    With no_attn_vec [D] a learned column is prepended to the context (:31-34) and stripped from the returned
    attention only (:60-61)."""
    if not query.is_cuda:
        raise RuntimeError("gloria_b200 runs on CUDA tensors only (sm_100a); there is no CPU fallback")
    B, ih, iw = context.shape[0], context.shape[2], context.shape[3]
    ctx = _context(context, no_attn_vec)
    wctx, attn = ops.attention_fwd(query.float(), ctx, float(temp1))
    if no_attn_vec is not None:
        attn = attn[:, :, 1:]
    return wctx, attn.reshape(B, query.shape[2], ih, iw)


def global_loss(cnn_code, rnn_code, eps=1e-8, temp3=10.0):
    """gloria_loss.py:66-88 -> (loss0, loss1)."""
    cosm, _, _ = ops.global_sim_fwd(cnn_code.float(), rnn_code.float(), float(eps))
    losses, _, _ = ops.ce_bidir_fwd(cosm, float(temp3))
    return losses[0], losses[1]


def kl_divergence(attn1, attn2):
    """gloria_loss.py:91-92"""
    return (attn1 * torch.log(attn1 / attn2)).sum(-1)


def entropy(attn):
    """gloria_loss.py:95-96"""
    return -(attn * torch.log(attn)).sum(-1)


def local_loss(
    img_features, words_emb, cap_lens, temp1=4.0, temp2=5.0, temp3=10.0, agg="sum", no_attn_vec=None,
    no_attn_loss_weight=None, attention_divergence_loss_weight=None, attention_entropy_loss_weight=None
):
    """gloria_loss.py:99-201 -> (loss0, loss1, no_attn_loss, kl_loss, entropy_loss, att_maps).

    att_maps[i] is the [1, L_i, H, W] attention map of pair (i, i), differentiable (the supervised-attention
    loss of GLoRIA.calc_loss back-propagates through it).
    """
    batch_size = img_features.shape[0]
    ih, iw = img_features.shape[2], img_features.shape[3]
    has_v = no_attn_vec is not None
    want_mean = (no_attn_loss_weight is not None or attention_divergence_loss_weight is not None
                 or attention_entropy_loss_weight is not None)
    sim, diag, mean, lens = local_similarities(img_features, words_emb, cap_lens, temp1, temp2, agg, no_attn_vec,
                                               word_offset=0, want_attn_maps=True, want_mean_attn=want_mean)
    s0 = 1 if has_v else 0                                          # returned maps drop the no-attn column (:60-61)
    att_maps = _att_maps(diag, lens, cap_lens, s0, ih, iw, words_emb.shape[0])

    losses, _, _ = ops.ce_bidir_fwd(sim, float(temp3))              # :164-170
    loss0, loss1 = losses[0], losses[1]

    no_attn_loss, kl_loss, entropy_loss = 0, 0, 0
    if want_mean:
        flat = mean[:, :, s0:]                                      # [B_img, B_cap, S]  (:132)
        ar = torch.arange(batch_size, device=sim.device)
        if no_attn_loss_weight is not None:                         # :129-130, 173-177 (diagonal pairs only)
            no_attn_scores = torch.log(1 - flat[ar, ar].sum(-1))
            no_attn_loss = no_attn_loss_weight * no_attn_scores.mean()
        if has_v:                                                   # :133-135
            flat = torch.cat([1 - flat.sum(-1, keepdim=True), flat], -1)
        if attention_entropy_loss_weight is not None:               # :136-137, 195-197 (weight not applied in ref)
            entropy_loss = entropy(flat).mean()
        if attention_divergence_loss_weight is not None:            # :138-139, 180-192
            cur = flat[ar, ar].unsqueeze(1)                         # attention of pair (i, i)
            sym = (kl_divergence(cur, flat) + kl_divergence(flat, cur)) / 2      # [B_img, B_cap]
            off_diag = ~torch.eye(batch_size, dtype=torch.bool, device=sim.device)
            kl_loss = attention_divergence_loss_weight * (-sym[off_diag].mean())
    return loss0, loss1, no_attn_loss, kl_loss, entropy_loss, att_maps
