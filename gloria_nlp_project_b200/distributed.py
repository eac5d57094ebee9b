"""Caption-sharded GLoRIA loss across the GPUs of one node.

The reference has no multi-device path of its own (Lightning 'dp' computes the loss on each replica's sub-batch,
SURVEY.md section 2a); this module gives the *full-batch* loss -- identical to the single-device reference on the
concatenated batch -- with the B_img x B_cap pair grid split by caption columns:

  rank r owns images and captions [r*B/G, (r+1)*B/G)            (its data-parallel shard)
  forward : all_gather(region features), all_gather(global image features)         [NCCL over NVLink]
            sim[:, own captions]  <- fused local-similarity kernels                [no collective inside]
            all_gather(logit column blocks)  -> both cross entropies on every rank (B x B floats)
  backward: d(sim block) = own columns of d(logits);  d(words), d(text global) stay local;
            d(region features), d(global image features) are partial sums over the rank's captions for ALL images
            -> reduce_scatter(sum) back to the owners.

Every rank returns the same loss value (the global-batch loss) and the gradient of that loss w.r.t. its own inputs.
The block-similarity functions are injectable so that the collective logic is testable with gloo on CPU
(tests/test_distributed_cpu.py plugs the oracle in); the defaults are the CUDA ops and fail on CPU tensors.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch
import torch.distributed as dist

__all__ = ["sharded_loss", "gather_cat", "gather_reduce_scatter"]


class _GatherSliceGrad(torch.autograd.Function):
    """all_gather along dim 0; backward keeps this rank's slice of the incoming gradient (for tensors whose
    consumer is computed identically on every rank, so the incoming gradient is already the full one)."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        ctx.n = x.shape[0]
        world = dist.get_world_size(group)
        out = x.new_empty((world * x.shape[0],) + tuple(x.shape[1:]))
        dist.all_gather_into_tensor(out, x.contiguous(), group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        r = dist.get_rank(ctx.group)
        return g[r * ctx.n:(r + 1) * ctx.n].contiguous(), None


class _GatherReduceScatterGrad(torch.autograd.Function):
    """all_gather along dim 0; backward reduce_scatters (sum) the incoming gradient (for tensors whose consumer on
    each rank produces only that rank's partial gradient for every shard)."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        world = dist.get_world_size(group)
        out = x.new_empty((world * x.shape[0],) + tuple(x.shape[1:]))
        dist.all_gather_into_tensor(out, x.contiguous(), group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        world = dist.get_world_size(ctx.group)
        g = g.contiguous()
        out = g.new_empty((g.shape[0] // world,) + tuple(g.shape[1:]))
        dist.reduce_scatter_tensor(out, g, op=dist.ReduceOp.SUM, group=ctx.group)
        return out, None


def gather_cat(x: torch.Tensor, group=None) -> torch.Tensor:
    return _GatherSliceGrad.apply(x, group)


def gather_reduce_scatter(x: torch.Tensor, group=None) -> torch.Tensor:
    return _GatherReduceScatterGrad.apply(x, group)


_SIDE_STREAMS: "dict[int, torch.cuda.Stream]" = {}


def _side_stream(dev: torch.device) -> "torch.cuda.Stream":
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _SIDE_STREAMS:
        _SIDE_STREAMS[idx] = torch.cuda.Stream(device=idx)
    return _SIDE_STREAMS[idx]


class _ShardedLocalSim(torch.autograd.Function):
    """sim[:, own captions] from this rank's shard of images and captions, with the collectives placed around the
    kernels: forward all_gathers the region features and runs the fused training kernel on [B, B/G] pairs; backward
    computes the image-side gradient first, and its reduce_scatter (on a side stream, released by an event the
    library records once d_ctx is final) overlaps the caption-side GEMM of the same backward."""

    @staticmethod
    def forward(ctx, img_emb_l, text_emb_l, dev_lens, lcap, temp1, temp2, agg, eps, mode, group):
        from . import ops
        world = dist.get_world_size(group)
        n, D = img_emb_l.shape[0], img_emb_l.shape[1]
        x = img_emb_l.reshape(n, D, -1).float().contiguous()
        img_all = x.new_empty((world * n, D, x.shape[2]))
        dist.all_gather_into_tensor(img_all, x, group=group)
        words = text_emb_l.float().contiguous()
        if any(ctx.needs_input_grad[:2]):
            # tells the op (called here with autograd off) to run the fused training forward and keep its state
            words = words.detach().requires_grad_(True)
        sim, _, _, stats = ops.local_sim_fwd(img_all, words, dev_lens, lcap, 0, temp1, temp2, agg, eps, False, False,
                                             mode)
        ctx.save_for_backward(img_all, words.detach(), dev_lens, stats)
        ctx.args = (lcap, temp1, temp2, agg, eps, mode, group, img_emb_l.shape, img_emb_l.dtype, text_emb_l.dtype)
        return sim

    @staticmethod
    def backward(ctx, dsim):
        from . import ops
        img_all, words, dev_lens, stats = ctx.saved_tensors
        lcap, temp1, temp2, agg, eps, mode, group, img_shape, img_dtype, txt_dtype = ctx.args
        world = dist.get_world_size(group)
        dev = img_all.device
        ready = torch.cuda.Event()
        ready.record()                                   # materialise the handle; the op re-records it
        d_ctx_all, d_words = ops.local_sim_bwd(img_all, words, dev_lens, lcap, 0, temp1, temp2, agg, eps,
                                               dsim.contiguous(), None, None, stats if stats.numel() > 0 else None,
                                               mode, ready.cuda_event)
        d_ctx = d_ctx_all.new_empty((d_ctx_all.shape[0] // world,) + tuple(d_ctx_all.shape[1:]))
        main, side = torch.cuda.current_stream(dev), _side_stream(dev)
        side.wait_event(ready)                           # d_ctx_all is final here; the caption-side GEMM still runs
        with torch.cuda.stream(side):
            dist.reduce_scatter_tensor(d_ctx, d_ctx_all, op=dist.ReduceOp.SUM, group=group)
        main.wait_stream(side)
        return (d_ctx.reshape(img_shape).to(img_dtype), d_words.to(txt_dtype), None, None, None, None, None, None,
                None, None)


def _sharded_local_sim(img_emb_l, text_emb_l, cap_lens, temp1, temp2, agg, group):
    """Default CUDA path of the sharded local similarity -> sim[:, own captions] ([B, B/G])."""
    from . import gloria_loss, ops
    if not img_emb_l.is_cuda:
        raise RuntimeError("gloria_b200 runs on CUDA tensors only (sm_100a); there is no CPU fallback")
    Bc, _, Lw = text_emb_l.shape
    dev_lens, lens = gloria_loss._cap_lens(cap_lens, Bc, 0, Lw, img_emb_l.device)
    mode = gloria_loss._mode(img_emb_l, text_emb_l)
    return _ShardedLocalSim.apply(img_emb_l, text_emb_l, dev_lens, max(lens), float(temp1), float(temp2),
                                  ops.AGG[agg], 1e-8, mode, group)


def _default_local_sim(img_all, words_local, cap_lens_local, temp1, temp2, agg):
    from . import gloria_loss
    sim, _, _, _ = gloria_loss.local_similarities(img_all, words_local, cap_lens_local, temp1, temp2, agg)
    return sim                                                       # [B_all, B_local]


def _default_global_cos(img_g_all, txt_g_local, eps):
    from . import ops
    cosm, _, _ = ops.global_sim_fwd(img_g_all.float(), txt_g_local.float(), float(eps))
    return cosm                                                      # [B_all, B_local]


def _default_ce(m, scale):
    from . import ops
    losses, _, _ = ops.ce_bidir_fwd(m, float(scale))
    return losses[0], losses[1]


def sharded_loss(img_emb_l: torch.Tensor, text_emb_l: torch.Tensor, img_emb_g: torch.Tensor,
                 text_emb_g: torch.Tensor, cap_lens: Sequence[int], temp1: float = 4.0, temp2: float = 5.0,
                 temp3: float = 10.0, agg: str = "sum", eps: float = 1e-8, group=None,
                 local_sim_fn: Optional[Callable] = None, global_cos_fn: Optional[Callable] = None,
                 ce_fn: Optional[Callable] = None):
    """Full-batch local + global GLoRIA losses from per-rank shards -> (l_loss0, l_loss1, g_loss0, g_loss1).

    All ranks must hold the same number of pairs.  Equals `local_loss` / `global_loss` of the reference
    (gloria_loss.py:99-170, 66-88) evaluated on the concatenation of all ranks' shards.
    """
    local_sim_fn = local_sim_fn or _default_local_sim
    global_cos_fn = global_cos_fn or _default_global_cos
    ce_fn = ce_fn or _default_ce
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        sim = local_sim_fn(img_emb_l, text_emb_l, cap_lens, temp1, temp2, agg)
        cosm = global_cos_fn(img_emb_g, text_emb_g, eps)
        return (*ce_fn(sim, temp3), *ce_fn(cosm, temp3))
    img_g_all = gather_reduce_scatter(img_emb_g, group)              # [B, D]
    if local_sim_fn is _default_local_sim:
        sim_blk = _sharded_local_sim(img_emb_l, text_emb_l, cap_lens, temp1, temp2, agg, group)   # [B, B/G]
    else:
        img_all = gather_reduce_scatter(img_emb_l, group)            # [B, D, H, W]
        sim_blk = local_sim_fn(img_all, text_emb_l, cap_lens, temp1, temp2, agg)      # [B, B/G]
    cos_blk = global_cos_fn(img_g_all, text_emb_g, eps)                                # [B, B/G]
    # one small collective for both logit blocks: rows = captions of every rank after the transpose
    both = torch.stack([sim_blk.t(), cos_blk.t()], 1)                # [B/G, 2, B]
    both_all = gather_cat(both, group)                               # [B, 2, B]   (caption-major)
    sim = both_all[:, 0].t()                                         # [B_img, B_cap]
    cosm = both_all[:, 1].t()
    return (*ce_fn(sim.contiguous(), temp3), *ce_fn(cosm.contiguous(), temp3))
