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
    img_all = gather_reduce_scatter(img_emb_l, group)                # [B, D, H, W]
    img_g_all = gather_reduce_scatter(img_emb_g, group)              # [B, D]
    sim_blk = local_sim_fn(img_all, text_emb_l, cap_lens, temp1, temp2, agg)          # [B, B/G]
    cos_blk = global_cos_fn(img_g_all, text_emb_g, eps)                                # [B, B/G]
    # one small collective for both logit blocks: rows = captions of every rank after the transpose
    both = torch.stack([sim_blk.t(), cos_blk.t()], 1)                # [B/G, 2, B]
    both_all = gather_cat(both, group)                               # [B, 2, B]   (caption-major)
    sim = both_all[:, 0].t()                                         # [B_img, B_cap]
    cosm = both_all[:, 1].t()
    return (*ce_fn(sim.contiguous(), temp3), *ce_fn(cosm.contiguous(), temp3))
