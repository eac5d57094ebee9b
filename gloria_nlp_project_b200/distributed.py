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

import os

import torch
import torch.distributed as dist

__all__ = ["sharded_loss", "sharded_calc_loss", "gather_cat", "gather_reduce_scatter", "part_major_rows"]


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
        sim, _, _, stats = ops.local_sim_fwd(img_all, words, dev_lens, lcap, 0, temp1, temp2, agg, eps, False, False,
                                             mode, bool(any(ctx.needs_input_grad[:2])))
        ctx.save_for_backward(img_all, words, dev_lens, stats)
        ctx.args = (lcap, temp1, temp2, agg, eps, mode, group, img_emb_l.shape, img_emb_l.dtype, text_emb_l.dtype)
        return sim

    @staticmethod
    def backward(ctx, dsim):
        from . import ops
        img_all, words, dev_lens, stats = ctx.saved_tensors
        lcap, temp1, temp2, agg, eps, mode, group, img_shape, img_dtype, txt_dtype = ctx.args
        world = dist.get_world_size(group)
        dev = img_all.device
        ops.check_state_unconsumed(ctx, img_all, words, lcap, stats, mode)
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


_PACKED_RS = os.environ.get("GLORIA_B200_SHARD_PACKED_RS", "1") != "0"   # reduce_scatter the packed gradient rows, unpack own images only
_N_PARTS = int(os.environ.get("GLORIA_B200_SHARD_PARTS", "2"))   # image parts per rank shard: gather / reduce_scatter of one part overlap the kernels of the other


def part_major_rows(n_per_rank: int, world: int, parts: int) -> torch.Tensor:
    """Row of every image in the part-major gathered batch: image g = r * n + p * m + k (rank r, part p of `parts`,
    k < m = n / parts) is gathered to row  p * (world * m) + r * m + k,  because part p of every rank lands in one
    contiguous block of the all_gather output.  Returns the int64 index [world * n] (CPU)."""
    m = n_per_rank // parts
    g = torch.arange(world * n_per_rank)
    r, q = g // n_per_rank, g % n_per_rank
    return (q // m) * (world * m) + r * m + (q % m)


_PERM_CACHE: "dict[tuple, torch.Tensor]" = {}


def _part_major_rows_on(dev: torch.device, n_per_rank: int, world: int, parts: int) -> torch.Tensor:
    """`part_major_rows` on the device, built once per configuration: a per-step pageable H2D copy of it blocks the host
    and queues behind the next batch's input prefetch on the copy engine (bench.py's e2e loop at 8 ranks)."""
    key = (str(dev), n_per_rank, world, parts)
    if key not in _PERM_CACHE:
        _PERM_CACHE[key] = part_major_rows(n_per_rank, world, parts).to(dev)
    return _PERM_CACHE[key]


class _ShardedLocalSimParts(torch.autograd.Function):
    """bf16 training path of the sharded local similarity, pipelined over image parts.

    Each rank's n images are cut into P parts; part p of every rank is all_gathered into one contiguous block, so the
    gathered batch is *part-major* ([p][rank][n/P]; a fixed permutation of the natural image order that is undone on
    the rows of sim).  Forward: each rank packs its own images and the 16-bit copies are gathered on a side stream
    while the fused training kernel already runs on the rank's own images (from the local pack); the other ranks' images
    follow part by part as their gathers land.  Backward: the library finishes d_ctx part by part and records an event per part; the
    reduce_scatter of part p (side stream) overlaps the GEMMs of part p+1 and the caption-side GEMM.
    Nothing but the 16-bit packed copies and the operand workspace is kept for the backward."""

    @staticmethod
    def forward(ctx, img_emb_l, text_emb_l, dev_lens, lcap, temp1, temp2, agg, eps, group, P):
        from . import _lib
        from .ops import _stream
        L = _lib.lib()
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        n, D = img_emb_l.shape[0], img_emb_l.shape[1]
        x = img_emb_l.reshape(n, D, -1).float().contiguous()
        S, B, m = x.shape[2], world * n, n // P
        words = text_emb_l.float().contiguous()
        Bc, _, Lw = words.shape
        dev = x.device
        main, side = torch.cuda.current_stream(dev), _side_stream(dev)
        with torch.cuda.device(dev):
            spad, sp = L.gloria_b200_tc_spad(S), L.gloria_b200_tc_sp(S)
            lpad, lp = L.gloria_b200_tc_lpad(lcap), L.gloria_b200_tc_lp(lcap)
            s = _stream(x)
            # 1. every rank packs ITS OWN images only (before: all B images on every rank) ...
            own_h = torch.empty((n, spad, D), dtype=torch.float16, device=dev)
            own_t = torch.empty((n, sp, D), dtype=torch.bfloat16, device=dev)
            _lib.check(L.gloria_b200_tc_prepack_ctx(x.data_ptr(), n, D, S, own_h.data_ptr(), own_t.data_ptr(), None, s),
                       "tc_prepack_ctx")
            packed = torch.cuda.Event()
            packed.record(main)
            # 2. ... and the 16-bit copies are what is gathered (the same bytes as the fp32 features), part by part on the
            # side stream, into the part-major arrays the kernels and the backward read
            ctx_h = torch.empty((B, spad, D), dtype=torch.float16, device=dev)
            ctx_t = torch.empty((B, sp, D), dtype=torch.bfloat16, device=dev)
            nj = world * m
            gathered = []
            side.wait_event(packed)
            with torch.cuda.stream(side):
                for p in range(P):
                    dist.all_gather_into_tensor(ctx_h[p * nj:(p + 1) * nj], own_h[p * m:(p + 1) * m], group=group)
                    dist.all_gather_into_tensor(ctx_t[p * nj:(p + 1) * nj], own_t[p * m:(p + 1) * m], group=group)
                    ev = torch.cuda.Event()
                    ev.record(side)
                    gathered.append(ev)
            words_h = torch.empty((Bc, lpad, D), dtype=torch.float16, device=dev)
            words_t = torch.empty((Bc, lp, D), dtype=torch.bfloat16, device=dev)
            wnorm = torch.empty((Bc, lpad), dtype=torch.float32, device=dev)
            nbytes = L.gloria_b200_tc_train_workspace(B, Bc, D, S, lcap)
            ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
            sim_pm = torch.empty((B, Bc), dtype=torch.float32, device=dev)
            _lib.check(L.gloria_b200_tc_prepack_words(words.data_ptr(), dev_lens.data_ptr(), Bc, D, Lw, lcap, 0,
                                                      words_h.data_ptr(), words_t.data_ptr(), wnorm.data_ptr(), s),
                       "tc_prepack_words")

            def launch(h, t, j0, cnt, flags):
                _lib.check(L.gloria_b200_tc_local_sim_fwd_train_range(
                    h.data_ptr(), t.data_ptr(), words_h.data_ptr(), wnorm.data_ptr(), dev_lens.data_ptr(), B, j0, cnt, Bc,
                    D, S, lcap, temp1, temp2, agg, eps, sim_pm.data_ptr(), ws.data_ptr(), nbytes, flags, None, 0, s),
                    "tc_local_sim_fwd_train_range")

            # 3. one launch per part as its gather lands (part 0's gather is the only exposed one).  Running the rank's own
            # images first from the local pack was tried and is slower: it takes 3 P launches instead of P, and at 8 ranks a
            # launch over 32 images x 64 captions fills the 148 CTAs for 14 pair slots only.
            for p in range(P):
                main.wait_event(gathered[p])
                launch(ctx_h[p * nj:], ctx_t[p * nj:], p * nj, nj, (5 if p == 0 else 0) | (2 if p == P - 1 else 0))
            for t in (own_h, own_t):
                t.record_stream(side)
            perm = _part_major_rows_on(dev, n, world, P)
            sim = sim_pm.index_select(0, perm)
        ctx.save_for_backward(ctx_t, words_t, dev_lens, ws, perm)
        ctx.args = (lcap, group, P, S, Lw, img_emb_l.shape, img_emb_l.dtype, text_emb_l.dtype)
        return sim

    @staticmethod
    def backward(ctx, dsim):
        import ctypes
        from . import _lib
        from .ops import _stream
        L = _lib.lib()
        from . import ops
        ctx_t, words_t, dev_lens, ws, perm = ctx.saved_tensors
        lcap, group, P, S, Lw, img_shape, img_dtype, txt_dtype = ctx.args
        if ops._BWD_MUTATES:
            if getattr(ctx, "_gloria_consumed", False):
                raise RuntimeError("gloria_b200: the fused training state was already consumed by a backward pass")
            ctx._gloria_consumed = True
        world = dist.get_world_size(group)
        B, _, D = ctx_t.shape
        Bc = words_t.shape[0]
        n = B // world
        m = n // P
        dev = ctx_t.device
        main, side = torch.cuda.current_stream(dev), _side_stream(dev)
        with torch.cuda.device(dev):
            dsim_pm = torch.empty((B, Bc), dtype=torch.float32, device=dev)
            dsim_pm.index_copy_(0, perm, dsim.float().contiguous())
            d_words = torch.empty((Bc, D, Lw), dtype=torch.float32, device=dev)
            d_ctx = torch.empty((n, D, S), dtype=torch.float32, device=dev)
            events = []
            for _ in range(P):
                e = torch.cuda.Event()
                e.record(main)                                    # materialise the handle; the library re-records it
                events.append(e)
            handles = (ctypes.c_void_p * P)(*[e.cuda_event for e in events])
            if _PACKED_RS:
                # the image-side gradient is reduce_scattered in the library's packed layout ([image, region row, D], straight
                # out of the workspace) and only this rank's n images are transposed into [n, D, S] afterwards: before,
                # every rank transposed all B images (0.4 ms of an 8-rank step) ahead of the collective
                sp = L.gloria_b200_tc_sp(S)
                off = L.gloria_b200_tc_train_drt_offset(B, Bc, D, S, lcap)
                grad_all = ws[off:off + B * sp * D * 4].view(torch.float32).view(B, sp, D)          # part-major
                own = torch.empty((n, sp, D), dtype=torch.float32, device=dev)
                out_ptr = None
            else:
                grad_all = torch.empty((B, D, S), dtype=torch.float32, device=dev)                  # part-major
                own = d_ctx
                out_ptr = grad_all.data_ptr()
            _lib.check(L.gloria_b200_tc_local_sim_bwd_train_parts(
                ctx_t.data_ptr(), words_t.data_ptr(), dev_lens.data_ptr(), B, Bc, D, S, Lw, lcap, 0, dsim_pm.data_ptr(),
                out_ptr, d_words.data_ptr(), ws.data_ptr(), ws.numel(), P,
                ctypes.cast(handles, ctypes.c_void_p), _stream(ctx_t)), "tc_local_sim_bwd_train_parts")
            for p in range(P):
                side.wait_event(events[p])                        # rows of part p are final; later parts still compute
                with torch.cuda.stream(side):
                    dist.reduce_scatter_tensor(own[p * m:(p + 1) * m], grad_all[p * world * m:(p + 1) * world * m],
                                               op=dist.ReduceOp.SUM, group=group)
                    if _PACKED_RS:
                        _lib.check(L.gloria_b200_tc_unpack_dctx(own[p * m:].data_ptr(), d_ctx[p * m:].data_ptr(), m, D, S,
                                                                side.cuda_stream), "tc_unpack_dctx")
            main.wait_stream(side)
        return (d_ctx.reshape(img_shape).to(img_dtype), d_words.to(txt_dtype), None, None, None, None, None, None,
                None, None)


_AGREED: "dict[tuple, bool]" = {}


def _agree(dev, lcap: int, parts_ok: bool, group):
    """Every rank must issue the same sequence of collectives: the pipelined path runs P all_gathers / reduce_scatters of
    n / P rows, the fallback one of n rows, and the choice depends on rank-local facts (free memory, the longest local
    caption).  One small all_reduce makes it global: lcap = max over ranks, parts path only if every rank can take it."""
    t = torch.tensor([int(lcap), -int(bool(parts_ok))], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    lcap_all, neg_ok = (int(v) for v in t.tolist())
    return lcap_all, neg_ok == -1


def _agree_cached(key, dev, lcap: int, parts_ok: bool, group):
    """`_agree` costs a host synchronisation, so the answer is asked once per configuration (shapes, world size, padded
    caption length) and reused: the sharded path pads to the word axis, which is the same on every rank, and the
    workspace question has the same answer every step unless memory runs out (which then fails loudly, not silently)."""
    if key not in _AGREED:
        _, ok = _agree(dev, lcap, parts_ok, group)
        _AGREED[key] = ok
    return _AGREED[key]


def _sharded_local_sim(img_emb_l, text_emb_l, cap_lens, temp1, temp2, agg, group, eps=1e-8):
    """Default CUDA path of the sharded local similarity -> sim[:, own captions] ([B, B/G])."""
    from . import _lib, gloria_loss, ops
    if not img_emb_l.is_cuda:
        raise RuntimeError("gloria_b200 runs on CUDA tensors only (sm_100a); there is no CPU fallback")
    Bc, D, Lw = text_emb_l.shape
    n = img_emb_l.shape[0]
    S = img_emb_l.shape[2] * img_emb_l.shape[3]
    dev_lens, lens = gloria_loss._cap_lens(cap_lens, Bc, 0, Lw, img_emb_l.device)
    mode = gloria_loss._mode(img_emb_l, text_emb_l)
    lcap = Lw          # pad to the word axis: the same on every rank without asking (the kernels mask per caption)
    L = _lib.lib()
    world = dist.get_world_size(group)
    P = _N_PARTS if n % _N_PARTS == 0 and n >= 2 * _N_PARTS else 1
    want_parts = (mode == ops.MODE_BF16 and ops._FUSED_TRAIN and agg != "max" and torch.is_grad_enabled()
                  and (img_emb_l.requires_grad or text_emb_l.requires_grad))
    parts_ok = False
    if want_parts and L.gloria_b200_tc_supported(D, S, Lw) == 0:
        # sized for the longest caption any rank may hold (Lw), so the answer cannot flip when lcap is agreed below
        nbytes = L.gloria_b200_tc_train_workspace(world * n, Bc, D, S, Lw)
        parts_ok = 0 < nbytes <= min(ops._TC_WS_BUDGET, int(ops._available_bytes(img_emb_l.device, nbytes) * 0.92))
    if want_parts:
        parts_ok = _agree_cached((world, n, Bc, D, S, Lw, id(group)), img_emb_l.device, lcap, parts_ok, group)
    if parts_ok:
        return _ShardedLocalSimParts.apply(img_emb_l, text_emb_l, dev_lens, lcap, float(temp1), float(temp2),
                                           ops.AGG[agg], float(eps), group, P)
    return _ShardedLocalSim.apply(img_emb_l, text_emb_l, dev_lens, lcap, float(temp1), float(temp2),
                                  ops.AGG[agg], float(eps), mode, group)


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

    Gradient scale under DDP: every rank gets the gradient of the FULL-batch loss w.r.t. its own shard of the inputs.
    DistributedDataParallel then AVERAGES parameter gradients over ranks, which leaves them at 1 / world_size of the
    single-device full-batch gradient: multiply the returned loss by world_size (or register a SUM communication
    hook) to reproduce single-device training exactly.
    """
    local_sim_fn = local_sim_fn or _default_local_sim
    global_cos_fn = global_cos_fn or _default_global_cos
    ce_fn = ce_fn or _default_ce
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        sim = local_sim_fn(img_emb_l, text_emb_l, cap_lens, temp1, temp2, agg)
        cosm = global_cos_fn(img_emb_g, text_emb_g, eps)
        return (*ce_fn(sim, temp3), *ce_fn(cosm, temp3))
    if local_sim_fn is _default_local_sim:
        sim_blk = _sharded_local_sim(img_emb_l, text_emb_l, cap_lens, temp1, temp2, agg, group, eps)   # [B, B/G]
    else:
        img_all = gather_reduce_scatter(img_emb_l, group)            # [B, D, H, W]
        sim_blk = local_sim_fn(img_all, text_emb_l, cap_lens, temp1, temp2, agg)      # [B, B/G]
    # (gathered AFTER the local term was built: autograd runs backward nodes in reverse creation order, so the small
    # reduce_scatter of d_img_g then precedes the local backward instead of trailing the whole step by a launch latency)
    img_g_all = gather_reduce_scatter(img_emb_g, group)              # [B, D]
    cos_blk = global_cos_fn(img_g_all, text_emb_g, eps)                                # [B, B/G]
    # one small collective for both logit blocks: rows = captions of every rank after the transpose
    both = torch.stack([sim_blk.t(), cos_blk.t()], 1)                # [B/G, 2, B]
    both_all = gather_cat(both, group)                               # [B, 2, B]   (caption-major)
    sim = both_all[:, 0].t()                                         # [B_img, B_cap]
    cosm = both_all[:, 1].t()
    return (*ce_fn(sim.contiguous(), temp3), *ce_fn(cosm.contiguous(), temp3))


def _global_mean(local_mean: torch.Tensor, group) -> torch.Tensor:
    """Mean over the global batch of a per-rank mean (equal shard sizes): the VALUE is the all-reduced one, the
    gradient that flows back is this rank's share of it (1 / world of its local mean)."""
    world = dist.get_world_size(group)
    tot = local_mean.detach().clone()
    dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    return local_mean / world + (tot - local_mean.detach()) / world


def sharded_calc_loss(model, img_emb_l: torch.Tensor, img_emb_g: torch.Tensor, text_emb_l: torch.Tensor,
                      text_emb_g: torch.Tensor, sents, segmentation_labels: Optional[torch.Tensor] = None, group=None,
                      attn_maps_fn: Optional[Callable] = None, seg_loss_fn: Optional[Callable] = None, **loss_fns):
    """`GLoRIA.calc_loss` (gloria/models/gloria_model.py:132-150) from per-rank shards -> (loss, attn_maps).

    `model` supplies the attributes the reference's __init__ sets (temp1/2/3, local_loss_weight, global_loss_weight,
    segmentation_loss_weight, no_attn_vec).  The contrastive terms are `sharded_loss` (full-batch negatives); the
    attention maps of the diagonal pairs and the supervised-attention term (:143-147) are local by nature -- pair (i, i)
    lives on the rank that owns image and caption i -- so they are computed on the rank's own pairs, and the term's
    batch mean is all-reduced (value) while its gradient stays with the owner.  `attn_maps` holds the maps of THIS
    rank's pairs.  The optional no-attention / KL / entropy regularisers (gloria_loss.py:108-139) compare every image's
    own map with its maps under all other captions; they are not sharded: configure them on a single device.
    """
    from . import gloria_loss
    from .gloria_model import cap_lens_from_sents
    if any(getattr(model, k, None) is not None for k in
           ("no_attn_loss_weight", "attention_divergence_loss_weight", "attention_entropy_loss_weight")):
        raise RuntimeError("sharded_calc_loss: the attention regularisers need every pair's word-mean attention on one "
                           "device; run calc_loss unsharded for these configurations")
    if isinstance(sents, gloria_loss.DeviceCapLens) or (len(sents) and isinstance(sents[0], int)):
        cap_lens = sents                       # caption lengths given directly
    else:
        cap_lens = cap_lens_from_sents(sents)  # word lists (or this package's LazySentences)
    nav = getattr(model, "no_attn_vec", None)
    sharded = dist.is_initialized() and dist.get_world_size(group) > 1
    loss = 0
    lw, gw = model.local_loss_weight, model.global_loss_weight
    if lw != 0 or gw != 0:
        if nav is not None and lw != 0:
            raise RuntimeError("sharded_calc_loss: no_attn_vec is not supported on the sharded contrastive path")
        l0, l1, g0, g1 = sharded_loss(img_emb_l, text_emb_l, img_emb_g, text_emb_g, cap_lens, temp1=model.temp1,
                                      temp2=model.temp2, temp3=model.temp3, group=group, **loss_fns)
        loss = (l0 + l1) * lw + (g0 + g1) * gw
    attn_maps_fn = attn_maps_fn or (lambda i, t, c: gloria_loss.diagonal_attention_maps(i, t, c, temp1=model.temp1,
                                                                                       no_attn_vec=nav))
    attn_maps = attn_maps_fn(img_emb_l, text_emb_l, cap_lens)
    if segmentation_labels is not None and getattr(model, "segmentation_loss_weight", None):
        seg = (seg_loss_fn or gloria_loss.supervised_attention_loss)(attn_maps, segmentation_labels)
        if sharded:
            seg = _global_mean(seg, group)
        loss = loss + seg * model.segmentation_loss_weight
    return loss, attn_maps
