"""torch custom ops (with autograd) over the C ABI of libgloria_b200.so.

PyTorch is plumbing here: it owns device memory and streams and records the autograd graph; every number is
computed by the CUDA kernels behind `include/gloria_b200.h`.  The ops are opaque to torch.compile
(`torch.library.custom_op` + `register_fake` + `register_autograd`; the byte state the bf16 forward hands to its
backward has a data-dependent length, declared as such in the fake) and hold no process-global state.  There is no CPU implementation: calling an
op with CPU tensors raises RuntimeError.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import threading
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib

AGG = {"sum": 0, "mean": 1, "max": 2}
MODE_FP32, MODE_BF16 = 0, 1

_WS_BUDGET = int(os.environ.get("GLORIA_B200_WS_BYTES", str(4 << 30)))            # fp32 kernels
_TC_WS_BUDGET = int(os.environ.get("GLORIA_B200_TC_WS_BYTES", str(96 << 30)))     # bf16 backward operand matrices
_F32_TC = os.environ.get("GLORIA_B200_F32_TC", "1") != "0"                         # fp32 mode on the tensor cores (tc_f32.cu)
_F32_TC_WS_BUDGET = int(os.environ.get("GLORIA_B200_F32_TC_WS_BYTES", str(16 << 30)))
_PACKED_PROMPTS = os.environ.get("GLORIA_B200_PACKED_PROMPTS", "1") != "0"           # packed short-caption inference kernel
_FUSED_TRAIN = os.environ.get("GLORIA_B200_FUSED_TRAIN", "1") != "0"               # fused forward+backward-operand kernel
_FUSED_DIAG = os.environ.get("GLORIA_B200_FUSED_DIAG", "1") != "0"                 # ... which also emits the diagonal attention maps


def _f32_tc(L, D: int, S: int, lcap: int) -> bool:
    """fp32 mode on the tensor cores (split-precision tcgen05 GEMMs, tc_f32.cu) where the shape allows it
    (D % 64 == 0, captions of <= 128 words), else the CUDA-core kernels (simt_f32.cu)."""
    return _F32_TC and L.gloria_b200_f32tc_supported(D, S, lcap) == 0


def _f32_forward(L, ctx, words, cap_lens, lcap, word_off, temp1, temp2, agg, eps, sim, diag, mean, need_grad) -> Tensor:
    """fp32-mode forward; returns the state for the backward (uint8; empty = the backward recomputes)."""
    Bi, D, S = ctx.shape
    Bc, _, Lw = words.shape
    dev, st = ctx.device, _stream(ctx)
    empty = torch.empty((0,), dtype=torch.uint8, device=dev)
    args = (Bc, D, S, Lw, lcap, word_off, temp1, temp2, agg, eps)
    if not _f32_tc(L, D, S, lcap):
        nbytes = L.gloria_b200_local_f32_workspace(Bi, Bc, D, S, Lw, lcap, _WS_BUDGET)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        rc = L.gloria_b200_local_sim_fwd_f32(ctx.data_ptr(), words.data_ptr(), cap_lens.data_ptr(), Bi, *args, sim.data_ptr(),
                                             _ptr(diag), _ptr(mean), ws.data_ptr(), nbytes, st)
        _lib.check(rc, "local_sim_fwd_f32")
        return empty
    if need_grad and agg != AGG["max"]:
        # training: keep the forward's operand pieces, P, A and contexts for the backward when they fit in one chunk
        nbytes = L.gloria_b200_local_f32tc_workspace(Bi, Bc, D, S, Lw, lcap, 0, 1)
        if nbytes <= min(_F32_TC_WS_BUDGET, int(_available_bytes(dev, nbytes) * 0.9)):
            state = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
            rc = L.gloria_b200_local_sim_fwd_f32tc_train(ctx.data_ptr(), words.data_ptr(), cap_lens.data_ptr(), Bi, *args,
                                                         sim.data_ptr(), _ptr(diag), _ptr(mean), state.data_ptr(), nbytes, st)
            _lib.check(rc, "local_sim_fwd_f32tc_train")
            return state
    # forward only (or a state too large to keep): captions are chunked inside the library; many images against few
    # captions (zero-shot scoring) are cut into image blocks here so that every block's buffers fit the budget
    nj = Bi
    if diag.numel() == 0:
        while nj > 64 and L.gloria_b200_local_f32tc_workspace(nj, Bc, D, S, Lw, lcap, 0, 0) > _F32_TC_WS_BUDGET:
            nj = (nj + 1) // 2
    nbytes = L.gloria_b200_local_f32tc_workspace(nj, Bc, D, S, Lw, lcap, _F32_TC_WS_BUDGET, 0)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    for j0 in range(0, Bi, nj):
        j1 = min(Bi, j0 + nj)
        rc = L.gloria_b200_local_sim_fwd_f32tc(ctx[j0:j1].data_ptr(), words.data_ptr(), cap_lens.data_ptr(), j1 - j0, *args,
                                               sim[j0:j1].data_ptr(), _ptr(diag), _ptr(mean[j0:j1]) if mean.numel() else None,
                                               ws.data_ptr(), nbytes, st)
        _lib.check(rc, "local_sim_fwd_f32tc")
    return empty


def _f32_backward(L, ctx, words, cap_lens, lcap, word_off, temp1, temp2, agg, eps, dsim, d_diag, d_mean, d_ctx, d_words,
                  state) -> None:
    Bi, D, S = ctx.shape
    Bc, _, Lw = words.shape
    dev, st = ctx.device, _stream(ctx)
    args = (Bi, Bc, D, S, Lw, lcap, word_off, temp1, temp2, agg, eps, dsim.data_ptr(), _ptr(d_diag), _ptr(d_mean),
            d_ctx.data_ptr(), d_words.data_ptr())
    if not _f32_tc(L, D, S, lcap):
        nbytes = L.gloria_b200_local_f32_workspace(Bi, Bc, D, S, Lw, lcap, _WS_BUDGET)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        rc = L.gloria_b200_local_sim_bwd_f32(ctx.data_ptr(), words.data_ptr(), cap_lens.data_ptr(), *args, ws.data_ptr(), nbytes, st)
        _lib.check(rc, "local_sim_bwd_f32")
        return
    if state is not None and state.numel() > 0:
        rc = L.gloria_b200_local_sim_bwd_f32tc(ctx.data_ptr(), words.data_ptr(), cap_lens.data_ptr(), *args, state.data_ptr(),
                                               state.numel(), 1, st)
    else:
        nbytes = L.gloria_b200_local_f32tc_workspace(Bi, Bc, D, S, Lw, lcap, _F32_TC_WS_BUDGET, 1)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        rc = L.gloria_b200_local_sim_bwd_f32tc(ctx.data_ptr(), words.data_ptr(), cap_lens.data_ptr(), *args, ws.data_ptr(), nbytes,
                                               0, st)
    _lib.check(rc, "local_sim_bwd_f32tc")


class Packed:
    """TMA-legal 16-bit copies of one (context, words) pair (see gloria_b200_tc_prepack in include/gloria_b200.h)."""
    __slots__ = ("ctx_h", "ctx_t", "ctx_n", "words_h", "words_t", "wnorm")

    def __init__(self, *t):
        self.ctx_h, self.ctx_t, self.ctx_n, self.words_h, self.words_t, self.wnorm = t


_SHARED = threading.local()


@contextlib.contextmanager
def shared_ctx_pack(ctx: Tensor):
    """Within the block, every tc_prepack of this very context tensor (same storage, shape and version; the block keeps
    it alive) reuses one set of packed region copies -- the length-bucketed path launches once per bucket on the same
    images."""
    _SHARED.key = (ctx.data_ptr(), tuple(ctx.shape), ctx._version, ctx.device)
    _SHARED.keep, _SHARED.val = ctx, None
    try:
        yield
    finally:
        _SHARED.key = _SHARED.keep = _SHARED.val = None


def tc_prepack(ctx: Tensor, words: Tensor, cap_lens: Tensor, lcap: int, word_off: int, ctx_t: Optional[Tensor] = None,
               words_t: Optional[Tensor] = None) -> Packed:
    """fp32 native layouts -> ctx_h/ctx_t [Bi,Spad,D] (fp16/bf16), ctx_n [Bi,D,Spad] bf16, words_h/words_t [Bc,Lpad,D]
    (fp16/bf16), wnorm [Bc,Lpad] fp32.  ctx_t / words_t may be given as uint8 views to write into (the training state
    carries them to the backward); ctx_n is then not produced (only the inference / recompute kernels read it)."""
    L = _lib.lib()
    Bi, D, S = ctx.shape
    Bc, _, Lw = words.shape
    spad, lpad = L.gloria_b200_tc_spad(S), L.gloria_b200_tc_lpad(lcap)
    dev = ctx.device
    words_h = torch.empty((Bc, lpad, D), dtype=torch.float16, device=dev)
    lean = ctx_t is not None
    if words_t is None:
        words_t = torch.empty((Bc, L.gloria_b200_tc_lp(lcap), D), dtype=torch.bfloat16, device=dev)
    else:
        words_t = words_t[:Bc * L.gloria_b200_tc_lp(lcap) * D * 2].view(torch.bfloat16).view(Bc, L.gloria_b200_tc_lp(lcap), D)
    wnorm = torch.empty((Bc, lpad), dtype=torch.float32, device=dev)
    key = (ctx.data_ptr(), tuple(ctx.shape), ctx._version, ctx.device)
    shared = getattr(_SHARED, "key", None) == key
    sp = L.gloria_b200_tc_sp(S)
    if shared and _SHARED.val is not None:
        ctx_h, ctx_t0, ctx_n = _SHARED.val
        if lean:      # the state needs its own copy of the (shared) ctx_t: one device copy instead of a second pack
            ctx_t = ctx_t[:Bi * sp * D * 2].view(torch.bfloat16).view(Bi, sp, D)
            ctx_t.copy_(ctx_t0)
        else:
            ctx_t = ctx_t0
    else:
        ctx_h = torch.empty((Bi, spad, D), dtype=torch.float16, device=dev)
        if lean:
            ctx_t = ctx_t[:Bi * sp * D * 2].view(torch.bfloat16).view(Bi, sp, D)
        else:
            ctx_t = torch.empty((Bi, sp, D), dtype=torch.bfloat16, device=dev)
        ctx_n = None if (lean and not shared) else torch.empty((Bi, D, spad), dtype=torch.bfloat16, device=dev)
        rc = L.gloria_b200_tc_prepack_ctx(ctx.data_ptr(), Bi, D, S, ctx_h.data_ptr(), ctx_t.data_ptr(), _ptr(ctx_n),
                                          _stream(ctx))
        _lib.check(rc, "tc_prepack_ctx")
        if shared:
            _SHARED.val = (ctx_h, ctx_t, ctx_n)
    rc = L.gloria_b200_tc_prepack_words(words.data_ptr(), cap_lens.data_ptr(), Bc, D, Lw, lcap, word_off,
                                        words_h.data_ptr(), words_t.data_ptr(), wnorm.data_ptr(), _stream(ctx))
    _lib.check(rc, "tc_prepack_words")
    return Packed(ctx_h, ctx_t, ctx_n, words_h, words_t, wnorm)


_TOTAL_MEM: "dict[int, int]" = {}


def _available_bytes(dev: torch.device, want: int) -> int:
    """Bytes a new allocation could get on `dev`.  cudaMemGetInfo costs ~2 ms of host time per call (more than a whole
    B=48 step), so it is only asked when the request is large against what the caching allocator's own counters say
    is left; small requests are answered from those counters (a wrong guess surfaces as torch's OOM error)."""
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _TOTAL_MEM:
        _TOTAL_MEM[idx] = torch.cuda.get_device_properties(idx).total_memory
    optimistic = _TOTAL_MEM[idx] - torch.cuda.memory_allocated(idx)
    if want * 4 <= optimistic:
        return optimistic
    free, _ = torch.cuda.mem_get_info(idx)
    return free + torch.cuda.memory_reserved(idx) - torch.cuda.memory_allocated(idx)


def upload_ints(values, device: torch.device) -> Tensor:
    """int32 device tensor from a short host list (caption lengths) through the kernel parameter buffer: asynchronous, no
    copy engine (a pageable H2D copy blocks the host and queues behind the next batch's prefetch on the copy engine)."""
    n = len(values)
    out = torch.empty((n,), dtype=torch.int32, device=device)
    if n:
        arr = (C.c_int32 * n)(*values)
        with torch.cuda.device(device):
            _lib.check(_lib.lib().gloria_b200_upload_ints(arr, n, out.data_ptr(), torch.cuda.current_stream(device).cuda_stream),
                       "upload_ints")
    return out


def _need_cuda(*ts: Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("gloria_b200 ops run on CUDA tensors only (sm_100a); there is no CPU fallback")


def _ptr(t: Optional[Tensor]):
    return None if t is None or t.numel() == 0 else t.data_ptr()


def _stream(t: Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def _f32c(t: Tensor) -> Tensor:
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


# ----------------------------------------------------------------------------------------------------------------
# local similarity
# ----------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("gloria_b200::local_sim_fwd", mutates_args=())
def local_sim_fwd(ctx: Tensor, words: Tensor, cap_lens: Tensor, lcap: int, word_off: int, temp1: float,
                  temp2: float, agg: int, eps: float, want_diag: bool, want_mean: bool,
                  mode: int, need_grad: bool = True) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """sim [Bi, Bc], attn_diag [Bc, lcap, S] (or empty), attn_mean [Bi, Bc, S] (or empty), state (uint8 [n]: what the
    bf16 backward needs -- the fused training workspace followed by the 16-bit operand copies, or the per-word
    scalars of the recompute path; empty in fp32 mode and when need_grad is false).

    need_grad is the CALLER's statement that a backward will follow (the host wrappers pass
    `torch.is_grad_enabled() and an input requires grad`): only then is the training forward run and its state
    (41.7 GB at B = 512) allocated.

    ctx [Bi, D, S] fp32, words [Bc, D, Lw] fp32, cap_lens int32 [Bc] on the same device.
    Replaces the caption loop of gloria_loss.py:116-162 (attention_fn + cosine_similarity + aggregation).
    """
    _need_cuda(ctx, words, cap_lens)
    L = _lib.lib()
    Bi, D, S = ctx.shape
    Bc, D2, Lw = words.shape
    if D2 != D:
        raise RuntimeError(f"feature dims differ: context {D} vs words {D2}")
    if cap_lens.dtype != torch.int32 or cap_lens.numel() != Bc:
        raise RuntimeError("cap_lens must be an int32 tensor with one entry per caption")
    ctx, words, cap_lens = ctx.contiguous(), words.contiguous(), cap_lens.contiguous()
    dev = ctx.device
    sim = torch.empty((Bi, Bc), dtype=torch.float32, device=dev)
    diag = torch.empty((Bc, lcap, S) if want_diag else (0,), dtype=torch.float32, device=dev)
    mean = torch.empty((Bi, Bc, S) if want_mean else (0,), dtype=torch.float32, device=dev)
    stats = torch.empty((0,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        if mode == MODE_FP32:
            stats = _f32_forward(L, ctx, words, cap_lens, lcap, word_off, temp1, temp2, agg, eps, sim, diag, mean, need_grad)
        else:
            if want_mean and agg == AGG["max"]:
                raise RuntimeError("word-mean attention output is not available with agg='max'")
            if L.gloria_b200_tc_supported(D, S, lcap) != 0:
                raise RuntimeError(f"bf16 tensor-core kernels need D % 128 == 0, S <= 384, cap_len <= 128; got "
                                   f"D={D} S={S} Lcap={lcap} (use set_precision('fp32'))")
            need_grad = bool(need_grad) and agg != AGG["max"]
            if not need_grad and not want_mean and not want_diag and lcap <= 16 and Bc >= 2 and _PACKED_PROMPTS:
                # forward-only scoring of short prompts (zero-shot): up to 8 captions share one word tile
                _tc_packed_fwd(L, ctx, words, cap_lens, word_off, temp1, temp2, agg, eps, sim)
                return sim, diag, mean, stats
            fused = diag_done = False
            lpad = L.gloria_b200_tc_lpad(lcap)
            if need_grad and not want_mean and _FUSED_TRAIN:
                # fused training forward: sim AND the backward's operand rows (for dsim = 1) in one kernel.  The state
                # tensor is [workspace | ctx_t | words_t]: the backward needs nothing else (no global hand-over cache).
                # Falls back to forward + recompute-backward when the workspace does not fit.
                nbytes = L.gloria_b200_tc_train_workspace(Bi, Bc, D, S, lcap)
                n_ct, n_wt = _ctx_t_bytes(L, Bi, D, S), _words_t_bytes(L, Bc, D, lcap)
                total = nbytes + n_ct + n_wt
                if 0 < nbytes and total <= min(_TC_WS_BUDGET, int(_available_bytes(dev, total) * 0.92)):
                    stats = torch.empty((total,), dtype=torch.uint8, device=dev)
                    packed = tc_prepack(ctx, words, cap_lens, lcap, word_off, ctx_t=stats[nbytes:nbytes + n_ct],
                                        words_t=stats[nbytes + n_ct:])
                    if want_diag and Bi == Bc and _FUSED_DIAG:
                        # the attention maps of the diagonal pairs come out of the same kernel (its softmax stores them)
                        rc = L.gloria_b200_tc_local_sim_fwd_train_diag(
                            packed.ctx_h.data_ptr(), packed.ctx_t.data_ptr(), packed.words_h.data_ptr(),
                            packed.wnorm.data_ptr(), cap_lens.data_ptr(), Bi, Bc, D, S, lcap, temp1, temp2, agg, eps,
                            sim.data_ptr(), stats.data_ptr(), nbytes, diag.data_ptr(), lcap, _stream(ctx))
                        _lib.check(rc, "tc_local_sim_fwd_train_diag")
                        diag_done = True
                    else:
                        rc = L.gloria_b200_tc_local_sim_fwd_train(packed.ctx_h.data_ptr(), packed.ctx_t.data_ptr(),
                                                                  packed.words_h.data_ptr(), packed.wnorm.data_ptr(),
                                                                  cap_lens.data_ptr(), Bi, Bc, D, S, lcap, temp1, temp2,
                                                                  agg, eps, sim.data_ptr(), stats.data_ptr(), nbytes,
                                                                  _stream(ctx))
                        _lib.check(rc, "tc_local_sim_fwd_train")
                    fused = True
            if not fused:
                packed = tc_prepack(ctx, words, cap_lens, lcap, word_off)
                fstats = None
                if need_grad:           # per-word scalars for the recompute backward, carried as bytes
                    stats = torch.empty((Bi * Bc * 2 * lpad * 4,), dtype=torch.uint8, device=dev)
                    fstats = stats.view(torch.float32)
                if want_mean:
                    # regulariser configs: one kernel gives sim + the word-mean attention of every pair (+ the per-word
                    # scalars); their backward is the recompute kernel, which takes d(attn_mean) next to dsim
                    nbytes = L.gloria_b200_tc_mean_workspace(Bi, Bc, D, S, lcap)
                    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
                    rc = L.gloria_b200_tc_local_sim_fwd_mean(packed.ctx_h.data_ptr(), packed.ctx_t.data_ptr(),
                                                             packed.words_h.data_ptr(), packed.wnorm.data_ptr(),
                                                             cap_lens.data_ptr(), Bi, Bc, D, S, lcap, temp1, temp2, agg,
                                                             eps, sim.data_ptr(), mean.data_ptr(), _ptr(fstats),
                                                             ws.data_ptr(), nbytes, _stream(ctx))
                    _lib.check(rc, "tc_local_sim_fwd_mean")
                else:
                    rc = L.gloria_b200_tc_local_sim_fwd(packed.ctx_h.data_ptr(), packed.ctx_n.data_ptr(),
                                                        packed.words_h.data_ptr(), packed.wnorm.data_ptr(),
                                                        cap_lens.data_ptr(), Bi, Bc, D, S, lcap, temp1, temp2, agg, eps,
                                                        sim.data_ptr(), _ptr(fstats), _stream(ctx))
                    _lib.check(rc, "tc_local_sim_fwd")
            if want_diag and not diag_done:
                if Bi != Bc:
                    raise RuntimeError(f"diagonal attention maps need as many images as captions, got {Bi} x {Bc}")
                # diagonal attention maps: B pairs (not B^2) through the exact fp32 kernels
                nbytes = L.gloria_b200_diag_attn_workspace(Bc, D, S, Lw, lcap)
                ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
                rc = L.gloria_b200_diag_attn_fwd_f32(ctx.data_ptr(), words.data_ptr(), cap_lens.data_ptr(), Bc, D, S,
                                                     Lw, lcap, word_off, temp1, diag.data_ptr(), ws.data_ptr(),
                                                     nbytes, _stream(ctx))
                _lib.check(rc, "diag_attn_fwd_f32")
    return sim, diag, mean, stats


def _ctx_t_bytes(L, Bi, D, S) -> int:
    return (Bi * L.gloria_b200_tc_sp(S) * D * 2 + 1023) // 1024 * 1024


def _words_t_bytes(L, Bc, D, lcap) -> int:
    return (Bc * L.gloria_b200_tc_lp(lcap) * D * 2 + 1023) // 1024 * 1024


def _tc_packed_fwd(L, ctx, words, cap_lens, word_off, temp1, temp2, agg, eps, sim) -> None:
    """sim[Bi, Bc] through the packed-prompt kernel (gloria_b200_tc_local_sim_fwd_packed).  (Packing image parts on a side
    stream under the scoring kernel of the previous part was tried and measured slower, 11.05 vs 10.81 ms at 10 000 x 25:
    the scoring kernel's 384 threads x 168 registers leave no room for a second resident block.)"""
    Bi, D, S = ctx.shape
    Bc, _, Lw = words.shape
    dev = ctx.device
    per, groups = L.gloria_b200_tc_packed_per(Bc), L.gloria_b200_tc_packed_groups(Bc)
    spad = L.gloria_b200_tc_spad(S)
    ctx_h = torch.empty((Bi, spad, D), dtype=torch.float16, device=dev)
    ctx_n = torch.empty((Bi, D, spad), dtype=torch.bfloat16, device=dev)
    words_h = torch.empty((groups, 16 * per, D), dtype=torch.float16, device=dev)
    wnorm = torch.empty((groups, 16 * per), dtype=torch.float32, device=dev)
    st = _stream(ctx)
    _lib.check(L.gloria_b200_tc_prepack_ctx(ctx.data_ptr(), Bi, D, S, ctx_h.data_ptr(), None, ctx_n.data_ptr(), st),
               "tc_prepack_ctx")
    _lib.check(L.gloria_b200_tc_prepack_words_packed(words.data_ptr(), cap_lens.data_ptr(), Bc, D, Lw, word_off,
                                                     words_h.data_ptr(), wnorm.data_ptr(), st), "tc_prepack_words_packed")
    _lib.check(L.gloria_b200_tc_local_sim_fwd_packed(ctx_h.data_ptr(), ctx_n.data_ptr(), words_h.data_ptr(),
                                                     wnorm.data_ptr(), cap_lens.data_ptr(), Bi, Bc, D, S, temp1, temp2,
                                                     agg, eps, sim.data_ptr(), st), "tc_local_sim_fwd_packed")


@local_sim_fwd.register_fake
def _(ctx, words, cap_lens, lcap, word_off, temp1, temp2, agg, eps, want_diag, want_mean, mode, need_grad=True):
    Bi, D, S = ctx.shape
    Bc = words.shape[0]
    if (mode == MODE_BF16 or _F32_TC) and need_grad and agg != AGG["max"]:
        # fused-training workspace / recompute statistics (bf16), kept forward state or none (fp32): which one is decided
        # from free memory at run time, so the length of the byte state is a data-dependent size
        n = torch.library.get_ctx().new_dynamic_size()
        state = ctx.new_empty((n,), dtype=torch.uint8)
    else:
        state = ctx.new_empty((0,), dtype=torch.uint8)
    return (ctx.new_empty((Bi, Bc)), ctx.new_empty((Bc, lcap, S) if want_diag else (0,)),
            ctx.new_empty((Bi, Bc, S) if want_mean else (0,)), state)


# (`stats` is the forward's opaque byte state.  On the fused bf16 path the library scales the operand rows inside it by dsim
# in place -- it is produced by one forward and consumed by one backward, `check_state_unconsumed` and a device-side flag
# refuse a second use -- so no caller-visible tensor value changes; it is therefore not listed in mutates_args, which
# would make the functionaliser clone 41.7 GB at B = 512.)
@torch.library.custom_op("gloria_b200::local_sim_bwd", mutates_args=())
def local_sim_bwd(ctx: Tensor, words: Tensor, cap_lens: Tensor, lcap: int, word_off: int, temp1: float,
                  temp2: float, agg: int, eps: float, dsim: Optional[Tensor], d_diag: Optional[Tensor],
                  d_mean: Optional[Tensor], stats: Optional[Tensor], mode: int,
                  dctx_event: int = 0) -> Tuple[Tensor, Tensor]:
    """Closed-form backward by recomputation: returns d_ctx [Bi, D, S], d_words [Bc, D, Lw].

    dsim None = no gradient reaches the similarity matrix (attention fine-tune with both contrastive weights 0,
    gloria_model.py:138-147): only the B diagonal pairs are differentiated.
    dctx_event: raw cudaEvent_t handle (0 = none) recorded on the current stream once d_ctx is final -- on the fused
    training path that is before the caption-side GEMM, so a caption-sharded caller can overlap its reduce_scatter."""
    _need_cuda(ctx, words, cap_lens)
    L = _lib.lib()
    Bi, D, S = ctx.shape
    Bc, _, Lw = words.shape
    ctx, words, cap_lens = ctx.contiguous(), words.contiguous(), cap_lens.contiguous()
    if stats is not None and stats.numel() == 0:
        stats = None
    dsim = None if dsim is None else _f32c(dsim)
    d_diag = None if d_diag is None else _f32c(d_diag)
    d_mean = None if d_mean is None else _f32c(d_mean)
    dev = ctx.device
    d_ctx = torch.empty_like(ctx)
    d_words = torch.empty_like(words)
    st = _stream(ctx)
    with torch.cuda.device(dev):
        all_pairs = dsim is not None or d_mean is not None
        diag_separately = d_diag is not None and (not all_pairs or mode == MODE_BF16)
        if all_pairs:
            if dsim is None:
                dsim = torch.zeros((Bi, Bc), dtype=torch.float32, device=dev)
            if mode == MODE_BF16:
                ev_done = tc_local_sim_bwd(L, ctx, words, cap_lens, stats, lcap, word_off, temp1, temp2, agg, eps, dsim,
                                           d_mean, d_ctx, d_words, 0 if diag_separately else dctx_event)
                if ev_done:
                    dctx_event = 0
            else:
                _f32_backward(L, ctx, words, cap_lens, lcap, word_off, temp1, temp2, agg, eps, dsim,
                              None if diag_separately else d_diag, d_mean, d_ctx, d_words, stats)
        if diag_separately:
            nbytes = L.gloria_b200_diag_attn_workspace(Bc, D, S, Lw, lcap)
            ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
            rc = L.gloria_b200_diag_attn_bwd_f32(ctx.data_ptr(), words.data_ptr(), cap_lens.data_ptr(), Bc, D, S, Lw,
                                                 lcap, word_off, temp1, d_diag.data_ptr(), d_ctx.data_ptr(),
                                                 d_words.data_ptr(), 1 if all_pairs else 0, ws.data_ptr(), nbytes, st)
            _lib.check(rc, "diag_attn_bwd_f32")
        elif not all_pairs:
            d_ctx.zero_()
            d_words.zero_()
        if dctx_event:
            _lib.check(L.gloria_b200_record_event(dctx_event, st), "record_event")
    return d_ctx, d_words


def tc_local_sim_bwd(L, ctx, words, cap_lens, stats, lcap, word_off, temp1, temp2, agg, eps, dsim, d_mean, d_ctx,
                     d_words, dctx_event=0) -> bool:
    """bf16 tensor-core backward behind the C ABI: from the fused training state, or by recomputation (prepack, fused
    recompute kernel, accumulation GEMMs).  Returns True when `dctx_event` was recorded by the library."""
    Bi, D, S = ctx.shape
    Bc, _, Lw = words.shape
    have = stats is not None and stats.numel() > 0
    lpad = L.gloria_b200_tc_lpad(lcap)
    n_stats = Bi * Bc * 2 * lpad * 4
    if have and stats.numel() != n_stats:
        # state of the fused training forward = [workspace | ctx_t | words_t]: scale by dsim + accumulation GEMMs, nothing
        # is recomputed or re-packed.  (A second backward over the same state is caught on the device: the library
        # poisons its result with NaN rather than scaling the operands twice.)
        nbytes = L.gloria_b200_tc_train_workspace(Bi, Bc, D, S, lcap)
        n_ct = _ctx_t_bytes(L, Bi, D, S)
        if stats.numel() != nbytes + n_ct + _words_t_bytes(L, Bc, D, lcap):
            raise RuntimeError(f"gloria_b200: unexpected training-state size {stats.numel()}")
        rc = L.gloria_b200_tc_local_sim_bwd_train_ev(stats[nbytes:].data_ptr(), stats[nbytes + n_ct:].data_ptr(),
                                                     cap_lens.data_ptr(), Bi, Bc, D, S, Lw, lcap, word_off,
                                                     dsim.data_ptr(), d_ctx.data_ptr(), d_words.data_ptr(),
                                                     stats.data_ptr(), nbytes, dctx_event or None, _stream(ctx))
        _lib.check(rc, "tc_local_sim_bwd_train")
        return bool(dctx_event)
    packed = tc_prepack(ctx, words, cap_lens, lcap, word_off)
    fstats = stats.view(torch.float32) if have else None
    free, _ = torch.cuda.mem_get_info(ctx.device)
    budget = min(_TC_WS_BUDGET, int(free * 0.9) + torch.cuda.memory_reserved(ctx.device)
                 - torch.cuda.memory_allocated(ctx.device))
    nbytes = L.gloria_b200_tc_bwd_workspace(Bi, Bc, D, S, lcap, 1 if have else 0, max(budget, 1 << 30))
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=ctx.device)
    rc = L.gloria_b200_tc_local_sim_bwd(packed.ctx_h.data_ptr(), packed.ctx_t.data_ptr(), packed.ctx_n.data_ptr(),
                                        packed.words_h.data_ptr(), packed.words_t.data_ptr(), packed.wnorm.data_ptr(),
                                        cap_lens.data_ptr(), fstats.data_ptr() if have else None,
                                        Bi, Bc, D, S, Lw, lcap, word_off, temp1, temp2, agg, eps, dsim.data_ptr(),
                                        _ptr(d_mean), d_ctx.data_ptr(), d_words.data_ptr(), ws.data_ptr(), nbytes,
                                        _stream(ctx))
    _lib.check(rc, "tc_local_sim_bwd")
    return False


@local_sim_bwd.register_fake
def _(ctx, words, cap_lens, lcap, word_off, temp1, temp2, agg, eps, dsim, d_diag, d_mean, stats, mode, dctx_event=0):
    return torch.empty_like(ctx), torch.empty_like(words)


_BWD_MUTATES = os.environ.get("GLORIA_B200_BWD_GEMM", "cublas") not in ("inflight", "0")


def check_state_unconsumed(node, ctx: Tensor, words: Tensor, lcap: int, stats: Optional[Tensor], mode: int) -> None:
    """The fused training backward scales its operand state in place (unless GLORIA_B200_BWD_GEMM=inflight), so a second
    backward through the same forward (retain_graph=True) must not run: the flag lives on the autograd node of that
    forward -- per graph, nothing process-wide.  (The library additionally poisons such a result with NaN on the
    device, for callers of the C ABI.)"""
    if mode != MODE_BF16 or stats is None or not _BWD_MUTATES:
        return
    n = stats.numel()
    if not isinstance(n, int) or n == 0:              # symbolic under tracing: the device-side guard remains
        return
    lpad = (lcap + 15) // 16 * 16
    if n == ctx.shape[0] * words.shape[0] * 2 * lpad * 4:
        return                                        # per-word statistics of the recompute path: re-usable
    if getattr(node, "_gloria_consumed", False):
        raise RuntimeError("gloria_b200: the fused training state was already consumed by a backward pass (set "
                           "GLORIA_B200_FUSED_TRAIN=0 or GLORIA_B200_BWD_GEMM=inflight to differentiate the same "
                           "forward twice)")
    node._gloria_consumed = True


def _local_setup(ctx, inputs, output):
    feats, words, cap_lens, lcap, word_off, temp1, temp2, agg, eps, want_diag, want_mean, mode = inputs[:12]
    ctx.save_for_backward(feats, words, cap_lens, output[3])
    ctx.args = (lcap, word_off, temp1, temp2, agg, eps, mode)
    ctx.set_materialize_grads(False)


def _local_backward(c, dsim, d_diag, d_mean, d_stats):
    ctx, words, cap_lens, stats = c.saved_tensors
    lcap, word_off, temp1, temp2, agg, eps, mode = c.args
    if d_diag is not None and d_diag.numel() == 0:
        d_diag = None
    if d_mean is not None and d_mean.numel() == 0:
        d_mean = None
    check_state_unconsumed(c, ctx, words, lcap, stats, mode)
    # the state goes in as it is (its length is data dependent under tracing); the op treats an empty one as absent
    d_ctx, d_words = local_sim_bwd(ctx, words, cap_lens, lcap, word_off, temp1, temp2, agg, eps, dsim, d_diag,
                                   d_mean, stats, mode)
    return d_ctx, d_words, None, None, None, None, None, None, None, None, None, None, None


local_sim_fwd.register_autograd(_local_backward, setup_context=_local_setup)


# ----------------------------------------------------------------------------------------------------------------
# diagonal pairs only: attention maps of (image i, caption i)  (gloria_loss.py:141-143; gloria_model.py:143-147,209-211)
# ----------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("gloria_b200::diag_attn_fwd", mutates_args=())
def diag_attn_fwd(ctx: Tensor, words: Tensor, cap_lens: Tensor, lcap: int, word_off: int, temp1: float) -> Tensor:
    """attn [B, lcap, S] of the B diagonal pairs (rows beyond each caption's length are zero) -- B pairs of work, where
    the reference's local_loss runs all B^2 to return these maps."""
    _need_cuda(ctx, words, cap_lens)
    L = _lib.lib()
    B, D, S = ctx.shape
    Bc, D2, Lw = words.shape
    if Bc != B or D2 != D:
        raise RuntimeError(f"diagonal attention maps need matching batches / feature dims, got {tuple(ctx.shape)} and "
                           f"{tuple(words.shape)}")
    ctx, words, cap_lens = ctx.contiguous(), words.contiguous(), cap_lens.contiguous()
    diag = torch.empty((B, lcap, S), dtype=torch.float32, device=ctx.device)
    with torch.cuda.device(ctx.device):
        nbytes = L.gloria_b200_diag_attn_workspace(B, D, S, Lw, lcap)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=ctx.device)
        rc = L.gloria_b200_diag_attn_fwd_f32(ctx.data_ptr(), words.data_ptr(), cap_lens.data_ptr(), B, D, S, Lw, lcap,
                                             word_off, temp1, diag.data_ptr(), ws.data_ptr(), nbytes, _stream(ctx))
        _lib.check(rc, "diag_attn_fwd_f32")
    return diag


@diag_attn_fwd.register_fake
def _(ctx, words, cap_lens, lcap, word_off, temp1):
    return ctx.new_empty((ctx.shape[0], lcap, ctx.shape[2]))


@torch.library.custom_op("gloria_b200::diag_attn_bwd", mutates_args=())
def diag_attn_bwd(ctx: Tensor, words: Tensor, cap_lens: Tensor, lcap: int, word_off: int, temp1: float,
                  d_diag: Tensor) -> Tuple[Tensor, Tensor]:
    _need_cuda(ctx, words, cap_lens, d_diag)
    L = _lib.lib()
    B, D, S = ctx.shape
    Lw = words.shape[2]
    ctx, words, cap_lens, d_diag = ctx.contiguous(), words.contiguous(), cap_lens.contiguous(), _f32c(d_diag)
    d_ctx, d_words = torch.empty_like(ctx), torch.empty_like(words)
    with torch.cuda.device(ctx.device):
        nbytes = L.gloria_b200_diag_attn_workspace(B, D, S, Lw, lcap)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=ctx.device)
        rc = L.gloria_b200_diag_attn_bwd_f32(ctx.data_ptr(), words.data_ptr(), cap_lens.data_ptr(), B, D, S, Lw, lcap,
                                             word_off, temp1, d_diag.data_ptr(), d_ctx.data_ptr(), d_words.data_ptr(), 0,
                                             ws.data_ptr(), nbytes, _stream(ctx))
        _lib.check(rc, "diag_attn_bwd_f32")
    return d_ctx, d_words


@diag_attn_bwd.register_fake
def _(ctx, words, cap_lens, lcap, word_off, temp1, d_diag):
    return torch.empty_like(ctx), torch.empty_like(words)


def _diag_setup(ctx, inputs, output):
    feats, words, cap_lens, lcap, word_off, temp1 = inputs
    ctx.save_for_backward(feats, words, cap_lens)
    ctx.args = (lcap, word_off, temp1)


def _diag_backward(c, d_diag):
    feats, words, cap_lens = c.saved_tensors
    d_ctx, d_words = diag_attn_bwd(feats, words, cap_lens, *c.args, d_diag)
    return d_ctx, d_words, None, None, None, None


diag_attn_fwd.register_autograd(_diag_backward, setup_context=_diag_setup)


# ----------------------------------------------------------------------------------------------------------------
# global cosine similarity  (gloria_loss.py:75-80, gloria_model.py:164-169)
# ----------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("gloria_b200::global_sim_fwd", mutates_args=())
def global_sim_fwd(x: Tensor, y: Tensor, eps: float) -> Tuple[Tensor, Tensor, Tensor]:
    """cos [Bi, Bc] = <x_a, y_b> / max(|x_a||y_b|, eps), plus the saved norms."""
    _need_cuda(x, y)
    L = _lib.lib()
    x, y = x.contiguous(), y.contiguous()
    Bi, D = x.shape
    Bc = y.shape[0]
    if y.shape[1] != D:
        raise RuntimeError("global feature dims differ")
    cosm = torch.empty((Bi, Bc), dtype=torch.float32, device=x.device)
    xn = torch.empty((Bi,), dtype=torch.float32, device=x.device)
    yn = torch.empty((Bc,), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(L.gloria_b200_global_sim_fwd(x.data_ptr(), y.data_ptr(), Bi, Bc, D, eps, cosm.data_ptr(),
                                                xn.data_ptr(), yn.data_ptr(), _stream(x)), "global_sim_fwd")
    return cosm, xn, yn


@global_sim_fwd.register_fake
def _(x, y, eps):
    return x.new_empty((x.shape[0], y.shape[0])), x.new_empty((x.shape[0],)), x.new_empty((y.shape[0],))


@torch.library.custom_op("gloria_b200::global_sim_bwd", mutates_args=())
def global_sim_bwd(x: Tensor, y: Tensor, xn: Tensor, yn: Tensor, dcos: Tensor, eps: float) -> Tuple[Tensor, Tensor]:
    _need_cuda(x, y, dcos)
    L = _lib.lib()
    x, y, dcos = x.contiguous(), y.contiguous(), _f32c(dcos)
    Bi, D = x.shape
    Bc = y.shape[0]
    dx, dy = torch.empty_like(x), torch.empty_like(y)
    with torch.cuda.device(x.device):
        _lib.check(L.gloria_b200_global_sim_bwd(x.data_ptr(), y.data_ptr(), xn.data_ptr(), yn.data_ptr(),
                                                dcos.data_ptr(), Bi, Bc, D, eps, dx.data_ptr(), dy.data_ptr(),
                                                _stream(x)), "global_sim_bwd")
    return dx, dy


@global_sim_bwd.register_fake
def _(x, y, xn, yn, dcos, eps):
    return torch.empty_like(x), torch.empty_like(y)


def _global_setup(ctx, inputs, output):
    x, y, eps = inputs
    _, xn, yn = output
    ctx.save_for_backward(x, y, xn, yn)
    ctx.eps = eps
    ctx.set_materialize_grads(False)


def _global_backward(c, dcos, dxn, dyn):
    x, y, xn, yn = c.saved_tensors
    if dcos is None:
        return torch.zeros_like(x), torch.zeros_like(y), None
    dx, dy = global_sim_bwd(x, y, xn, yn, dcos, c.eps)
    return dx, dy, None


global_sim_fwd.register_autograd(_global_backward, setup_context=_global_setup)


# ----------------------------------------------------------------------------------------------------------------
# bidirectional cross entropy with arange labels  (gloria_loss.py:86-87, 164-170)
# ----------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("gloria_b200::ce_bidir_fwd", mutates_args=())
def ce_bidir_fwd(m: Tensor, scale: float) -> Tuple[Tensor, Tensor, Tensor]:
    """losses [2] = (CE(scale*m, arange), CE(scale*m^T, arange)); also the row/column log-sum-exps."""
    _need_cuda(m)
    L = _lib.lib()
    m = m.contiguous()
    if m.dim() != 2 or m.shape[0] != m.shape[1]:
        raise RuntimeError(f"cross entropy with arange labels needs a square logit matrix, got {tuple(m.shape)}")
    B = m.shape[0]
    losses = torch.empty((2,), dtype=torch.float32, device=m.device)
    row = torch.empty((B,), dtype=torch.float32, device=m.device)
    col = torch.empty((B,), dtype=torch.float32, device=m.device)
    with torch.cuda.device(m.device):
        _lib.check(L.gloria_b200_ce_bidir_fwd(m.data_ptr(), B, scale, losses.data_ptr(), row.data_ptr(),
                                              col.data_ptr(), _stream(m)), "ce_bidir_fwd")
    return losses, row, col


@ce_bidir_fwd.register_fake
def _(m, scale):
    return m.new_empty((2,)), m.new_empty((m.shape[0],)), m.new_empty((m.shape[0],))


@torch.library.custom_op("gloria_b200::ce_bidir_bwd", mutates_args=())
def ce_bidir_bwd(m: Tensor, scale: float, row: Tensor, col: Tensor, g: Tensor) -> Tensor:
    _need_cuda(m, g)
    L = _lib.lib()
    m, g = m.contiguous(), _f32c(g)
    dm = torch.empty_like(m)
    with torch.cuda.device(m.device):
        _lib.check(L.gloria_b200_ce_bidir_bwd(m.data_ptr(), m.shape[0], scale, row.data_ptr(), col.data_ptr(),
                                              g.data_ptr(), dm.data_ptr(), _stream(m)), "ce_bidir_bwd")
    return dm


@ce_bidir_bwd.register_fake
def _(m, scale, row, col, g):
    return torch.empty_like(m)


def _ce_setup(ctx, inputs, output):
    m, scale = inputs
    _, row, col = output
    ctx.save_for_backward(m, row, col)
    ctx.scale = scale
    ctx.set_materialize_grads(False)


def _ce_backward(c, g, drow, dcol):
    m, row, col = c.saved_tensors
    if g is None:
        return torch.zeros_like(m), None
    return ce_bidir_bwd(m, c.scale, row, col, g), None


ce_bidir_fwd.register_autograd(_ce_backward, setup_context=_ce_setup)


# ----------------------------------------------------------------------------------------------------------------
# stand-alone attention_fn (gloria_loss.py:19-63) and row-wise cosine_similarity (gloria_loss.py:11-16)
# ----------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("gloria_b200::attention_fwd", mutates_args=())
def attention_fwd(query: Tensor, ctx: Tensor, temp1: float) -> Tuple[Tensor, Tensor]:
    """query [B, D, L], ctx [B, D, S] (paired) -> weighted context [B, D, L], attention [B, L, S]."""
    _need_cuda(query, ctx)
    L = _lib.lib()
    query, ctx = _f32c(query), _f32c(ctx)
    B, D, Lq = query.shape
    S = ctx.shape[2]
    if ctx.shape[0] != B or ctx.shape[1] != D:
        raise RuntimeError(f"attention_fn: query {tuple(query.shape)} and context {tuple(ctx.shape)} do not pair up")
    wctx = torch.empty((B, D, Lq), dtype=torch.float32, device=query.device)
    attn = torch.empty((B, Lq, S), dtype=torch.float32, device=query.device)
    with torch.cuda.device(query.device):
        n = L.gloria_b200_attention_workspace(B, D, S, Lq)
        ws = torch.empty((n,), dtype=torch.uint8, device=query.device)
        _lib.check(L.gloria_b200_attention_fwd_f32(query.data_ptr(), ctx.data_ptr(), B, D, S, Lq, temp1,
                                                   wctx.data_ptr(), attn.data_ptr(), ws.data_ptr(), n,
                                                   _stream(query)), "attention_fwd_f32")
    return wctx, attn


@attention_fwd.register_fake
def _(query, ctx, temp1):
    B, D, Lq = query.shape
    return query.new_empty((B, D, Lq)), query.new_empty((B, Lq, ctx.shape[2]))


@torch.library.custom_op("gloria_b200::attention_bwd", mutates_args=())
def attention_bwd(query: Tensor, ctx: Tensor, temp1: float, d_wctx: Optional[Tensor],
                  d_attn: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    _need_cuda(query, ctx)
    L = _lib.lib()
    query, ctx = _f32c(query), _f32c(ctx)
    B, D, Lq = query.shape
    S = ctx.shape[2]
    d_wctx = None if d_wctx is None else _f32c(d_wctx)
    d_attn = None if d_attn is None else _f32c(d_attn)
    dq, dc = torch.empty_like(query), torch.empty_like(ctx)
    with torch.cuda.device(query.device):
        n = L.gloria_b200_attention_workspace(B, D, S, Lq)
        ws = torch.empty((n,), dtype=torch.uint8, device=query.device)
        _lib.check(L.gloria_b200_attention_bwd_f32(query.data_ptr(), ctx.data_ptr(), B, D, S, Lq, temp1,
                                                   _ptr(d_wctx), _ptr(d_attn), dq.data_ptr(), dc.data_ptr(),
                                                   ws.data_ptr(), n, _stream(query)), "attention_bwd_f32")
    return dq, dc


@attention_bwd.register_fake
def _(query, ctx, temp1, d_wctx, d_attn):
    return torch.empty_like(query), torch.empty_like(ctx)


def _attn_setup(ctx, inputs, output):
    q, c, temp1 = inputs
    ctx.save_for_backward(q, c)
    ctx.temp1 = temp1
    ctx.set_materialize_grads(False)


def _attn_backward(c, d_wctx, d_attn):
    q, cx = c.saved_tensors
    if d_wctx is None and d_attn is None:
        return torch.zeros_like(q), torch.zeros_like(cx), None
    dq, dc = attention_bwd(q, cx, c.temp1, d_wctx, d_attn)
    return dq, dc, None


attention_fwd.register_autograd(_attn_backward, setup_context=_attn_setup)


@torch.library.custom_op("gloria_b200::row_cosine_fwd", mutates_args=())
def row_cosine_fwd(x1: Tensor, x2: Tensor, eps: float) -> Tuple[Tensor, Tensor]:
    """cos [N] of the rows of two [N, D] matrices, plus the saved (dot, |x1|, |x2|) per row."""
    _need_cuda(x1, x2)
    L = _lib.lib()
    x1, x2 = _f32c(x1), _f32c(x2)
    N, D = x1.shape
    out = torch.empty((N,), dtype=torch.float32, device=x1.device)
    stats = torch.empty((N, 3), dtype=torch.float32, device=x1.device)
    with torch.cuda.device(x1.device):
        _lib.check(L.gloria_b200_row_cosine_fwd(x1.data_ptr(), x2.data_ptr(), N, D, eps, out.data_ptr(),
                                                stats.data_ptr(), _stream(x1)), "row_cosine_fwd")
    return out, stats


@row_cosine_fwd.register_fake
def _(x1, x2, eps):
    return x1.new_empty((x1.shape[0],)), x1.new_empty((x1.shape[0], 3))


@torch.library.custom_op("gloria_b200::row_cosine_bwd", mutates_args=())
def row_cosine_bwd(x1: Tensor, x2: Tensor, stats: Tensor, g: Tensor, eps: float) -> Tuple[Tensor, Tensor]:
    _need_cuda(x1, x2, g)
    L = _lib.lib()
    x1, x2, g = _f32c(x1), _f32c(x2), _f32c(g)
    N, D = x1.shape
    d1, d2 = torch.empty_like(x1), torch.empty_like(x2)
    with torch.cuda.device(x1.device):
        _lib.check(L.gloria_b200_row_cosine_bwd(x1.data_ptr(), x2.data_ptr(), stats.data_ptr(), g.data_ptr(), N, D, eps,
                                                d1.data_ptr(), d2.data_ptr(), _stream(x1)), "row_cosine_bwd")
    return d1, d2


@row_cosine_bwd.register_fake
def _(x1, x2, stats, g, eps):
    return torch.empty_like(x1), torch.empty_like(x2)


def _rc_setup(ctx, inputs, output):
    x1, x2, eps = inputs
    ctx.save_for_backward(x1, x2, output[1])
    ctx.eps = eps
    ctx.set_materialize_grads(False)


def _rc_backward(c, g, gstats):
    x1, x2, stats = c.saved_tensors
    if g is None:
        return torch.zeros_like(x1), torch.zeros_like(x2), None
    d1, d2 = row_cosine_bwd(x1, x2, stats, g, c.eps)
    return d1, d2, None


row_cosine_fwd.register_autograd(_rc_backward, setup_context=_rc_setup)


# ----------------------------------------------------------------------------------------------------------------
# word-piece aggregation of the text encoder (text_model.py:32-90) -- the step right before the loss path
# ----------------------------------------------------------------------------------------------------------------
_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def word_ranges(caption_ids: Tensor, is_continuation: Tensor, sep_id: int, is_bracket: Optional[Tensor] = None):
    """caption_ids [B, T] int64, is_continuation [vocab] uint8 (1 = the entry starts with "##") ->
    word_range [B, T, 2] int32, token_word [B, T] int32, n_words [B] int32 (all on the device, no sync).
    With is_bracket [vocab] uint8 (1 = the entry's text starts with '[') a fourth tensor follows: cap_lens [B] int32,
    the caption lengths of gloria_model.py:107-109."""
    _need_cuda(caption_ids, is_continuation)
    L = _lib.lib()
    ids = caption_ids.to(torch.int64).contiguous()
    B, T = ids.shape
    dev = ids.device
    wr = torch.empty((B, T, 2), dtype=torch.int32, device=dev)
    tw = torch.empty((B, T), dtype=torch.int32, device=dev)
    nw = torch.empty((B,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        if is_bracket is None:
            rc = L.gloria_b200_word_ranges(ids.data_ptr(), is_continuation.data_ptr(), is_continuation.numel(),
                                           int(sep_id), B, T, wr.data_ptr(), tw.data_ptr(), nw.data_ptr(), _stream(ids))
            _lib.check(rc, "word_ranges")
            return wr, tw, nw
        cl = torch.empty((B,), dtype=torch.int32, device=dev)
        rc = L.gloria_b200_word_ranges_cap_lens(ids.data_ptr(), is_continuation.data_ptr(), is_bracket.data_ptr(),
                                                is_continuation.numel(), int(sep_id), B, T, wr.data_ptr(), tw.data_ptr(),
                                                nw.data_ptr(), cl.data_ptr(), _stream(ids))
    _lib.check(rc, "word_ranges_cap_lens")
    return wr, tw, nw, cl


@torch.library.custom_op("gloria_b200::aggregate_tokens", mutates_args=())
def aggregate_tokens(embeddings: Tensor, word_range: Tensor, token_word: Tensor) -> Tensor:
    """out[b, layer, w, :] = sum of embeddings[b, layer, t, :] over the tokens of word w ([B, layers, T, D], zero rows
    beyond each caption's words).  token_word is only carried for the backward."""
    _need_cuda(embeddings, word_range, token_word)
    if embeddings.dtype not in _DTYPES:
        raise RuntimeError(f"aggregate_tokens: unsupported dtype {embeddings.dtype}")
    L = _lib.lib()
    emb = embeddings.contiguous()
    B, layers, T, D = emb.shape
    out = torch.empty_like(emb)
    with torch.cuda.device(emb.device):
        rc = L.gloria_b200_aggregate_tokens_fwd(emb.data_ptr(), _DTYPES[emb.dtype], word_range.data_ptr(), B, layers, T, D,
                                                out.data_ptr(), _stream(emb))
    _lib.check(rc, "aggregate_tokens_fwd")
    return out


@aggregate_tokens.register_fake
def _(embeddings, word_range, token_word):
    return torch.empty_like(embeddings)


@torch.library.custom_op("gloria_b200::aggregate_tokens_bwd", mutates_args=())
def aggregate_tokens_bwd(d_out: Tensor, token_word: Tensor) -> Tensor:
    _need_cuda(d_out, token_word)
    L = _lib.lib()
    g = d_out.contiguous()
    B, layers, T, D = g.shape
    d_emb = torch.empty_like(g)
    with torch.cuda.device(g.device):
        rc = L.gloria_b200_aggregate_tokens_bwd(g.data_ptr(), _DTYPES[g.dtype], token_word.data_ptr(), B, layers, T, D,
                                                d_emb.data_ptr(), _stream(g))
    _lib.check(rc, "aggregate_tokens_bwd")
    return d_emb


@aggregate_tokens_bwd.register_fake
def _(d_out, token_word):
    return torch.empty_like(d_out)


def _agg_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[2])


def _agg_backward(ctx, d_out):
    (token_word,) = ctx.saved_tensors
    return aggregate_tokens_bwd(d_out, token_word), None, None


aggregate_tokens.register_autograd(_agg_backward, setup_context=_agg_setup)
