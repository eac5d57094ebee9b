"""ctypes binding of libgloria_b200.so (the C ABI declared in include/gloria_b200.h).

There is deliberately no CPU or PyTorch fallback: if the library is missing or a call fails, a RuntimeError is
raised (the reference surfaces faults as RuntimeError too).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from .build import LIB_PATH

_lock = threading.Lock()
_lib = None

_i, _f, _p, _z = C.c_int, C.c_float, C.c_void_p, C.c_size_t

# name -> (restype, argtypes); mirrors include/gloria_b200.h one to one
PROTOTYPES = {
    "gloria_b200_version": (_i, []),
    "gloria_b200_build_id": (C.c_char_p, []),
    "gloria_b200_last_error": (C.c_char_p, []),
    "gloria_b200_launch_count": (C.c_longlong, [_i]),
    "gloria_b200_set_timer_events": (_i, [_i, _p, _p]),
    "gloria_b200_record_event": (_i, [_p, _p]),
    "gloria_b200_debug_phase_clocks": (None, [_p]),
    "gloria_b200_upload_ints": (_i, [_p, _i, _p, _p]),
    "gloria_b200_local_f32_workspace": (_z, [_i, _i, _i, _i, _i, _i, _z]),
    "gloria_b200_local_sim_fwd_f32": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _f, _i, _f, _p, _p, _p, _p, _z,
                                           _p]),
    "gloria_b200_local_sim_bwd_f32": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _f, _i, _f, _p, _p, _p, _p, _p,
                                           _p, _z, _p]),
    "gloria_b200_f32tc_supported": (_i, [_i, _i, _i]),
    "gloria_b200_local_f32tc_workspace": (_z, [_i, _i, _i, _i, _i, _i, _z, _i]),
    "gloria_b200_local_sim_fwd_f32tc": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _f, _i, _f, _p, _p, _p, _p, _z,
                                             _p]),
    "gloria_b200_local_sim_fwd_f32tc_train": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _f, _i, _f, _p, _p, _p, _p,
                                                   _z, _p]),
    "gloria_b200_local_sim_bwd_f32tc": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _f, _i, _f, _p, _p, _p, _p, _p,
                                             _p, _z, _i, _p]),
    "gloria_b200_diag_attn_workspace": (_z, [_i, _i, _i, _i, _i]),
    "gloria_b200_diag_attn_fwd_f32": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _p, _p, _z, _p]),
    "gloria_b200_diag_attn_bwd_f32": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _p, _p, _p, _i, _p, _z, _p]),
    "gloria_b200_attention_workspace": (_z, [_i, _i, _i, _i]),
    "gloria_b200_attention_fwd_f32": (_i, [_p, _p, _i, _i, _i, _i, _f, _p, _p, _p, _z, _p]),
    "gloria_b200_attention_bwd_f32": (_i, [_p, _p, _i, _i, _i, _i, _f, _p, _p, _p, _p, _p, _z, _p]),
    "gloria_b200_row_cosine_fwd": (_i, [_p, _p, C.c_longlong, _i, _f, _p, _p, _p]),
    "gloria_b200_row_cosine_bwd": (_i, [_p, _p, _p, _p, C.c_longlong, _i, _f, _p, _p, _p]),
    "gloria_b200_tc_spad": (_i, [_i]),
    "gloria_b200_tc_lpad": (_i, [_i]),
    "gloria_b200_tc_lp": (_i, [_i]),
    "gloria_b200_tc_sp": (_i, [_i]),
    "gloria_b200_tc_supported": (_i, [_i, _i, _i]),
    "gloria_b200_tc_prepack": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "gloria_b200_tc_local_sim_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _f, _i, _f, _p, _p, _p]),
    "gloria_b200_tc_packed_groups": (_i, [_i]),
    "gloria_b200_tc_packed_per": (_i, [_i]),
    "gloria_b200_tc_prepack_words_packed": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "gloria_b200_tc_local_sim_fwd_packed": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _f, _i, _f, _p, _p]),
    "gloria_b200_tc_bwd_workspace": (_z, [_i, _i, _i, _i, _i, _i, _z]),
    "gloria_b200_tc_local_sim_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _f, _i, _f,
                                          _p, _p, _p, _p, _p, _z, _p]),
    "gloria_b200_tc_mean_workspace": (_z, [_i, _i, _i, _i, _i]),
    "gloria_b200_tc_local_sim_fwd_mean": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _f, _i, _f, _p, _p, _p, _p, _z,
                                               _p]),
    "gloria_b200_tc_train_workspace": (_z, [_i, _i, _i, _i, _i]),
    "gloria_b200_tc_local_sim_fwd_train": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _f, _i, _f, _p, _p, _z, _p]),
    "gloria_b200_tc_local_sim_bwd_train": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _z, _p]),
    "gloria_b200_tc_prepack_ctx": (_i, [_p, _i, _i, _i, _p, _p, _p, _p]),
    "gloria_b200_tc_prepack_words": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "gloria_b200_tc_local_sim_fwd_train_part": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _f, _i, _f, _p,
                                                     _p, _z, _p]),
    "gloria_b200_tc_local_sim_fwd_train_range": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _f, _i, _f, _p,
                                                      _p, _z, _i, _p, _i, _p]),
    "gloria_b200_tc_local_sim_fwd_train_diag": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _f, _i, _f, _p, _p, _z,
                                                     _p, _i, _p]),
    "gloria_b200_tc_local_sim_bwd_train_parts": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _z, _i,
                                                      _p, _p]),
    "gloria_b200_tc_train_drt_offset": (_z, [_i, _i, _i, _i, _i]),
    "gloria_b200_tc_unpack_dctx": (_i, [_p, _p, _i, _i, _i, _p]),
    "gloria_b200_tc_local_sim_bwd_train_ev": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _z, _p, _p]),
    "gloria_b200_acc_gemm": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _i, _i, _i, _i, _i, _p]),
    "gloria_b200_acc_gemm_planes": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "gloria_b200_word_ranges": (_i, [_p, _p, _i, C.c_longlong, _i, _i, _p, _p, _p, _p]),
    "gloria_b200_word_ranges_cap_lens": (_i, [_p, _p, _p, _i, C.c_longlong, _i, _i, _p, _p, _p, _p, _p]),
    "gloria_b200_aggregate_tokens_fwd": (_i, [_p, _i, _p, _i, _i, _i, _i, _p, _p]),
    "gloria_b200_aggregate_tokens_bwd": (_i, [_p, _i, _p, _i, _i, _i, _i, _p, _p]),
    "gloria_b200_global_sim_fwd": (_i, [_p, _p, _i, _i, _i, _f, _p, _p, _p, _p]),
    "gloria_b200_global_sim_bwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _f, _p, _p, _p]),
    "gloria_b200_ce_bidir_fwd": (_i, [_p, _i, _f, _p, _p, _p, _p]),
    "gloria_b200_ce_bidir_bwd": (_i, [_p, _i, _f, _p, _p, _p, _p, _p]),
}


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises RuntimeError if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                path = os.environ.get("GLORIA_B200_LIB", LIB_PATH)      # development: A/B two builds on one box
                if not os.path.exists(path):
                    raise RuntimeError(
                        f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a). There is no CPU fallback for this path.")
                h = C.CDLL(path)
                for name, (res, args) in PROTOTYPES.items():
                    fn = getattr(h, name)      # AttributeError if a declared symbol is not exported
                    fn.restype, fn.argtypes = res, args
                if "GLORIA_B200_LIB" not in os.environ and os.environ.get("GLORIA_B200_SKIP_ID_CHECK") != "1":
                    from .build import source_id
                    built, now = h.gloria_b200_build_id().decode(), source_id()
                    if built != now:
                        raise RuntimeError(
                            f"{path} was built from other sources (library id {built}, sources on disk {now}): rebuild "
                            "with `python -c 'import __graft_entry__ as g; g.build()'`")
                _lib = h
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().gloria_b200_last_error().decode(errors="replace")
        raise RuntimeError(f"libgloria_b200: {what} failed with status {rc}: {msg}")
