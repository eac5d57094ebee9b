"""Arithmetic-mode selection for the local similarity kernels."""
from __future__ import annotations

import contextlib
import os
import threading

_state = threading.local()
_VALID = ("auto", "fp32", "bf16")
_default = os.environ.get("GLORIA_B200_PRECISION", "auto")


def set_precision(mode: str) -> None:
    """'fp32': CUDA-core kernels, logits within 1e-5 of the un-autocast reference.
    'bf16': tcgen05 tensor-core kernels (bf16 operands, fp32 accumulate/softmax), logits within 2e-3.
    'auto': bf16 under torch autocast or for half/bfloat16 inputs (the reference's AMP setting), else fp32."""
    if mode not in _VALID:
        raise ValueError(f"precision must be one of {_VALID}")
    _state.mode = mode


def get_precision() -> str:
    return getattr(_state, "mode", _default)


@contextlib.contextmanager
def precision(mode: str):
    old = get_precision()
    set_precision(mode)
    try:
        yield
    finally:
        set_precision(old)
