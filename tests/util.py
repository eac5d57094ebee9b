"""Shared helpers for the parity tests."""
import numpy as np
import torch


def relerr(a, b):
    """max-norm relative error of a vs the trusted b."""
    a = np.asarray(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a, dtype=np.float64)
    b = np.asarray(b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.all(np.isfinite(a)), "non-finite values in result"
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def cu(x, grad=False, dtype=torch.float32):
    t = torch.tensor(np.asarray(x), dtype=dtype, device="cuda")
    return t.requires_grad_(grad)


class Holder:
    """Stand-in for the reference GLoRIA nn.Module: just the attributes its __init__ sets (gloria_model.py:60-75)."""

    def __init__(self, **kw):
        d = dict(local_loss_weight=1.0, global_loss_weight=1.0, sparse_attn_loss_weight=None,
                 no_attn_loss_weight=None, attention_divergence_loss_weight=None, attention_entropy_loss_weight=None,
                 segmentation_loss_weight=None, temp1=4.0, temp2=5.0, temp3=10.0, no_attn_vec=None)
        d.update(kw)
        self.__dict__.update(d)
