"""B200-native GLoRIA local/global contrastive similarity + loss (drop-in for the reference's loss path).

Public surface (mirrors strongbeamsprout/gloria-nlp-project):
  gloria_loss   -- module substitutable for gloria/loss/gloria_loss.py
  gloria_model  -- GLoRIALossMixin / patch_gloria(): calc_loss, get_local_similarities, get_global_similarities,
                   get_attn_maps of gloria/models/gloria_model.py
  text_model    -- BertEncoder.aggregate_tokens of gloria/models/text_model.py (word-piece sums on the device)
  zero_shot     -- get_similarities / zero_shot_classification of gloria/gloria.py (one launch for all classes' prompts)
  distributed   -- caption-sharded loss across the GPUs of one node (NCCL)
  set_precision -- "fp32" (CUDA-core, 1e-5), "bf16" (tcgen05 tensor cores, 2e-3) or "auto"
"""
from ._config import get_precision, precision, set_precision  # noqa: F401

__all__ = ["set_precision", "get_precision", "precision", "gloria_loss", "gloria_model", "text_model", "zero_shot", "distributed"]
__version__ = "0.1.0"


def __getattr__(name):
    if name in ("gloria_loss", "gloria_model", "text_model", "zero_shot", "distributed", "ops"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
