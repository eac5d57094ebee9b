"""Drop-in for the word-piece aggregation of the reference text encoder
(gloria/models/text_model.py:32-90, `BertEncoder.aggregate_tokens`) and for the caption lengths the loss derives from
its sentences (gloria/models/gloria_model.py:107-109).

The reference walks every token of every caption in Python with `word_id.item()` -- B x T device syncs per step -- and
stacks / sums the word pieces one word at a time.  Here the token ids are read back once (B x T integers, needed anyway
for the returned word strings), the word boundaries are derived on the device from a "starts with ##" vocabulary table,
and the aggregation is one streaming kernel with autograd (`ops.aggregate_tokens`).  SURVEY.md section 8f, row 3.

    from gloria.models.text_model import BertEncoder
    from gloria_nlp_project_b200.text_model import patch_bert_encoder
    patch_bert_encoder(BertEncoder)          # aggregate_tokens now runs on the device
"""
from __future__ import annotations

from collections.abc import Sequence as _SequenceABC
from typing import Dict, List, Sequence, Tuple

import torch

from . import gloria_loss, ops

__all__ = ["aggregate_tokens", "AggregateTokensMixin", "patch_bert_encoder", "cap_lens_from_sents", "VocabTable",
           "LazySentences"]


class VocabTable:
    """Device-side view of `idxtoword`: which ids continue a word ("##...") and which id is "[SEP]"."""

    def __init__(self, idxtoword: Dict[int, str]):
        self.idxtoword = idxtoword
        self.size = max(idxtoword) + 1
        cont = torch.zeros(self.size, dtype=torch.uint8)
        for i, w in idxtoword.items():
            if w.startswith("##"):
                cont[i] = 1
        self._cont_cpu = cont
        self._cont = {}
        seps = [i for i, w in idxtoword.items() if w == "[SEP]"]
        self.sep_id = seps[0] if seps else -1
        # host tables for the word strings: the piece text ("##" stripped) and the continuation flag per id
        self.piece = [""] * self.size
        self.cont = [False] * self.size
        brk = torch.zeros(self.size, dtype=torch.uint8)
        for i, w in idxtoword.items():
            c = w.startswith("##")
            self.cont[i] = c
            self.piece[i] = w[2:] if c else w
            if self.piece[i].startswith("["):
                brk[i] = 1
        self._brk_cpu = brk
        self._brk = {}

    def is_continuation(self, device: torch.device) -> torch.Tensor:
        key = str(device)
        if key not in self._cont:
            self._cont[key] = self._cont_cpu.to(device)
        return self._cont[key]

    def is_bracket(self, device: torch.device) -> torch.Tensor:
        """1 where a word that STARTS with this entry starts with '[' ([CLS], [SEP], [PAD], ...): what the caption
        length of gloria_model.py:107-109 skips."""
        key = str(device)
        if key not in self._brk:
            self._brk[key] = self._brk_cpu.to(device)
        return self._brk[key]


def _sentences(ids: Sequence[Sequence[int]], vocab) -> List[List[str]]:
    """The word strings of text_model.py:43-82 from the token ids (host side, no tensors involved).  `vocab` is a
    VocabTable or a plain idxtoword dict."""
    if not isinstance(vocab, VocabTable):
        vocab = VocabTable(vocab)
    piece, cont, sep = vocab.piece, vocab.cont, vocab.sep_id
    out = []
    for row in ids:
        n_tok = len(row)
        try:
            p = row.index(sep)
        except ValueError:
            p = -1
        words, cur = [], None
        for v in (row[:p] if p >= 0 else row):
            if cont[v] and cur is not None:
                cur.append(piece[v])
            else:
                if cur is not None:
                    words.append("".join(cur))
                cur = [piece[v]]
        if p >= 0:                                   # [SEP] closes the open word and is a word of its own (:50-58);
            words.append("".join(cur) if cur is not None else "")
            words.append("[SEP]")                    # without it the word still open is never emitted
        out.append(words + ["[PAD]"] * (n_tok - len(words)))
    return out


class LazySentences(_SequenceABC):
    """The `sentences` BertEncoder.aggregate_tokens returns (a list of word-string lists), built only when somebody
    reads them (logging, attention plots): the training step needs just the caption lengths, and those arrive as
    `cap_lens` -- a `gloria_loss.DeviceCapLens` computed by the word-boundary kernel -- so a step involves no read-back
    of the token ids and no Python string work at all (5.9 ms of host time per step at B = 512 before)."""

    def __init__(self, caption_ids: torch.Tensor, vocab: "VocabTable", cap_lens: torch.Tensor, n_words: torch.Tensor):
        self._ids, self._vocab, self._built = caption_ids, vocab, None
        self.cap_lens = gloria_loss.DeviceCapLens(cap_lens)
        self.n_words = n_words

    def _materialise(self) -> List[List[str]]:
        if self._built is None:
            self._built = _sentences(self._ids.tolist(), self._vocab)   # one read-back instead of B x T .item() syncs
        return self._built

    def __len__(self):
        return self._ids.shape[0]

    def __getitem__(self, i):
        return self._materialise()[i]

    def __iter__(self):
        return iter(self._materialise())

    def __eq__(self, other):
        return list(self) == list(other)

    def __repr__(self):
        return repr(self._materialise())


def aggregate_tokens(embeddings: torch.Tensor, caption_ids: torch.Tensor, vocab: VocabTable
                     ) -> Tuple[torch.Tensor, "LazySentences"]:
    """embeddings [B, layers, T, D] (CUDA; fp32 / fp16 / bf16), caption_ids [B, T] -> (aggregated [B, layers, T, D],
    sentences) as BertEncoder.aggregate_tokens returns them; differentiable w.r.t. the embeddings.  `sentences` behaves
    like the reference's list of word lists but is built lazily, and carries `.cap_lens` for the loss (see
    LazySentences)."""
    if not embeddings.is_cuda:
        raise RuntimeError("gloria_b200 runs on CUDA tensors only (sm_100a); there is no CPU fallback")
    dev = embeddings.device
    ids_dev = caption_ids.to(dev, non_blocking=True)
    wr, tw, nw, cl = ops.word_ranges(ids_dev, vocab.is_continuation(dev), vocab.sep_id, vocab.is_bracket(dev))
    agg = ops.aggregate_tokens(embeddings, wr, tw)
    return agg, LazySentences(caption_ids, vocab, cl, nw)


def cap_lens_from_sents(sents):
    """gloria_model.py:107-109: words not starting with '[' plus one (the device tensor of a LazySentences as it is)."""
    dev = getattr(sents, "cap_lens", None)
    if isinstance(dev, gloria_loss.DeviceCapLens):
        return dev
    return [len([w for w in sent if not w.startswith("[")]) + 1 for sent in sents]


class AggregateTokensMixin:
    """Place before the reference BertEncoder in an MRO, or use patch_bert_encoder()."""

    def aggregate_tokens(self, embeddings, caption_ids):
        table = getattr(self, "_gloria_b200_vocab", None)
        if table is None or table.idxtoword is not self.idxtoword:
            table = VocabTable(self.idxtoword)
            object.__setattr__(self, "_gloria_b200_vocab", table)
        return aggregate_tokens(embeddings, caption_ids, table)


def patch_bert_encoder(target):
    """Monkey-patch the reference BertEncoder class (or an instance) so aggregate_tokens runs on the device."""
    import types
    fn = AggregateTokensMixin.aggregate_tokens
    if isinstance(target, type):
        target.aggregate_tokens = fn
    else:
        object.__setattr__(target, "aggregate_tokens", types.MethodType(fn, target))
    return target
