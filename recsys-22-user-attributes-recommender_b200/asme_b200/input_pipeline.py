"""Batched GPU versions of the reference's per-sample dataset processors: the step just before the hot path.

    cloze_mask(batch, ...)        ClozeMaskProcessor      data/datasets/processors/cloze_mask.py:50-92
    pos_neg_sample(batch, ...)    PositiveNegativeSamplerProcessor  data/datasets/processors/pos_neg_sampler.py:41-114

Both take the right-padded int64 id tensors the reference's collate produces and return new batch dictionaries with the
reference's entry names.  Streams are counter-based (pure functions of seed / sequence / position): the distributions and
invariants are the reference's, the individual draws are not (the reference draws from Python's global generator)."""
import ctypes
from typing import Dict, Optional, Sequence

import torch

from . import _lib
from .ops import _i64, _p, _stream

ITEM_SEQ_ENTRY_NAME = "item"
TARGET_ENTRY_NAME = "item.target"
POSITIVE_SAMPLES_ENTRY_NAME = "positive_samples"
NEGATIVE_SAMPLES_ENTRY_NAME = "negative_samples"


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def cloze_mask(batch: Dict[str, torch.Tensor], vocab_sizes: Dict[str, int], mask_prob: float, only_last_item_mask_prob: float,
               seed: int, masking_targets: Optional[Sequence[str]] = None, mask_token_id: int = 1, pad_token_id: int = 0,
               mask_ids: Optional[Dict[str, int]] = None) -> Dict[str, torch.Tensor]:
    """returns a copy of ``batch`` whose masking targets are masked and that carries ``item.target`` (N,S)"""
    names = [ITEM_SEQ_ENTRY_NAME] + [n for n in (masking_targets or []) if n != ITEM_SEQ_ENTRY_NAME]
    ins = [_i64(batch[n]) for n in names]
    B, S = ins[0].shape
    for n, t in zip(names, ins):
        if tuple(t.shape) != (B, S):
            raise NotImplementedError(f"cloze_mask: feature {n} has shape {tuple(t.shape)}; (N,S) id sequences only")
    outs = [torch.empty_like(t) for t in ins]
    target = torch.empty_like(ins[0])
    mids = (ctypes.c_int64 * len(names))(*[int((mask_ids or {}).get(n, mask_token_id)) for n in names])
    voc = (ctypes.c_int64 * len(names))(*[int(vocab_sizes[n]) for n in names])
    _lib.call("asme_b200_cloze_mask", B, S, len(names), _ptr_array(ins), _ptr_array(outs), mids, voc, _p(target), int(pad_token_id),
              float(mask_prob), float(only_last_item_mask_prob), int(seed), _stream())
    out = dict(batch)
    for n, t in zip(names, outs):
        out[n] = t
    out[TARGET_ENTRY_NAME] = target
    return out


def pos_neg_sample(batch: Dict[str, torch.Tensor], item_vocab_size: int, seed: int, n_special_tokens: int = 3,
                   pad_token_id: int = 0) -> Dict[str, torch.Tensor]:
    """``item`` (N,S) -> ``item`` = seq[:-1], ``positive_samples`` = seq[1:], ``negative_samples`` (all (N,S-1))"""
    seq = _i64(batch[ITEM_SEQ_ENTRY_NAME])
    B, S1 = seq.shape
    x = torch.empty(B, S1 - 1, dtype=torch.int64, device=seq.device)
    pos, neg = torch.empty_like(x), torch.empty_like(x)
    _lib.call("asme_b200_pos_neg_sample", _p(seq), B, S1, int(item_vocab_size), int(n_special_tokens), int(pad_token_id), int(seed),
              _p(x), _p(pos), _p(neg), _stream())
    out = dict(batch)
    out[ITEM_SEQ_ENTRY_NAME], out[POSITIVE_SAMPLES_ENTRY_NAME], out[NEGATIVE_SAMPLES_ENTRY_NAME] = x, pos, neg
    return out
