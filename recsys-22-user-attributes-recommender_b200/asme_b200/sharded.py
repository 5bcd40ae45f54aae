"""Vocab-sharded full-catalog evaluation across the GPUs of one box (new in this build -- the reference
has no sharded scoring, its only multi-GPU mode is Lightning DDP).

The item table (and the output bias) is row-sharded: rank g owns the catalog slice [v0, v1) = shard_range(V, G, g).
One evaluation step for a global batch of users:

  1. all-gather the (B_local, H) hidden rows and the targets        -> every rank holds all B = G * B_local users
  2. local:  tcgen05 scoring of all B users against the own slice   -> top-k (score, id) per user, score of the target
             (written only by the rank that owns the target's row)
  3. all-reduce(SUM) the target scores  [B floats],  all-gather the per-shard top-k lists  [G, B, k, 2 x 4 bytes]
  4. local:  K-way merge of the G lists with (score desc, id asc) order; rank = position of the target in the merged list
             (or, for the ``rank`` / full ``MRR`` metrics, a count-only sweep over the slice + all-reduce(SUM) of the counts)
  5. every rank keeps the rows of its own users; metric sums are all-reduced once per epoch (RankingMetric.sync)

Every message is a few hundred KB at most, so the exchange is latency-bound; it is kept to three NCCL calls per step.
The collective plumbing is plain ``torch.distributed`` and is exercised on CPU with the gloo backend (tests/
test_sharded_cpu.py) by injecting the per-shard scorer; on the GPU the scorer is :func:`tc_local_scorer`.
"""
from typing import Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(V: int, G: int, g: int) -> Tuple[int, int]:
    """rows [v0, v1) of rank g: ceil(V / G) rows per rank, the last shard may be shorter"""
    per = (V + G - 1) // G
    return min(V, g * per), min(V, (g + 1) * per)


def _world(group) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def all_gather_rows(x: torch.Tensor, group=None) -> torch.Tensor:
    """(B_local, ...) -> (G * B_local, ...) in rank order"""
    G, _ = _world(group)
    if G == 1:
        return x
    out = torch.empty((G * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def merge_topk_host(vals: torch.Tensor, idx: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """reference merge of (G,B,k) lists used by the CPU tests: (score desc, id asc); empty slots carry id -1"""
    G, B, kk = vals.shape
    v = vals.permute(1, 0, 2).reshape(B, G * kk).to(torch.float64)
    i = idx.permute(1, 0, 2).reshape(B, G * kk).to(torch.int64)
    v = torch.where(i < 0, torch.full_like(v, float("-inf")), v)
    key_id = torch.where(i < 0, torch.full_like(i, 2 ** 40), i)
    order = torch.argsort(key_id, dim=1, stable=True)                       # secondary key first ...
    v, i = torch.gather(v, 1, order), torch.gather(i, 1, order)
    order = torch.argsort(-v, dim=1, stable=True)                           # ... then stable sort by the primary key
    v, i = torch.gather(v, 1, order)[:, :k], torch.gather(i, 1, order)[:, :k]
    return v.to(torch.float32), i.to(torch.int32)


def sharded_topk_rank(hidden_local: torch.Tensor, target_local: torch.Tensor, k: int,
                      local_scorer: Callable[..., Dict[str, Optional[torch.Tensor]]],
                      merge: Callable[[torch.Tensor, torch.Tensor, int], Tuple[torch.Tensor, torch.Tensor]],
                      full_rank: bool = False, group=None) -> Dict[str, torch.Tensor]:
    """steps 1-5 above.  ``local_scorer(hidden_all, target_all, k, target_score_in)`` scores all users against the own slice
    and returns dict(topk_val (B,k), topk_idx (B,k) int32 global ids, target_score (B) [owner rows, 0 elsewhere],
    n_greater, n_tie_lower (B) int32 when target_score_in is given)."""
    G, g = _world(group)
    B_local = hidden_local.shape[0]
    hidden_all = all_gather_rows(hidden_local, group)
    target_all = all_gather_rows(target_local, group)
    part = local_scorer(hidden_all, target_all, k, None)
    ts = part["target_score"].clone()
    if G > 1:
        dist.all_reduce(ts, op=dist.ReduceOp.SUM, group=group)
        packed = torch.stack([part["topk_val"], part["topk_idx"].view(torch.float32)], dim=-1).contiguous()     # (B,k,2)
        gathered = torch.empty((G * packed.shape[0],) + tuple(packed.shape[1:]), dtype=packed.dtype, device=packed.device)
        dist.all_gather_into_tensor(gathered, packed, group=group)          # concatenated along dim 0 (gloo and NCCL both accept it)
        gathered = gathered.view((G,) + tuple(packed.shape))
        vals, idx = gathered[..., 0].contiguous(), gathered[..., 1].contiguous().view(torch.int32)
        val, ids = merge(vals, idx, k)
    else:
        val, ids = part["topk_val"], part["topk_idx"]
    sl = slice(g * B_local, (g + 1) * B_local)
    out = dict(topk_val=val[sl], topk_idx=ids[sl], target_score=ts[sl])
    if full_rank:
        cnt = local_scorer(hidden_all, target_all, 0, ts)
        counts = torch.stack([cnt["n_greater"], cnt["n_tie_lower"]]).to(torch.int32)
        if G > 1:
            dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
        out["rank"] = (counts[0] + counts[1] + 1)[sl].to(torch.int32)
    else:
        hit = out["topk_idx"].eq(target_local.to(torch.int32).unsqueeze(1))
        pos = hit.to(torch.int32).argmax(dim=1).to(torch.int32)
        out["rank"] = torch.where(hit.any(dim=1), pos + 1, torch.full_like(pos, k + 1))
    return out


def tc_local_scorer(wb_shard: torch.Tensor, bias_shard: Optional[torch.Tensor], v0: int, folded: bool = False):
    """the production per-shard scorer: tcgen05 scoring kernel over the rank's (Vloc, Kp) bf16 slice; ``folded``: the slice
    carries the bias in two extra K columns (models.projection_operands_folded) and the hidden rows get the matching ones"""
    from . import ops

    def score(hidden_all, target_all, k, target_score_in):
        hb = ops.cast_bf16_ext(hidden_all) if folded else ops.cast_bf16(hidden_all, ld_out=wb_shard.shape[1])
        return ops.tc_score_topk(hb, wb_shard, bias_shard, k, target=target_all, target_score_in=target_score_in, v0=v0,
                                 capture_target=target_score_in is None)

    return score


def tc_merge(vals: torch.Tensor, idx: torch.Tensor, k: int):
    from . import ops
    return ops.topk_merge(vals, idx, k)
