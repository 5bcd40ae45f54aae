"""Vocab-sharded full-catalog scoring across the GPUs of one box (new in this build -- the reference has no
sharded scoring, its only multi-GPU mode is Lightning DDP).

The item table (and the output bias) is row-sharded: rank g owns the catalog slice [v0, v1) = shard_range(V, G, g).

Evaluation step for a global batch of users (:func:`sharded_topk_rank`), three NCCL calls:

  1. ONE all-gather of [hidden row | target] per user                -> every rank holds all B = G * B_local users
  2. local: tcgen05 sweep of all B users over the own slice           -> per user the slice's exact top-k (score, id), the target's
            score (only the owner's is non-zero), optionally the slice's (row max, sum-exp) for the validation loss
  3. ONE all-to-all: rank r receives, from every shard, the results of ITS OWN B_local users only (G lists per user instead
            of G * B: each rank merges 1/G of what an all-gather would make it merge)
  4. local: G-way merge (score desc, id asc); target score = sum of the G contributions; rank = position of the target in the
            merged list, or -- for the ``rank`` / full ``MRR`` metrics -- a count-only sweep and a reduce-scatter(SUM) of the counts;
            loss from the merged (max, sum-exp)
  5. metric sums are all-reduced once per epoch (RankingMetric.sync)

Training with a sharded scoring layer (:func:`sharded_ce`): every rank scores the SAME rows against its slice; the softmax
statistics meet in all-reduce(MAX) of the row maxima and ONE packed all-reduce(SUM) of (rescaled sum-exp, target logit);
the backward all-reduces dH (every shard contributes its slice's part), dW / dbias stay local to the slice.

Every message is a few hundred KB at most: the exchange is latency-bound, so calls are packed and counted.  The collective plumbing
is plain ``torch.distributed`` (NCCL on the GPUs, captured into the evaluation graph; gloo in tests/test_sharded_cpu.py, where the
per-shard scorer is injected).
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


def all_to_all_rows(x: torch.Tensor, group=None) -> torch.Tensor:
    """x (G, n, ...): block g goes to rank g; returns (G, n, ...) whose block g came from rank g"""
    G, _ = _world(group)
    if G == 1:
        return x
    x = x.contiguous()
    out = torch.empty_like(x)
    try:
        dist.all_to_all_single(out, x, group=group)
    except (RuntimeError, NotImplementedError):          # a backend without all-to-all (CPU tests): all-gather and keep the own blocks
        _, g = _world(group)
        full = torch.empty((G,) + tuple(x.shape), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(full.view(G * x.shape[0], *x.shape[1:]), x, group=group)
        out = full[:, g].contiguous()
    return out


def reduce_scatter_rows(x: torch.Tensor, group=None) -> torch.Tensor:
    """x (G * n, ...) summed over the ranks; rank g keeps rows [g n, (g+1) n)"""
    G, g = _world(group)
    if G == 1:
        return x
    n = x.shape[0] // G
    x = x.contiguous()
    if dist.get_backend(group) == "gloo":                # gloo has no reduce-scatter
        dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)
        return x[g * n:(g + 1) * n].contiguous()
    out = torch.empty((n,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.reduce_scatter_tensor(out, x, op=dist.ReduceOp.SUM, group=group)
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


def combine_ce_host(rmax: torch.Tensor, rsum: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(G, B) per-shard row maxima / sum-exps -> (B) global ones (CPU tests; the GPU path uses ops.ce_combine)"""
    m = rmax.max(dim=0).values
    s = (rsum * torch.exp(rmax - m.unsqueeze(0))).sum(dim=0)
    return m, s


def _pack_users(hidden_local: torch.Tensor, target_local: torch.Tensor) -> torch.Tensor:
    """[hidden row | target] as one fp32 row per user: the int64 target rides in two fp32 slots (bit pattern, not value)"""
    t = target_local.to(torch.int64).contiguous().view(torch.float32).view(-1, 2)
    return torch.cat([hidden_local, t.to(hidden_local.device)], dim=1).contiguous()


def _unpack_users(packed: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    H = packed.shape[1] - 2
    return packed[:, :H].contiguous(), packed[:, H:].contiguous().view(torch.int64).view(-1)


def sharded_topk_rank(hidden_local: torch.Tensor, target_local: torch.Tensor, k: int,
                      local_scorer: Callable[..., Dict[str, Optional[torch.Tensor]]],
                      merge: Callable[[torch.Tensor, torch.Tensor, int], Tuple[torch.Tensor, torch.Tensor]],
                      full_rank: bool = False, group=None, with_loss: bool = False, pad_id: int = 0,
                      combine_ce: Callable = combine_ce_host) -> Dict[str, torch.Tensor]:
    """steps 1-4 above.  ``local_scorer(hidden_all, target_all, k, target_score_in, want_ce=...)`` scores all users against the
    own slice and returns dict(topk_val (B,k), topk_idx (B,k) int32 global ids, target_score (B) [owner rows, 0 elsewhere],
    pivot (B) [what a later count call compares against; defaults to target_score], rmax / rsum (/ tl) (B) when ``want_ce``,
    n_greater / n_tie_lower (B) int32 when ``target_score_in`` is given)."""
    G, g = _world(group)
    B_local = hidden_local.shape[0]
    if G > 1:
        hidden_all, target_all = _unpack_users(all_gather_rows(_pack_users(hidden_local, target_local), group))
    else:
        hidden_all, target_all = hidden_local, target_local
    part = local_scorer(hidden_all, target_all, k, None, want_ce=with_loss)
    n_extra = 4 if with_loss else 1
    if G > 1:
        B = G * B_local
        fields = [part["topk_val"], part["topk_idx"].view(torch.float32), part["target_score"].unsqueeze(1)]
        if with_loss:
            fields += [part["rmax"].unsqueeze(1), part["rsum"].unsqueeze(1), part.get("tl", part["target_score"]).unsqueeze(1)]
        packed = torch.cat(fields, dim=1).view(G, B_local, 2 * k + n_extra)          # block r = what this shard found for rank r's users
        got = all_to_all_rows(packed, group)                                         # block s = what shard s found for MY users
        vals = got[..., :k].contiguous()
        idx = got[..., k:2 * k].contiguous().view(torch.int32)
        val, ids = merge(vals, idx, k)
        ts = got[..., 2 * k].sum(dim=0)                                              # one owner, G - 1 zeros
        if with_loss:
            rmax, rsum = combine_ce(got[..., 2 * k + 1].contiguous(), got[..., 2 * k + 2].contiguous())
            tl = got[..., 2 * k + 3].sum(dim=0)                                      # the target's logit as the loss sweep computed it
    else:
        val, ids, ts = part["topk_val"], part["topk_idx"], part["target_score"]
        if with_loss:
            rmax, rsum, tl = part["rmax"], part["rsum"], part.get("tl", part["target_score"])
    out = dict(topk_val=val, topk_idx=ids, target_score=ts)
    if "n_uncertified" in part and part["n_uncertified"] is not None:
        out["n_uncertified"] = part["n_uncertified"]
    if with_loss:      # nn.CrossEntropyLoss(ignore_index=pad) over this rank's users (masked_training_module.py:150)
        out["lse"] = rmax + torch.log(rsum)
        keep = target_local.ne(pad_id)
        out["loss"] = ((out["lse"] - tl) * keep).sum() / keep.sum()
    if full_rank:
        # the pivot every shard counts against must be known everywhere: the owner's value, summed over the shards
        pivot = part.get("pivot", part["target_score"]).clone()
        if G > 1:
            dist.all_reduce(pivot, op=dist.ReduceOp.SUM, group=group)
        cnt = local_scorer(hidden_all, target_all, 0, pivot, want_ce=False)
        counts = torch.stack([cnt["n_greater"], cnt["n_tie_lower"]], dim=1).to(torch.int32)            # (B, 2)
        counts = reduce_scatter_rows(counts, group)
        full = (counts[:, 0] + counts[:, 1] + 1).to(torch.int32)
        hit = ids.eq(target_local.to(torch.int32).unsqueeze(1))
        pos = hit.to(torch.int32).argmax(dim=1).to(torch.int32) + 1
        # inside the merged list the position IS the rank (exact lists); outside, the count sweep's
        out["rank"] = torch.where(hit.any(dim=1), pos, torch.clamp(full, min=k + 1))
    else:
        hit = ids.eq(target_local.to(torch.int32).unsqueeze(1))
        pos = hit.to(torch.int32).argmax(dim=1).to(torch.int32)
        out["rank"] = torch.where(hit.any(dim=1), pos + 1, torch.full_like(pos, k + 1))
    return out


def tc_local_scorer(wb_shard: torch.Tensor, bias_shard: Optional[torch.Tensor], v0: int, folded: bool = False,
                    exact: Optional[Tuple[torch.Tensor, Optional[torch.Tensor], torch.Tensor]] = None, bias_bounds=None):
    """the production per-shard scorer: tcgen05 scoring kernel over the rank's (Vloc, Kp) bf16 slice; ``folded``: the slice
    carries the bias in two extra K columns (models.projection_operands_folded) and the hidden rows get the matching ones.
    ``exact = (w32_shard, b32_shard, norm_bound)``: the slice's lists are made exact (csrc/rescore.cu) before they are exchanged,
    so the merged lists are the fp32 path's.  ``bias_bounds`` (ops.bias_chunk_bounds(bias_shard), exact lists without a count sweep
    only): the slice is the plain (Vloc, H) table and the bias is added per chunk where a score could pass the threshold."""
    from . import ops
    # candidates per user and slice: a slice of 1/G of the catalog holds ~k/G of a user's global top k, and the certificate only needs
    # spare entries beyond the slice's OWN k best -- 32 instead of 64 halves the fp32 re-scoring of G x B_local users per rank
    G = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    k_out = 32 if G >= 4 else 64

    def score(hidden_all, target_all, k, target_score_in, want_ce: bool = False):
        hb = ops.cast_bf16_ext(hidden_all) if folded else ops.cast_bf16(hidden_all, ld_out=wb_shard.shape[1])
        if target_score_in is not None:          # count-only sweep
            return ops.tc_score_topk(hb, wb_shard, bias_shard, 0, target=target_all, target_score_in=target_score_in, v0=v0,
                                     capture_target=False)
        if exact is None:
            out = ops.tc_score_topk(hb, wb_shard, bias_shard, k, target=target_all, v0=v0)
        else:
            w32, b32, nb = exact
            if bias_bounds is not None:
                c = ops.tc_score_candidates(hb, wb_shard, bias_shard, k, k_out, v0=v0, bias_bounds=bias_bounds)
            else:
                c = ops.tc_score_candidates(hb, wb_shard, bias_shard, k, k_out, target=target_all, v0=v0)
            r = ops.topk_rescore(hidden_all, w32, b32, c["cand_idx"], c["cand_val"], k, nb, target_all, v0=v0, want_rank=False,
                                 cand_bound=c["bound"])
            ops.score_topk_flagged(hidden_all, w32, b32, target_all, r["target_score"], k, r["row_flag"], r["topk_val"], r["topk_idx"],
                                   None, v0=v0)
            out = dict(topk_val=r["topk_val"], topk_idx=r["topk_idx"], target_score=r["target_score"], pivot=c["target_score"],
                       n_uncertified=r["n_flagged"])
        if want_ce:
            out["rmax"], out["rsum"], out["tl"] = ops.tc_score_ce_partial(hb, wb_shard, bias_shard, target_all, v0=v0)
        return out

    return score


def tc_merge(vals: torch.Tensor, idx: torch.Tensor, k: int):
    from . import ops
    return ops.topk_merge(vals, idx, k)


def tc_combine_ce(rmax: torch.Tensor, rsum: torch.Tensor):
    from . import ops
    return ops.ce_combine(rmax, rsum)


# ----------------------------------------------------------------------------------------------------------------------------
# vocab-sharded cross entropy (training): every rank holds the same R rows and its own slice of the scoring layer
# ----------------------------------------------------------------------------------------------------------------------------
def sharded_ce(partial: Callable[[], Tuple[torch.Tensor, torch.Tensor, torch.Tensor]], target: torch.Tensor, pad_id: int = 0,
               group=None, rescale: Optional[Callable] = None) -> Dict[str, torch.Tensor]:
    """``partial()`` -> this shard's (row max, row sum-exp, target logit [owner rows, 0 elsewhere]) over its catalog slice
    (``ops.tc_score_ce_partial(..., v0=v0)`` / ``ops.score_ce_partial``).  Two collectives: all-reduce(MAX) of the row maxima,
    then ONE packed all-reduce(SUM) of (sum-exp rescaled to the global maximum, target logit).  Returns dict(lse (R), target_logit
    (R), loss = mean of lse - target_logit over the rows with target != pad) -- what nn.CrossEntropyLoss(ignore_index=pad) gives
    over the whole catalog (modules/masked_training_module.py:93-111)."""
    G, _ = _world(group)
    rmax, rsum, tl = partial()
    if G > 1:
        gmax = rmax.clone()
        dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=group)
        scaled = rescale(rsum, rmax, gmax) if rescale is not None else rsum * torch.exp(rmax - gmax)
        packed = torch.stack([scaled, tl])
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        rmax, rsum, tl = gmax, packed[0], packed[1]
    lse = rmax + torch.log(rsum)
    keep = target.ne(pad_id)
    return dict(lse=lse, target_logit=tl, loss=((lse - tl) * keep).sum() / keep.sum(), n_rows=keep.sum())


def sharded_ce_backward(backward: Callable[[torch.Tensor], torch.Tensor], lse: torch.Tensor, group=None) -> torch.Tensor:
    """``backward(lse)`` -> this shard's contribution to dH (R, H) for the GLOBAL log-sum-exp (it also accumulates the slice's dW /
    dbias, which need no exchange: every row of the batch is present on every rank).  dH = all-reduce(SUM) over the shards."""
    G, _ = _world(group)
    dh = backward(lse)
    if G > 1:
        dist.all_reduce(dh, op=dist.ReduceOp.SUM, group=group)
    return dh
