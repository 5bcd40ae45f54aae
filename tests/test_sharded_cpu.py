"""World-size-2 CPU tests (gloo) of the multi-GPU plumbing: vocab-sharded top-k / rank exchange (asme_b200.sharded) and
the metric-state all-reduce.  The per-shard scorer is injected: here it is the CPU oracle (test infrastructure), on the
GPU it is the tcgen05 scoring kernel (tests/test_gpu_tc.py covers the kernel itself, including a 4-shard merge)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "recsys-22-user-attributes-recommender_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_scorer(w_shard, b_shard, v0):
    from oracle import asme_oracle as O

    def score(hidden_all, target_all, k, target_score_in, want_ce=False):
        logits = (hidden_all.double() @ w_shard.double().t() + b_shard.double()).float().numpy()
        B, Vloc = logits.shape
        t = target_all.numpy() - v0
        own = (t >= 0) & (t < Vloc)
        out = dict(topk_val=None, topk_idx=None, target_score=None, n_greater=None, n_tie_lower=None)
        if k > 0:
            kk = min(k, Vloc)
            ids = O.topk_ids(logits, kk)
            val = np.take_along_axis(logits, ids, axis=1)
            pad = k - kk
            out["topk_val"] = torch.from_numpy(np.pad(val, ((0, 0), (0, pad)), constant_values=-np.inf))
            out["topk_idx"] = torch.from_numpy(np.pad(ids + v0, ((0, 0), (0, pad)), constant_values=-1).astype(np.int32))
        if target_score_in is None:
            ts = np.zeros(B, dtype=np.float32)
            ts[own] = logits[np.arange(B)[own], t[own]]
            out["target_score"] = torch.from_numpy(ts)
        else:
            st = target_score_in.numpy()[:, None]
            cols = np.arange(Vloc)[None, :] + v0
            out["n_greater"] = torch.from_numpy((logits > st).sum(1).astype(np.int32))
            out["n_tie_lower"] = torch.from_numpy(((logits == st) & (cols < target_all.numpy()[:, None])).sum(1).astype(np.int32))
        if want_ce:
            lt = torch.from_numpy(logits)
            out["rmax"] = lt.max(dim=1).values
            out["rsum"] = torch.exp(lt - out["rmax"].unsqueeze(1)).sum(dim=1)
        return out

    return score


def _worker(rank, world, port, V, H, B_local, k, seed, result_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from asme_b200 import sharded
    from asme_b200.metrics import build_metrics
    gen = torch.Generator().manual_seed(seed)
    w = torch.randint(-3, 4, (V, H), generator=gen).float()          # small integers: many exact ties
    b = torch.randint(-2, 3, (V,), generator=gen).float()
    h_all = torch.randint(-3, 4, (world * B_local, H), generator=gen).float()
    t_all = torch.randint(0, V, (world * B_local,), generator=gen)
    v0, v1 = sharded.shard_range(V, world, rank)
    sl = slice(rank * B_local, (rank + 1) * B_local)
    out = sharded.sharded_topk_rank(h_all[sl], t_all[sl], k, _oracle_scorer(w[v0:v1], b[v0:v1], v0), sharded.merge_topk_host,
                                    full_rank=True)
    lite = sharded.sharded_topk_rank(h_all[sl], t_all[sl], k, _oracle_scorer(w[v0:v1], b[v0:v1], v0), sharded.merge_topk_host,
                                     full_rank=False, with_loss=True, pad_id=0)
    # vocab-sharded cross entropy of a training step: every rank scores the SAME rows against its slice
    def partial():
        logits = torch.from_numpy((h_all.double() @ w[v0:v1].double().t() + b[v0:v1].double()).float().numpy())
        rmax = logits.max(dim=1).values
        rsum = torch.exp(logits - rmax.unsqueeze(1)).sum(dim=1)
        t = t_all - v0
        own = (t >= 0) & (t < v1 - v0)
        tl = torch.where(own, logits[torch.arange(len(t)), t.clamp(0, v1 - v0 - 1)], torch.zeros(len(t)))
        return rmax, rsum, tl
    ce = sharded.sharded_ce(partial, t_all, pad_id=0)

    def backward(lse):          # dH of this slice for dlogit = (softmax - onehot) / n_rows
        logits = h_all.double() @ w[v0:v1].double().t() + b[v0:v1].double()
        p = torch.exp(logits - lse.double().unsqueeze(1))
        t = t_all - v0
        own = (t >= 0) & (t < v1 - v0)
        p[torch.arange(len(t))[own], t[own]] -= 1.0
        p = p * t_all.ne(0).double().unsqueeze(1) / float(ce["n_rows"])
        return (p @ w[v0:v1].double()).float()
    dh = sharded.sharded_ce_backward(backward, ce["lse"])
    metrics = build_metrics({"recall": [1, 5], "ndcg": [5], "mrr": [5]})
    for c in metrics.containers:
        for m in c.metrics:
            m.update_from_ranks(out["rank"])
    metrics.sync()
    torch.save(dict(out=out, lite=lite, ce={k_: v for k_, v in ce.items()}, dh=dh, metrics={k_: float(v) for k_, v in metrics.compute().items()}, h=h_all, t=t_all, w=w, b=b),
               f"{result_path}.{rank}")
    dist.destroy_process_group()


@pytest.mark.parametrize("V,k", [(1001, 10), (37, 5)])
def test_sharded_topk_rank_two_ranks_gloo(tmp_path, V, k):
    from oracle import asme_oracle as O
    world, H, B_local = 2, 16, 9
    path = str(tmp_path / "res")
    mp.spawn(_worker, args=(world, _free_port(), V, H, B_local, k, 123, path), nprocs=world, join=True)
    res = [torch.load(f"{path}.{r}", weights_only=False) for r in range(world)]
    h, t, w, b = res[0]["h"], res[0]["t"], res[0]["w"], res[0]["b"]
    logits = (h.double() @ w.double().t() + b.double()).float().numpy()
    want_ids = O.topk_ids(logits, k)
    want_rank = O.target_rank(logits, t.numpy())
    got_ids = np.concatenate([r["out"]["topk_idx"].numpy() for r in res])
    got_rank = np.concatenate([r["out"]["rank"].numpy() for r in res])
    got_ts = np.concatenate([r["out"]["target_score"].numpy() for r in res])
    np.testing.assert_array_equal(got_ids, want_ids)
    np.testing.assert_array_equal(got_rank, want_rank)
    np.testing.assert_array_equal(got_ts, logits[np.arange(len(t)), t.numpy()])
    lite_rank = np.concatenate([r["lite"]["rank"].numpy() for r in res])
    np.testing.assert_array_equal(lite_rank, np.minimum(want_rank, k + 1))
    # validation loss of every rank's own users from the merged softmax statistics == CrossEntropyLoss(ignore_index=0) on the full logits
    lt, tt = torch.from_numpy(logits), t
    for r, part in enumerate(res):
        sl = slice(r * B_local, (r + 1) * B_local)
        want_loss = torch.nn.functional.cross_entropy(lt[sl], tt[sl], ignore_index=0)
        assert abs(float(part["lite"]["loss"]) - float(want_loss)) < 1e-4 * max(1.0, abs(float(want_loss)))
        torch.testing.assert_close(part["lite"]["lse"], torch.logsumexp(lt[sl], dim=1), rtol=1e-5, atol=1e-5)
    # vocab-sharded training cross entropy: loss and lse identical on every rank and equal to the unsharded values; dH all-reduced
    hd = h.clone().double().requires_grad_(True)
    full_loss = torch.nn.functional.cross_entropy(hd @ w.double().t() + b.double(), tt, ignore_index=0)
    full_loss.backward()
    for part in res:
        assert abs(float(part["ce"]["loss"]) - float(full_loss)) < 1e-5 * max(1.0, abs(float(full_loss)))
        torch.testing.assert_close(part["ce"]["lse"], torch.logsumexp(lt, dim=1), rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(part["dh"].double(), hd.grad, rtol=1e-4, atol=1e-6)
    # metric states all-reduced over the ranks == metrics of the whole batch
    want = {name: v / len(want_rank) for name, v in O.metrics_from_rank(want_rank, [1, 5]).items()}     # sums -> means
    for r in res:
        assert abs(r["metrics"]["recall@5"] - want["recall@5"]) < 1e-6
        assert abs(r["metrics"]["recall@1"] - want["recall@1"]) < 1e-6
        assert abs(r["metrics"]["NDCG@5"] - want["NDCG@5"]) < 1e-6
        assert abs(r["metrics"]["MRR@5"] - want["MRR@5"]) < 1e-6


def test_shard_range_covers_catalog():
    from asme_b200.sharded import shard_range
    for V in (1, 7, 1000003, 3709):
        for G in (1, 2, 4, 8):
            edges = [shard_range(V, G, g) for g in range(G)]
            assert edges[0][0] == 0 and edges[-1][1] == V
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
