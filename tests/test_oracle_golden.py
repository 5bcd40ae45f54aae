"""Pins the CPU oracle (oracle/asme_oracle.py) against
 (1) every golden vector of the reference's own metric tests (tests/golden/metric_vectors.json), and
 (2) outputs of the unmodified reference classes (tests/golden/*.npz, made by make_golden.py).
CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import asme_oracle as O

EPSILON = 10e-4          # the reference's own tolerance, tests/util_test_metric.py:11


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    w = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w::")}
    g = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad::")}
    return z, w, g


def _metric_cases(golden_dir=os.path.join(os.path.dirname(__file__), "golden")):
    with open(os.path.join(golden_dir, "metric_vectors.json")) as f:
        vec = json.load(f)
    return [(name, i, s) for name, samples in vec.items() for i, s in enumerate(samples)]


_FN = {"recall": O.recall_at_k, "ndcg": O.ndcg_at_k, "dcg": O.dcg_at_k, "mrr": O.mrr_at_k,
       "precision": O.precision_at_k, "f1": O.f1_at_k}


@pytest.mark.parametrize("name,idx,sample", _metric_cases(), ids=lambda v: str(v) if not isinstance(v, dict) else "")
def test_metric_golden_vectors(name, idx, sample):
    pred = np.array(sample["predictions"], dtype=np.float32)
    mask = np.array(sample["positive_mask"])
    value = _FN[name](pred, mask, sample["k"]).sum() / pred.shape[0]
    assert abs(value - sample["expected"]) < EPSILON


def test_metrics_dense_vs_reference_classes(golden_dir):
    z = np.load(os.path.join(golden_dir, "metrics_dense_small.npz"))
    pred, targets = z["predictions"], z["targets"]
    pm = O.multi_hot(pred.shape, targets)
    rank = O.target_rank(pred, targets)
    from_rank = O.metrics_from_rank(rank, z["ks"])
    for k in z["ks"]:
        k = int(k)
        dense = {f"recall@{k}": O.recall_at_k, f"NDCG@{k}": O.ndcg_at_k, f"MRR@{k}": O.mrr_at_k,
                 f"precision@{k}": O.precision_at_k, f"F1@{k}": O.f1_at_k, f"DCG@{k}": O.dcg_at_k}
        for name, fn in dense.items():
            got = fn(pred, pm, k).sum() / pred.shape[0]
            np.testing.assert_allclose(got, z["final::" + name], rtol=1e-6, atol=1e-7, err_msg=name)
            if name in from_rank:      # single-target shortcut: every @k metric is a function of the rank
                np.testing.assert_allclose(from_rank[name] / pred.shape[0], z["final::" + name], rtol=1e-6,
                                           atol=1e-7, err_msg="rank:" + name)
    np.testing.assert_allclose(O.rank_full(pred, pm).mean(), z["final::rank"], rtol=1e-6)
    np.testing.assert_allclose(rank.mean(), z["final::rank"], rtol=1e-6)
    np.testing.assert_allclose((1.0 / rank).mean(), z["final::MRR"], rtol=1e-6)


def _check_grads(w, g, loss_fn, tied=()):
    leaves = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    for alias, src in tied:
        leaves[alias] = leaves[src]
    loss = loss_fn(leaves)
    loss.backward()
    for name, ref in g.items():
        got = leaves[name].grad
        assert got is not None, name
        torch.testing.assert_close(got, ref, rtol=2e-4, atol=2e-6, msg=lambda m, n=name: f"{n}: {m}")
    return loss.detach()


def test_bert4rec_vs_reference(golden_dir):
    z, w, g = _load(golden_dir, "bert4rec_small.npz")
    inp, tgt = torch.from_numpy(z["input"]), torch.from_numpy(z["target"])
    heads, layers = int(z["heads"]), int(z["L"])
    w.pop("_projection_layer.embedding.weight")            # tied duplicate of the item table
    logits = O.bert4rec_logits(w, inp, heads, layers)
    torch.testing.assert_close(logits, torch.from_numpy(z["logits"]), rtol=1e-5, atol=1e-5)
    g.pop("_projection_layer.embedding.weight", None)
    loss = _check_grads(w, g, lambda lw: O.cross_entropy_ignore_pad(O.bert4rec_logits(lw, inp, heads, layers), tgt))
    torch.testing.assert_close(loss, torch.from_numpy(z["loss"]), rtol=1e-6, atol=1e-6)
    ev = torch.from_numpy(z["eval_input"])
    rows = O.select_masked_rows(O.bert4rec_logits(w, ev, heads, layers), ev)
    torch.testing.assert_close(rows, torch.from_numpy(z["eval_logits"]), rtol=1e-5, atol=1e-5)


def test_kebert4rec_vs_reference(golden_dir):
    z, w, g = _load(golden_dir, "kebert4rec_small.npz")
    inp, tgt = torch.from_numpy(z["input"]), torch.from_numpy(z["target"])
    attrs = {"category": torch.from_numpy(z["category"]), "tags": torch.from_numpy(z["tags"])}
    heads, layers = int(z["heads"]), int(z["L"])
    f = lambda lw: O.kebert4rec_logits(lw, inp, attrs, heads, layers, prefusion=("category", "tags"))
    torch.testing.assert_close(f(w), torch.from_numpy(z["logits"]), rtol=1e-5, atol=1e-5)
    loss = _check_grads(w, g, lambda lw: O.cross_entropy_ignore_pad(f(lw), tgt))
    torch.testing.assert_close(loss, torch.from_numpy(z["loss"]), rtol=1e-6, atol=1e-6)


def test_sasrec_full_vs_reference(golden_dir):
    z, w, g = _load(golden_dir, "sasrec_full_small.npz")
    inp, tgt = torch.from_numpy(z["input"]), torch.from_numpy(z["target"])
    heads, layers = int(z["heads"]), int(z["L"])
    f = lambda lw: O.sasrec_full_logits(lw, inp, {}, heads, layers)
    logits = f(w)
    torch.testing.assert_close(logits, torch.from_numpy(z["logits"]), rtol=1e-5, atol=1e-5)
    loss = _check_grads(w, g, lambda lw: O.cross_entropy_ignore_pad(f(lw), tgt))
    torch.testing.assert_close(loss, torch.from_numpy(z["loss"]), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(O.select_last_rows(logits, inp), torch.from_numpy(z["eval_logits"]), rtol=1e-5,
                               atol=1e-5)


def _user_attrs(z, category="category"):
    return {"user_id": torch.from_numpy(z["user_id"]), "gender": torch.from_numpy(z["gender"]), "category": torch.from_numpy(z[category])}


def test_ubert4rec_vs_reference(golden_dir):
    """user token + segment embedding + item attribute; (B, S+1, V) logits; loss with the pad column the
    reference module prepends to the targets; evaluation rows = MASK position shifted by the user position"""
    z, w, g = _load(golden_dir, "ubert4rec_small.npz")
    inp, tgt = torch.from_numpy(z["input"]), torch.from_numpy(z["target"])
    heads, layers = int(z["heads"]), int(z["L"])
    kw = dict(additional=("category",), user=("user_id", "gender"))
    f = lambda lw, seq=inp, a=_user_attrs(z): O.ubert4rec_logits(lw, seq, a, heads, layers, **kw)
    torch.testing.assert_close(f(w), torch.from_numpy(z["logits"]), rtol=1e-5, atol=1e-5)
    tgt1 = torch.cat([torch.zeros(tgt.shape[0], 1, dtype=tgt.dtype), tgt], dim=1)
    g.pop("_sequence_embedding_layer.segment_embedding.weight")     # row 2 of the table is never read: autograd leaves zeros there
    loss = _check_grads(w, g, lambda lw: O.cross_entropy_ignore_pad(f(lw), tgt1))
    torch.testing.assert_close(loss, torch.from_numpy(z["loss"]), rtol=1e-6, atol=1e-6)
    ev = torch.from_numpy(z["eval_input"])
    full = f(w, ev, _user_attrs(z, "eval_category"))
    rows = full[torch.cat([torch.zeros(ev.shape[0], 1, dtype=torch.bool), ev.eq(1)], dim=1)]
    torch.testing.assert_close(rows, torch.from_numpy(z["eval_logits"]), rtol=1e-5, atol=1e-5)


def test_usasrec_full_vs_reference(golden_dir):
    z, w, g = _load(golden_dir, "usasrec_full_small.npz")
    inp, tgt = torch.from_numpy(z["input"]), torch.from_numpy(z["target"])
    heads, layers = int(z["heads"]), int(z["L"])
    f = lambda lw: O.usasrec_full_logits(lw, inp, _user_attrs(z), heads, layers, additional=("category",), user=("user_id", "gender"))
    logits = f(w)
    torch.testing.assert_close(logits, torch.from_numpy(z["logits"]), rtol=1e-5, atol=2e-5)
    loss = _check_grads(w, g, lambda lw: O.cross_entropy_ignore_pad(f(lw)[:, 1:, :], tgt))
    torch.testing.assert_close(loss, torch.from_numpy(z["loss"]), rtol=1e-6, atol=1e-6)
    # the reference module's evaluation rows: index (length - 1) of the S+1 positions
    rows = logits[torch.arange(inp.shape[0]), inp.ne(0).sum(-1) - 1]
    torch.testing.assert_close(rows, torch.from_numpy(z["eval_logits"]), rtol=1e-5, atol=2e-5)


def test_sasrec_neg_vs_reference(golden_dir):
    z, w, g = _load(golden_dir, "sasrec_neg_small.npz")
    inp = torch.from_numpy(z["input"])
    pos, neg = torch.from_numpy(z["positive_samples"]), torch.from_numpy(z["negative_samples"])
    heads, layers = int(z["heads"]), int(z["L"])
    p, n = O.sasrec_neg_logits(w, inp, pos, neg, heads, layers)
    torch.testing.assert_close(p, torch.from_numpy(z["pos_logits"]), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(n, torch.from_numpy(z["neg_logits"]), rtol=1e-5, atol=1e-5)

    def loss_fn(lw):
        pp, nn_ = O.sasrec_neg_logits(lw, inp, pos, neg, heads, layers)
        return O.sasrec_bce(pp, nn_, inp.ne(0))

    loss = _check_grads(w, g, loss_fn)
    torch.testing.assert_close(loss, torch.from_numpy(z["loss"]), rtol=1e-6, atol=1e-6)
    h = O.sasrec_hidden(w, inp, {}, heads, layers)
    items = torch.arange(int(z["V"])).repeat(inp.shape[0], 1)
    ev = O.sasrec_score_items(h, inp.ne(0), w["_sequence_embedding_layer.item_embedding_layer.item_embedding.embedding.weight"], items)
    torch.testing.assert_close(ev, torch.from_numpy(z["eval_logits"]), rtol=1e-5, atol=1e-5)


def test_attention_fully_masked_rows_are_uniform():
    """Quirk Q4: -1e9 fill => a fully masked query row attends uniformly to ALL S keys."""
    torch.manual_seed(0)
    b, s, h, heads = 2, 5, 8, 2
    x = torch.randn(b, s, h)
    w = {}
    for i in range(3):
        w[f"a.linear_layers.{i}.weight"] = torch.randn(h, h)
        w[f"a.linear_layers.{i}.bias"] = torch.randn(h)
    w["a.output_linear.weight"] = torch.eye(h)
    w["a.output_linear.bias"] = torch.zeros(h)
    pm = torch.zeros(b, s, dtype=torch.bool)            # everything padded
    out = O.multi_head_attention(x, w, "a", heads, O.attention_mask(pm, b, s, True))
    v = x @ w["a.linear_layers.2.weight"].t() + w["a.linear_layers.2.bias"]
    torch.testing.assert_close(out, v.mean(dim=1, keepdim=True).expand(-1, s, -1), rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------------------------------
# input pipeline: the oracle's per-sample restatements draw from torch's global generator in the
# reference's order, so the reference's own seeded golden vectors pin them exactly
# ------------------------------------------------------------------------------------------------------------
def test_cloze_mask_golden_vectors_of_the_reference():
    """/root/reference/tests/test_cloze_mask.py:9-37 (seed_everything(42), example vocabulary of 13 tokens)"""
    torch.manual_seed(42)
    seq, tgt = O.cloze_mask_sequence([5, 8, 9, 7, 3, 4], 1.0, 1.0, 13)
    assert seq == [5, 8, 9, 7, 3, 1] and tgt == [0, 0, 0, 0, 0, 4]
    torch.manual_seed(42)
    seq, tgt = O.cloze_mask_sequence([5, 8, 9, 7, 3, 4, 12, 10, 11, 3], 0.5, 0.1, 13)
    assert seq == [5, 1, 9, 1, 3, 1, 12, 10, 1, 3] and tgt == [0, 8, 0, 7, 0, 4, 0, 0, 11, 0]


def test_pos_neg_sampler_golden_vector_of_the_reference():
    """/root/reference/tests/test_pos_neg.py:9-24"""
    torch.manual_seed(42)
    x, pos, neg = O.pos_neg_sequence([5, 8, 9, 7, 3, 4], 13)
    assert x == [5, 8, 9, 7, 3] and pos == [8, 9, 7, 3, 4] and neg == [6, 6, 6, 6, 11]
