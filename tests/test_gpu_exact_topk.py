"""Exact top-k on the tensor-core path (``-m gpu``).  Reference semantics: metrics/common.py:18-27 -- argsort of the fp32 logits
(ties -> lowest item id by this build's contract).

The bf16 sweep proposes KC candidates, ``ops.topk_rescore`` re-scores them from the fp32 operands with the fp32 path's arithmetic,
orders and certifies them, ``ops.score_topk_flagged`` re-runs the exact sweep for rows without a certificate.  Bars:
  * lists / scores / target positions EQUAL to the strict fp32 sweep (``ops.score_topk_rank``) -- bit for bit, ties included;
  * EQUAL to ``O.topk_ids`` / ``O.target_rank`` of logits computed from the fp32 operands (float64 accumulation) on the C2, C3 and
    C5 shapes, Recall / NDCG / MRR identical;
  * adversarial catalogs (thousands of items inside the bf16 error band) are caught by the certificate and still come out exact."""
import numpy as np
import pytest
import torch

from oracle import asme_oracle as O

pytestmark = pytest.mark.gpu


def _exact(ops, models, h, w, b, k, target=None, full_rank=False):
    wb = ops.cast_bf16_ext(w, b) if b is not None else ops.cast_bf16(w)
    hb = ops.cast_bf16_ext(h) if b is not None else ops.cast_bf16(h, ld_out=wb.shape[1])
    return models.score_rows_tc_exact(h, hb, wb, w, b, ops.table_norm_bound(w, b), target, k, full_rank)


def _fp32_sweep(ops, h, w, b, k, target):
    ts = ops.score_targets(h, w, b, target)
    val, idx, ng, nt = ops.score_topk_rank(h, w, b, k, target, ts)
    return val, idx, (ng + nt + 1), ts


@pytest.mark.parametrize("R,V,H,k,bias", [(64, 3709, 64, 10, True), (300, 12104, 64, 10, True), (33, 500, 128, 5, False), (128, 100003, 128, 10, True),
                                          (17, 40, 64, 10, True), (5, 7, 64, 10, False), (200, 30000, 256, 1, True), (100, 20000, 64, 20, True)])
def test_exact_topk_equals_the_fp32_sweep(R, V, H, k, bias):
    from asme_b200 import models, ops
    g = torch.Generator(device="cuda").manual_seed(R + V + k)
    h = torch.randn(R, H, generator=g, device="cuda")
    w = torch.randn(V, H, generator=g, device="cuda") * 0.1
    b = torch.randn(V, generator=g, device="cuda") * 0.1 if bias else None
    target = torch.randint(0, V, (R,), generator=g, device="cuda")
    out = _exact(ops, models, h, w, b, k, target)
    val, idx, rank, ts = _fp32_sweep(ops, h, w, b, k, target)
    assert torch.equal(out["topk_idx"], idx)
    assert torch.equal(out["topk_val"], val)
    assert torch.equal(out["target_score"], ts)
    # rows the certificate could not cover were re-run by the exact sweep and carry their exact FULL rank
    assert torch.equal(torch.clamp(out["rank"], max=k + 1), torch.clamp(rank, max=k + 1).to(torch.int32))
    if V >= 10000 and k <= 10:
        assert int(out["n_uncertified"]) == 0, "well separated random catalogs are expected to certify every row"


def test_exact_topk_with_planted_targets_and_full_rank():
    from asme_b200 import models, ops
    R, V, H, k = 256, 50021, 128, 10
    g = torch.Generator(device="cuda").manual_seed(5)
    h = torch.randn(R, H, generator=g, device="cuda")
    w = torch.randn(V, H, generator=g, device="cuda") * 0.05
    b = torch.randn(V, generator=g, device="cuda") * 0.05
    free = torch.randint(0, V, (R,), generator=g, device="cuda")
    first = _exact(ops, models, h, w, b, 20, free)
    col = torch.randint(0, 20, (R,), generator=g, device="cuda")
    planted = first["topk_idx"].gather(1, col.unsqueeze(1)).squeeze(1).to(torch.int64)
    target = torch.where(torch.arange(R, device="cuda") % 2 == 0, planted, free)
    out = _exact(ops, models, h, w, b, k, target, full_rank=True)
    val, idx, rank, ts = _fp32_sweep(ops, h, w, b, k, target)
    assert torch.equal(out["topk_idx"], idx) and torch.equal(out["topk_val"], val)
    inside = rank <= k
    assert int(inside.sum()) > R // 8                       # planted targets really land inside the lists
    assert torch.equal(out["rank"][inside], rank[inside].to(torch.int32))
    # outside the top k the rank comes from the bf16 count sweep: never inside, and close to the exact one
    assert bool((out["rank"][~inside] > k).all())
    diff = (out["rank"][~inside].double() - rank[~inside].double()).abs()
    assert bool((diff <= torch.clamp(0.05 * rank[~inside].double(), min=3.0)).all())


def test_adversarial_ties_are_caught_by_the_certificate():
    """a catalog in which thousands of items differ only BELOW bf16 resolution: the bf16 sweep cannot tell them apart, its lists are
    arbitrary among them -- the certificate must refuse and the exact sweep must repair the rows"""
    from asme_b200 import models, ops
    R, V, H, k = 40, 6000, 64, 10
    g = torch.Generator(device="cuda").manual_seed(11)
    h = torch.randn(R, H, generator=g, device="cuda")
    base = torch.randn(1, H, generator=g, device="cuda")
    w = base.repeat(V, 1) * (1.0 + 1e-5 * torch.randn(V, 1, generator=g, device="cuda")) + 1e-6 * torch.randn(V, H, generator=g, device="cuda")
    w[: V // 2] = torch.randn(V // 2, H, generator=g, device="cuda") * 0.01          # half the catalog is far away
    b = None
    target = torch.randint(0, V, (R,), generator=g, device="cuda")
    out = _exact(ops, models, h, w, b, k, target)
    val, idx, rank, ts = _fp32_sweep(ops, h, w, b, k, target)
    assert int(out["n_uncertified"]) > 0
    assert torch.equal(out["topk_idx"], idx) and torch.equal(out["topk_val"], val)
    assert torch.equal(torch.clamp(out["rank"], max=k + 1), torch.clamp(rank, max=k + 1).to(torch.int32))


def test_exact_ties_resolve_to_the_lowest_item_id():
    from asme_b200 import models, ops
    R, V, H, k = 16, 300, 64, 10
    g = torch.Generator(device="cuda").manual_seed(2)
    h = torch.randint(-2, 3, (R, H), generator=g, device="cuda").float()
    w = torch.randint(-2, 3, (V, H), generator=g, device="cuda").float()           # small integers: scores are exact and tie a lot
    target = torch.randint(0, V, (R,), generator=g, device="cuda")
    out = _exact(ops, models, h, w, None, k, target)
    logits = (h.double() @ w.double().t()).cpu().numpy()
    assert np.array_equal(out["topk_idx"].cpu().numpy(), O.topk_ids(logits, k))
    want_rank = O.target_rank(logits, target.cpu().numpy())
    assert np.array_equal(np.minimum(out["rank"].cpu().numpy(), k + 1), np.minimum(want_rank, k + 1))


def _model_rows(model, seq):
    """the hidden rows the evaluation scores (after the modifier), and the fp32 projection operands"""
    from asme_b200.models import mask_position_rows
    rows = mask_position_rows(seq, 1)
    h = model.encode_rows(seq, seq.ne(0), {}, rows, one_per_sequence=True)
    m_rows, _ = model.modify(h)
    return m_rows


def _oracle_check(model, seq, target, k=10, chunk=64):
    m_rows = _model_rows(model, seq)
    w, b = model.projection_operands()
    out = model.evaluate_rank(seq, seq.ne(0), {}, target, k=k, full_rank=False)
    assert int(out["n_uncertified"]) == 0
    B = seq.shape[0]
    for r0 in range(0, B, chunk):          # fp32 operands, float64 accumulation: the reference's logits without its rounding noise
        logits = (m_rows[r0:r0 + chunk].double() @ w.double().t() + b.double()).cpu().numpy()
        assert np.array_equal(out["topk_idx"][r0:r0 + chunk].cpu().numpy(), O.topk_ids(logits, k)), f"rows {r0}.."
        want = np.minimum(O.target_rank(logits, target[r0:r0 + chunk].cpu().numpy()), k + 1)
        assert np.array_equal(out["rank"][r0:r0 + chunk].cpu().numpy(), want), f"ranks of rows {r0}.."
        m = O.metrics_from_rank(want, [1, 5, 10])
        got = O.metrics_from_rank(out["rank"][r0:r0 + chunk].cpu().numpy(), [1, 5, 10])
        assert m == got
    return out


def _eval_batch(B, S, V, seed):
    g = torch.Generator().manual_seed(seed)
    seq = torch.randint(3, V, (B, S), generator=g)
    lengths = torch.randint(2, S, (B,), generator=g)
    seq = torch.where(torch.arange(S).unsqueeze(0) < lengths.unsqueeze(1), seq, torch.zeros_like(seq))
    seq[torch.arange(B), lengths] = 1
    return seq.cuda(), torch.randint(3, V, (B,), generator=g).cuda()


@pytest.mark.parametrize("name,V,S,H,B", [("C2", 3709, 200, 64, 256), ("C3", 12104, 50, 64, 256), ("C5", 1_000_003, 200, 128, 256)])
def test_model_evaluation_is_exact_on_the_benchmark_shapes(name, V, S, H, B):
    """bf16 policy, default settings: top-k ids, target positions and the @k metrics equal those of the fp32 logits of the same
    hidden rows -- half of the targets planted inside each user's own top 20 so that the positions are exercised"""
    from asme_b200.models import BERT4RecModel
    torch.manual_seed(0)
    model = BERT4RecModel(H, 2, 2, V, S, 0.0).cuda().eval()
    assert model.precision == "bf16" and model.exact_topk
    seq, free = _eval_batch(B, S, V, 7)
    first = model.evaluate_rank(seq, seq.ne(0), {}, free, k=20)
    g = torch.Generator(device="cuda").manual_seed(1)
    col = torch.randint(0, 20, (B,), generator=g, device="cuda")
    planted = first["topk_idx"].gather(1, col.unsqueeze(1)).squeeze(1).to(torch.int64)
    target = torch.where(torch.arange(B, device="cuda") % 2 == 0, planted, free)
    out = _oracle_check(model, seq, target)
    assert int((out["rank"] <= 10).sum()) > B // 8


@pytest.mark.parametrize("R,V,H,k,sigma_b", [(64, 3709, 64, 10, 0.1), (128, 100003, 128, 10, 0.001), (128, 100003, 128, 10, 0.3), (33, 20011, 128, 5, 0.05),
                                            (200, 30000, 64, 1, 0.1)])
def test_exact_topk_with_the_bias_bounded_per_chunk(R, V, H, k, sigma_b):
    """layers.py:138-143 (h E^T + output_bias): the sweep on the plain (V, H) table with the bias only BOUNDED per 32-item chunk
    (ops.bias_chunk_bounds; exact add where a chunk could pass the threshold) returns the lists of the strict fp32 sweep, for a bias
    far below and one as large as the spread of the scores (every chunk then takes the exact path)."""
    from asme_b200 import models, ops
    g = torch.Generator(device="cuda").manual_seed(R + V + k + 1)
    h = torch.randn(R, H, generator=g, device="cuda")
    w = torch.randn(V, H, generator=g, device="cuda") * 0.02
    b = torch.randn(V, generator=g, device="cuda") * sigma_b
    target = torch.randint(0, V, (R,), generator=g, device="cuda")
    bb = ops.bias_chunk_bounds(b)
    bref = b.view(-1)[: V // 32 * 32].view(-1, 32)
    assert torch.equal(bb[: V // 32, 0], bref.max(1).values) and torch.equal(bb[: V // 32, 1], bref.min(1).values)
    wb = ops.cast_bf16(w)
    hb = ops.cast_bf16(h, ld_out=wb.shape[1])
    out = models.score_rows_tc_exact(h, hb, wb, w, b, ops.table_norm_bound(w, b), target, k, False, bias_bounds=(b, bb))
    val, idx, rank, ts = _fp32_sweep(ops, h, w, b, k, target)
    assert torch.equal(out["topk_idx"], idx)
    assert torch.equal(out["topk_val"], val)
    assert torch.equal(out["target_score"], ts)
    assert torch.equal(torch.clamp(out["rank"], max=k + 1), torch.clamp(rank, max=k + 1).to(torch.int32))


@pytest.mark.parametrize("R,V,k", [(2500, 70001, 10), (8192, 125001, 10), (4096, 250001, 10), (1100, 40000, 3)])
def test_exact_topk_with_many_row_tiles(R, V, k):
    """a rank of the vocab-sharded evaluation scores ALL users of the box against its slice: so many row tiles that the catalog is
    not cut into 16 CTAs per row tile -- every CTA sweeps several parts in turn and its warpgroups write their lists unfolded
    (ScoreTcArgs::seq_parts).  Lists / scores equal to the strict fp32 sweep on a row sample, every row certified, and the same
    lists from the first scheme (knob 8 = 0)."""
    from asme_b200 import models, ops
    H = 128
    g = torch.Generator(device="cuda").manual_seed(R + V + k)
    h = torch.randn(R, H, generator=g, device="cuda")
    w = torch.randn(V, H, generator=g, device="cuda") * 0.03
    b = torch.randn(V, generator=g, device="cuda") * 0.01
    target = torch.randint(0, V, (R,), generator=g, device="cuda")
    out = _exact(ops, models, h, w, b, k, target)
    assert int(out["n_uncertified"]) == 0
    rows = torch.arange(0, R, max(1, R // 300), device="cuda")
    val, idx, rank, ts = _fp32_sweep(ops, h[rows].contiguous(), w, b, k, target[rows].contiguous())
    assert torch.equal(out["topk_idx"][rows], idx)
    assert torch.equal(out["topk_val"][rows], val)
    assert torch.equal(out["target_score"][rows], ts)
    ops._lib.call("asme_b200_tc_score_tune", 8, 0)
    try:
        first = _exact(ops, models, h, w, b, k, target)
    finally:
        ops._lib.call("asme_b200_tc_score_tune", 8, 1)
    assert torch.equal(first["topk_idx"], out["topk_idx"]) and torch.equal(first["topk_val"], out["topk_val"])
