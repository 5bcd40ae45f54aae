"""Model-level parity of the tensor-core policy (``precision = "bf16"``: tcgen05 GEMMs, attention and catalog scoring
with bf16 operands / fp32 accumulation; residual stream, LayerNorm, softmax statistics, loss and optimizer state in
fp32) against the reference fixtures and the fp32 CPU oracle.

Tolerances (north_star: "logits and losses within 1e-3 relative in bf16"):
  * loss                      1e-3 relative
  * gradients                 2e-2 of the tensor's norm (bf16 operand rounding is amplified by the backward chain)
  * evaluation                top-k lists and ranks are compared through the scores: every returned item's oracle score is
                              within 2e-3 of the logit scale of the oracle's score at the same list position, and ranks agree
                              wherever the oracle's neighbouring scores are further apart than that
"""
import numpy as np
import pytest
import torch

from oracle import asme_oracle as O
from test_gpu_models import _cpu_weights, _grads, _random_batch
from test_host_cpu import build_from_fixture

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-3
GRAD_NORM_TOL = 2e-2
SCORE_TOL = 2e-3


@pytest.fixture(autouse=True)
def _bf16_policy():
    from asme_b200 import models
    old = models.DEFAULT_PRECISION
    models.set_default_precision("bf16")
    yield
    models.set_default_precision(old)


def rel(a, b):
    return abs(float(a) - float(b)) / max(abs(float(b)), 1e-12)


def check_grads(got, want, skip=()):
    """norm-wise comparison; tensors whose true gradient is numerically zero (e.g. the key bias: softmax is invariant to a
    constant added to every key) are compared against an absolute floor tied to the largest gradient entry of the model"""
    want = {k: torch.as_tensor(v).float() for k, v in want.items() if v is not None}
    gmax = max(float(v.abs().max()) for v in want.values())
    for name, g_ref in want.items():
        if name in skip or name not in got or name.endswith("attention.linear_layers.1.bias"):
            continue          # key bias: its exact gradient is 0 (softmax invariance), both sides hold rounding noise only
        floor = 1e-3 * gmax * g_ref.numel() ** 0.5
        err = float((got[name].float() - g_ref).norm()) / (float(g_ref.norm()) + floor)
        assert err < GRAD_NORM_TOL, f"grad {name}: relative norm error {err:.4f}"


def check_eval(out, rows, target, k):
    """rows: oracle fp32 logits (B,V) of the selected positions"""
    rows = np.asarray(rows, dtype=np.float64)
    scale = np.abs(rows).max()
    idx = out["topk_idx"].cpu().numpy().astype(np.int64)
    srt = -np.sort(-rows, axis=1)
    got_scores = np.take_along_axis(rows, idx, axis=1)
    assert np.abs(got_scores - srt[:, :k]).max() <= SCORE_TOL * scale
    want_rank = O.target_rank(rows.astype(np.float32), target)
    got_rank = out["rank"].cpu().numpy()
    st = rows[np.arange(rows.shape[0]), target][:, None]
    near = (np.abs(rows - st) <= SCORE_TOL * scale).sum(axis=1) - 1          # competitors within bf16 noise of the target
    assert (np.abs(got_rank - want_rank) <= near).all(), (got_rank, want_rank, near)


def test_bert4rec_fixture_bf16(golden_dir):
    z, w, model = build_from_fixture(golden_dir, "bert4rec_small.npz")
    model.load_state_dict(w)
    model = model.cuda().train()
    assert model.precision == "bf16"
    inp, tgt = torch.from_numpy(z["input"]).cuda(), torch.from_numpy(z["target"]).cuda()
    loss, ctx = model.loss_ce(inp, inp.ne(0), {}, tgt)
    assert rel(loss, z["loss"]) < LOSS_RTOL
    model.loss_ce_backward(ctx)
    check_grads(_grads(model), {k[6:]: z[k] for k in z.files if k.startswith("grad::")})
    model.eval()
    ev, et = torch.from_numpy(z["eval_input"]).cuda(), torch.from_numpy(z["eval_target"]).cuda()
    out = model.evaluate_rank(ev, ev.ne(0), {}, et, k=10, with_loss=True)
    check_eval(out, z["eval_logits"], z["eval_target"], 10)
    ce = torch.nn.functional.cross_entropy(torch.from_numpy(z["eval_logits"]), torch.from_numpy(z["eval_target"]), ignore_index=0)
    assert rel(out["loss"], ce) < 2 * LOSS_RTOL
    # rank from the top-k list only (no count sweep): identical wherever the target is inside the list
    lite = model.evaluate_rank(ev, ev.ne(0), {}, et, k=10, full_rank=False)
    full, part = out["rank"].cpu().numpy(), lite["rank"].cpu().numpy()
    assert np.array_equal(np.minimum(full, 11), part)


def test_bert4rec_c2_shape_bf16_vs_oracle():
    """C2 shape (V=3709, S=200, H=64, L=2, heads=2), batch 16: tensor-core encoder + fused CE vs the fp32 oracle"""
    from asme_b200.models import BERT4RecModel
    torch.manual_seed(0)
    V, S, H, B = 3709, 200, 64, 16
    model = BERT4RecModel(H, 2, 2, V, S, 0.0, initializer_range=0.1).cuda().train()
    assert model.engine.use_tc()
    w = _cpu_weights(model)
    seq, target, _ = _random_batch(torch.Generator().manual_seed(1235), B, S, V)
    loss, ctx = model.loss_ce(seq.cuda(), seq.cuda().ne(0), {}, target.cuda())
    model.loss_ce_backward(ctx)
    leaves = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    ref = O.cross_entropy_ignore_pad(O.bert4rec_logits(leaves, seq, 2, 2), target)
    ref.backward()
    assert rel(loss, ref.detach()) < LOSS_RTOL
    check_grads(_grads(model), {k: v.grad for k, v in leaves.items()})


def test_kebert4rec_c3_shape_bf16_vs_oracle():
    from asme_b200.models import KeBERT4RecModel
    torch.manual_seed(0)
    V, S, H, B = 12104, 50, 64, 32
    pre = {"category": {"embedding_type": "content_embedding"}, "tags": {"embedding_type": "linear_upscale"}}
    model = KeBERT4RecModel(H, 2, 2, V, S, 0.0, prefusion_attributes=pre, attribute_vocab_sizes={"category": 256, "tags": 512},
                            initializer_range=0.1).cuda().train()
    w = _cpu_weights(model)
    gen = torch.Generator().manual_seed(1236)
    seq, target, _ = _random_batch(gen, B, S, V)
    cat = torch.randint(3, 256, (B, S), generator=gen)
    tags = torch.randint(0, 512, (B, S, 4), generator=gen)
    cat[seq == 0] = 0
    tags[seq == 0] = 0
    attrs = {"category": cat, "tags": tags}
    loss, ctx = model.loss_ce(seq.cuda(), seq.cuda().ne(0), {k: v.cuda() for k, v in attrs.items()}, target.cuda())
    model.loss_ce_backward(ctx)
    leaves = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    ref = O.cross_entropy_ignore_pad(O.kebert4rec_logits(leaves, seq, attrs, 2, 2, prefusion=("category", "tags")), target)
    ref.backward()
    assert rel(loss, ref.detach()) < LOSS_RTOL
    check_grads(_grads(model), {k: v.grad for k, v in leaves.items()})


def test_ubert4rec_bf16_vs_oracle():
    """user-attribute model on the tensor-core path: 51 positions (user token + 50 items), causal encoder, H=64, through the
    drop-in training module (targets get the pad column) and the fused evaluation with the module's MASK rows"""
    from asme_b200.models import UBERT4RecModel
    from asme_b200.modules import UBERTMaskedTrainingModule
    torch.manual_seed(0)
    V, S, H, B = 5003, 50, 64, 48
    kw = dict(additional_attributes={"category": {"embedding_type": "content_embedding"}},
              user_attributes={"user_id": {"embedding_type": "user_embedding"}, "gender": {"embedding_type": "content_embedding"}},
              attribute_vocab_sizes={"category": 40, "user_id": 900, "gender": 6})
    model = UBERT4RecModel(H, 2, 2, V, S, 0.0, segment_embedding=True, initializer_range=0.05, **kw).cuda().train()
    assert model.engine.use_tc()
    w = _cpu_weights(model)
    gen = torch.Generator().manual_seed(4321)
    seq, target, _ = _random_batch(gen, B, S, V)
    cat = torch.randint(3, 40, (B, S), generator=gen)
    cat[seq == 0] = 0
    uid = torch.randint(3, 900, (B, 1), generator=gen).repeat(1, S)
    gender = torch.randint(3, 6, (B, 1), generator=gen).repeat(1, S)
    attrs = {"user_id": uid, "gender": gender, "category": cat}
    module = UBERTMaskedTrainingModule(model, num_warmup_steps=0)
    batch = {"item": seq.cuda(), "item.target": target.cuda(), **{k: v.cuda() for k, v in attrs.items()}}
    out = module.training_step(batch, 0)
    out["loss"].backward()
    leaves = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    logits = O.ubert4rec_logits(leaves, seq, attrs, 2, 2, additional=("category",), user=("user_id", "gender"))
    tgt1 = torch.cat([torch.zeros(B, 1, dtype=target.dtype), target], dim=1)
    ref = O.cross_entropy_ignore_pad(logits, tgt1)
    ref.backward()
    assert rel(out["loss"], ref.detach()) < LOSS_RTOL
    check_grads(_grads(model), {k: v.grad for k, v in leaves.items()})
    # evaluation: one MASK per row, scored at position (1 + MASK position) of the 51
    model.eval()
    ev = seq.clone()
    lengths = ev.ne(0).sum(-1)
    for i in range(B):
        n = min(int(lengths[i]), S - 1)
        ev[i, n] = 1
        ev[i, n + 1:] = 0
    ev[ev.eq(1).cumsum(-1) > 1] = 3          # cloze masks left in the sequence: keep exactly one MASK (the appended one is last)
    first_mask = ev.eq(1).int().argmax(-1)
    et = torch.randint(3, V, (B,), generator=gen)
    c2 = cat.clone()
    c2[ev == 0] = 0
    eattrs = {"user_id": uid, "gender": gender, "category": c2}
    res = model.evaluate_rank(ev.cuda(), ev.cuda().ne(0), {k: v.cuda() for k, v in eattrs.items()}, et.cuda(), k=10,
                              rows=module._mask_rows(ev.cuda()))
    with torch.no_grad():
        full = O.ubert4rec_logits(w, ev, eattrs, 2, 2, additional=("category",), user=("user_id", "gender"))
    rows = full[torch.arange(B), first_mask + 1]
    check_eval(res, rows.numpy(), et.numpy(), 10)


def test_sasrec_neg_c4_shape_bf16_vs_oracle():
    from asme_b200.models import SASRecModel
    torch.manual_seed(0)
    V, S, H, B = 13047, 50, 64, 64
    model = SASRecModel(H, 2, 2, V, S, 0.0, mode="neg_sampling").cuda().train()
    w = _cpu_weights(model, drop=("_projection_layer",))
    gen = torch.Generator().manual_seed(1237)
    seq, _, lengths = _random_batch(gen, B, S, V, p_mask=0.0)
    seq[seq == 1] = 7
    pos, neg = torch.randint(3, V, (B, S), generator=gen), torch.randint(3, V, (B, S), generator=gen)
    pos[seq == 0] = 0
    neg[seq == 0] = 0
    mask = seq.ne(0)
    loss, ctx = model.loss_bce(seq.cuda(), mask.cuda(), {}, pos.cuda(), neg.cuda(), mask.cuda())
    model.loss_bce_backward(ctx)
    leaves = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    p, n = O.sasrec_neg_logits(leaves, seq, pos, neg, 2, 2)
    ref = O.sasrec_bce(p, n, mask)
    ref.backward()
    assert rel(loss, ref.detach()) < LOSS_RTOL
    check_grads(_grads(model), {k: v.grad for k, v in leaves.items()})


def test_training_with_dropout_runs_and_learns_bf16():
    """statistical check of the full tensor-core training step with dropout 0.2 (masks cannot match ATen's): the loss of
    a fixed batch decreases under the fused Adam"""
    from asme_b200.metrics import build_metrics
    from asme_b200.models import BERT4RecModel
    from asme_b200.modules import MaskedTrainingModule
    torch.manual_seed(0)
    V, S, H, B = 503, 40, 64, 64
    model = BERT4RecModel(H, 2, 2, V, S, 0.2)
    module = MaskedTrainingModule(model, metrics=build_metrics({"recall": [10]}), learning_rate=1e-2, num_warmup_steps=1).cuda()
    module.train()
    (opt,), (sched,) = module.configure_optimizers()
    seq, target, _ = _random_batch(torch.Generator().manual_seed(3), B, S, V)
    batch = {"item": seq.cuda(), "item.target": target.cuda()}
    losses = []
    for i in range(30):
        opt.zero_grad()
        out = module.training_step(batch, i)
        out["loss"].backward()
        opt.step()
        sched["scheduler"].step()
        losses.append(float(out["loss"].detach()))
    assert all(np.isfinite(losses))
    assert np.mean(losses[-5:]) < 0.93 * np.mean(losses[:5]), losses


def test_encode_rows_equals_full_encode_on_selected_rows():
    """evaluation encodes only the selected positions in the last layer: bit-identical to gathering from the full encode"""
    from asme_b200 import ops
    from asme_b200.models import BERT4RecModel, SASRecModel, mask_position_rows, last_position_rows
    torch.manual_seed(0)
    for cls, kw, S in ((BERT4RecModel, {}, 70), (SASRecModel, {"mode": "full"}, 70), (BERT4RecModel, {}, 200), (SASRecModel, {"mode": "full"}, 131)):
        V, H, B = 997, 128, 37
        model = cls(H, 2, 2, V, S, 0.1, **kw).cuda().eval()
        seq, _, lengths = _random_batch(torch.Generator().manual_seed(2), B, S, V, p_mask=0.0)
        seq = seq.cuda()
        seq[torch.arange(B), (lengths - 1).cuda()] = 1
        rows = mask_position_rows(seq, 1) if cls is BERT4RecModel else last_position_rows(seq, seq.ne(0))
        full, _ = model.encode(seq, seq.ne(0), {}, training=False)
        sel = model.encode_rows(seq, seq.ne(0), {}, rows)
        assert torch.equal(sel, ops.gather_rows(full, rows))
        # one position per sequence: the last layer's attention is the fp32 matrix-vector kernel; the tensor-core tile rounds the
        # probabilities to bf16 (2^-9 relative) before P V, so the rows agree to that rounding, not bit for bit
        sel1 = model.encode_rows(seq, seq.ne(0), {}, rows, one_per_sequence=True)
        assert float((sel1 - sel).norm() / sel.norm()) < 3e-3
        torch.testing.assert_close(sel1, sel, rtol=2e-2, atol=2e-2 * float(sel.abs().max()))


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_predict_topn_matches_softmax_sort_of_the_dense_logits(precision):
    """`predict` on the fused path (module.predict_topn -> asme_b200.evaluation evaluators) against the reference arithmetic
    (softmax + sort of predict_step's dense logits, evaluation/evaluation.py:176-178): same items in the same order wherever
    the dense scores are separated by more than the precision policy's noise, softmax scores within the policy's tolerance"""
    from asme_b200.models import BERT4RecModel, SASRecModel
    from asme_b200.modules import MaskedTrainingModule, NextItemPredictionTrainingModule
    from asme_b200.evaluation import top_predictions
    torch.manual_seed(1)
    V, S, H, B, n = 1203, 24, 64, 33, 20

    class Tok:
        pad_token_id, mask_token_id = 0, 1

    for cls, mod_cls, kw in ((BERT4RecModel, MaskedTrainingModule, {"initializer_range": 0.2}),
                             (SASRecModel, NextItemPredictionTrainingModule, {"mode": "full"})):
        model = cls(H, 2, 2, V, S, 0.1, **kw)
        model.precision = precision
        module = mod_cls(model, item_tokenizer=Tok()).cuda().eval()
        seq, _, lengths = _random_batch(torch.Generator().manual_seed(7), B, S, V, p_mask=0.0)
        if cls is BERT4RecModel:
            seq[torch.arange(B), lengths - 1] = 1
        batch = {"item": seq.cuda()}
        fused = module.predict_topn(batch, n)
        dense = module.predict_step(batch, 0).float()
        assert dense.shape == (B, V) and fused.topk_idx.shape == (B, n)
        got_s, got_i = top_predictions(fused, n)
        want_s, want_i = top_predictions(dense, n)
        tol = 2e-2 if precision == "bf16" else 1e-4
        np.testing.assert_allclose(got_s, want_s, rtol=tol, atol=tol * float(want_s.max()))
        # every listed item carries (up to the tolerance) the dense score of that item; the order may only differ inside near-ties
        probs = torch.softmax(dense, dim=-1).cpu().numpy()
        np.testing.assert_allclose(got_s, np.take_along_axis(probs, got_i, axis=1), rtol=tol, atol=tol * float(want_s.max()))
        if precision == "fp32":
            assert (got_i == want_i).mean() > 0.99
        assert (np.diff(got_s, axis=1) <= 1e-7).all()                      # best first
    with pytest.raises(ValueError, match="1..32"):
        module.predict_topn(batch, 40)


def test_per_sample_metrics_evaluator_fused_equals_dense():
    """PerSampleMetricsEvaluator: module.predict_topn(batch, n, with_rank=True) (exact target rank from the counting sweep) gives the
    per-sample metric values of the reference arithmetic on predict_step's dense logits (fp32 policy: identical ranks)"""
    from asme_b200.models import BERT4RecModel
    from asme_b200.modules import MaskedTrainingModule
    from asme_b200.metrics import build_metrics
    from asme_b200.evaluation import PerSampleMetricsEvaluator, top_predictions
    torch.manual_seed(3)
    V, S, H, B, n = 903, 20, 64, 40, 10

    class Tok:
        pad_token_id, mask_token_id = 0, 1

    model = BERT4RecModel(H, 2, 2, V, S, 0.1, initializer_range=0.2)
    model.precision = "fp32"
    names = {"recall": [1, 10], "ndcg": [10], "mrr": [10], "rank": []}
    seq, _, lengths = _random_batch(torch.Generator().manual_seed(9), B, S, V, p_mask=0.0)
    seq[torch.arange(B), lengths - 1] = 1
    batch = {"item": seq.cuda(), "item.target": torch.randint(3, V, (B,), generator=torch.Generator().manual_seed(4)).cuda()}
    rows = []
    for fused in (True, False):
        module = MaskedTrainingModule(model, item_tokenizer=Tok(), metrics=build_metrics(names)).cuda().eval()
        ev = PerSampleMetricsEvaluator(None, None, module)
        preds = module.predict_topn(batch, n, with_rank=True) if fused else module.predict_step(batch, 0).float()
        rows.append(np.asarray(ev.evaluate(0, batch, preds)))
        if fused:
            assert preds.rank is not None and preds.lse is not None and top_predictions(preds, n)[0].shape == (B, n)
    assert rows[0].shape == (B, 5)
    np.testing.assert_allclose(rows[0], rows[1], rtol=1e-6, atol=1e-7)
