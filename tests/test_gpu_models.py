"""Model-level parity (``-m gpu``): the drop-in model / module / metric classes, running on the sm_100a kernels
through the C ABI, against (1) the committed outputs of the unmodified reference (tests/golden/*.npz) and
(2) the CPU oracle on fresh seeded inputs.  fp32 tolerance: 1e-5 relative on logits and losses (north_star)."""
import os

import numpy as np
import pytest
import torch

from oracle import asme_oracle as O
from test_host_cpu import build_from_fixture

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _strict_fp32_policy():
    """these tests pin the strict-parity (fp32 SIMT) path at 1e-5; the tensor-core policy is covered by test_gpu_models_bf16.py"""
    from asme_b200 import models
    old = models.DEFAULT_PRECISION
    models.set_default_precision("fp32")
    yield
    models.set_default_precision(old)


RTOL, ATOL = 1e-5, 2e-5


def close(a, b, rtol=RTOL, atol=ATOL, msg=""):
    torch.testing.assert_close(a.detach().cpu().float(), torch.as_tensor(b).float(), rtol=rtol, atol=atol, msg=lambda m: f"{msg}: {m}")


def _grads(model):
    return {n: p.grad.detach().cpu() for n, p in model.named_parameters() if p.grad is not None}


def _check_grads(model, z, skip=()):
    got = _grads(model)
    for key in z.files:
        if not key.startswith("grad::"):
            continue
        name = key[6:]
        if name in skip or name not in got:
            assert name in skip or name.startswith("_projection_layer.embedding"), f"no gradient for {name}"
            continue
        close(got[name], z[key], rtol=2e-4, atol=2e-5, msg=f"grad {name}")


def test_bert4rec_vs_reference_fixture(golden_dir):
    from asme_b200.data import InputSequence
    z, w, model = build_from_fixture(golden_dir, "bert4rec_small.npz")
    model.load_state_dict(w)
    model = model.cuda().train()
    inp, tgt = torch.from_numpy(z["input"]).cuda(), torch.from_numpy(z["target"]).cuda()
    logits = model(InputSequence(inp, inp.ne(0), {}))
    close(logits, z["logits"], msg="logits")
    loss, ctx = model.loss_ce(inp, inp.ne(0), {}, tgt)
    close(loss, z["loss"], rtol=1e-5, atol=1e-5, msg="loss")
    model.loss_ce_backward(ctx)
    _check_grads(model, z)
    model.eval()
    ev = torch.from_numpy(z["eval_input"]).cuda()
    et = torch.from_numpy(z["eval_target"]).cuda()
    out = model.evaluate_rank(ev, ev.ne(0), {}, et, k=10, with_loss=True)
    rows = z["eval_logits"]
    close(out["target_score"], rows[np.arange(rows.shape[0]), z["eval_target"]], msg="target score")
    assert np.array_equal(out["rank"].cpu().numpy(), O.target_rank(rows, z["eval_target"]))
    assert np.array_equal(out["topk_idx"].cpu().numpy(), O.topk_ids(rows, 10))
    ce = torch.nn.functional.cross_entropy(torch.from_numpy(rows), torch.from_numpy(z["eval_target"]), ignore_index=0)
    close(out["loss"], ce, msg="eval loss")


def test_kebert4rec_vs_reference_fixture(golden_dir):
    from asme_b200.data import InputSequence
    z, w, model = build_from_fixture(golden_dir, "kebert4rec_small.npz")
    model.load_state_dict(w)
    model = model.cuda().train()
    inp, tgt = torch.from_numpy(z["input"]).cuda(), torch.from_numpy(z["target"]).cuda()
    attrs = {"category": torch.from_numpy(z["category"]).cuda(), "tags": torch.from_numpy(z["tags"]).cuda()}
    close(model(InputSequence(inp, inp.ne(0), attrs)), z["logits"], msg="logits")
    loss, ctx = model.loss_ce(inp, inp.ne(0), attrs, tgt)
    close(loss, z["loss"], msg="loss")
    model.loss_ce_backward(ctx)
    _check_grads(model, z)


def test_sasrec_full_vs_reference_fixture(golden_dir):
    from asme_b200.data import InputSequence
    z, w, model = build_from_fixture(golden_dir, "sasrec_full_small.npz")
    model.load_state_dict(w)
    model = model.cuda().train()
    inp, tgt = torch.from_numpy(z["input"]).cuda(), torch.from_numpy(z["target"]).cuda()
    close(model(InputSequence(inp, inp.ne(0), {})), z["logits"], msg="logits")
    loss, ctx = model.loss_ce(inp, inp.ne(0), {}, tgt)
    close(loss, z["loss"], msg="loss")
    model.loss_ce_backward(ctx)
    _check_grads(model, z)
    model.eval()
    et = torch.randint(3, int(z["V"]), (inp.shape[0],), generator=torch.Generator().manual_seed(0))
    out = model.evaluate_rank(inp, inp.ne(0), {}, et.cuda(), k=5, select="last")
    assert np.array_equal(out["rank"].cpu().numpy(), O.target_rank(z["eval_logits"], et.numpy()))


@pytest.mark.parametrize("name", ["kebert4rec_postfusion_add.npz", "kebert4rec_postfusion_multiply.npz", "sasrec_postfusion_add.npz",
                                  "sasrec_postfusion_multiply.npz"])
def test_postfusion_attributes_vs_reference_fixture(golden_dir, name):
    """row a9: attribute embeddings merged into the encoded sequence (add / multiply) before (KeBERT4Rec) or instead of (SASRec) the
    modifier transform -- logits, loss and every gradient (incl. the post-fused tables and the id-bag Linear) vs the unmodified
    reference (models/kebert4rec/components.py:97-116, models/sasrec/components.py:88-106)"""
    from asme_b200.data import InputSequence
    z, w, model = build_from_fixture(golden_dir, name)
    model.load_state_dict(w)
    model = model.cuda().train()
    inp, tgt = torch.from_numpy(z["input"]).cuda(), torch.from_numpy(z["target"]).cuda()
    attrs = {k: torch.from_numpy(z[k]).cuda() for k in ("category", "tags") if k in z.files}
    close(model(InputSequence(inp, inp.ne(0), attrs)), z["logits"], msg="logits")
    loss, ctx = model.loss_ce(inp, inp.ne(0), attrs, tgt)
    close(loss, z["loss"], msg="loss")
    model.loss_ce_backward(ctx)
    _check_grads(model, z)
    post = [n for n, p in model.named_parameters() if "postfusion_attribute_embeddings" in n and p.grad is not None]
    assert post and all(float(dict(model.named_parameters())[n].grad.abs().sum()) > 0 for n in post)


def test_sasrec_neg_vs_reference_fixture(golden_dir):
    from asme_b200.data import InputSequence
    z, w, model = build_from_fixture(golden_dir, "sasrec_neg_small.npz")
    model.load_state_dict(w, strict=False)
    model = model.cuda().train()
    inp = torch.from_numpy(z["input"]).cuda()
    pos, neg = torch.from_numpy(z["positive_samples"]).cuda(), torch.from_numpy(z["negative_samples"]).cuda()
    p, n = model(InputSequence(inp, inp.ne(0), {"positive_samples": pos, "negative_samples": neg}))
    close(p, z["pos_logits"], msg="pos logits")
    close(n, z["neg_logits"], msg="neg logits")
    loss, ctx = model.loss_bce(inp, inp.ne(0), {}, pos, neg, inp.ne(0))
    close(loss, z["loss"], msg="bce loss")
    model.loss_bce_backward(ctx)
    _check_grads(model, z)
    model.eval()
    ev = model(InputSequence(inp, inp.ne(0), {}))           # all-items eval branch
    close(ev, z["eval_logits"], msg="eval logits")
    items = torch.arange(int(z["V"])).repeat(inp.shape[0], 1).cuda()
    close(model(InputSequence(inp, inp.ne(0), {"positive_samples": items})), z["eval_logits"], msg="eval logits (items)")


def _user_attrs(z, category_key="category"):
    return {"user_id": torch.from_numpy(z["user_id"]).cuda(), "gender": torch.from_numpy(z["gender"]).cuda(),
            "category": torch.from_numpy(z[category_key]).cuda()}


def test_ubert4rec_vs_reference_fixture(golden_dir):
    """user token prepended (two user-attribute tables), item attribute, segment embedding, causal encoder;
    logits have S+1 positions; training / evaluation through the drop-in module exactly as the reference module slices them"""
    from asme_b200.data import InputSequence
    from asme_b200.modules import UBERTMaskedTrainingModule
    z, w, model = build_from_fixture(golden_dir, "ubert4rec_small.npz")
    model.load_state_dict(w)
    model = model.cuda().train()
    inp, tgt = torch.from_numpy(z["input"]).cuda(), torch.from_numpy(z["target"]).cuda()
    attrs = _user_attrs(z)
    logits = model(InputSequence(inp, inp.ne(0), attrs))
    assert tuple(logits.shape) == tuple(z["logits"].shape)
    close(logits, z["logits"], msg="logits")
    module = UBERTMaskedTrainingModule(model, num_warmup_steps=0)
    module.fused_eval = True
    out = module.training_step({"item": inp, "item.target": tgt, **attrs}, 0)
    close(out["loss"], z["loss"], rtol=1e-5, atol=1e-5, msg="loss")
    out["loss"].backward()
    _check_grads(model, z)
    model.eval()
    ev = torch.from_numpy(z["eval_input"]).cuda()
    rows = z["eval_logits"]
    et = torch.randint(3, int(z["V"]), (ev.shape[0],), generator=torch.Generator().manual_seed(1))
    res = model.evaluate_rank(ev, ev.ne(0), _user_attrs(z, "eval_category"), et.cuda(), k=10, rows=module._mask_rows(ev))
    close(res["target_score"], rows[np.arange(rows.shape[0]), et.numpy()], msg="target score")
    assert np.array_equal(res["rank"].cpu().numpy(), O.target_rank(rows, et.numpy()))
    assert np.array_equal(res["topk_idx"].cpu().numpy(), O.topk_ids(rows, 10))
    pred = module._get_prediction_for_masked_item({"item": ev, **_user_attrs(z, "eval_category")}, 0)
    close(pred, rows, msg="masked-item logits")


def test_usasrec_full_vs_reference_fixture(golden_dir):
    from asme_b200.data import InputSequence
    from asme_b200.modules import UserNextItemPredictionTrainingModule
    z, w, model = build_from_fixture(golden_dir, "usasrec_full_small.npz")
    model.load_state_dict(w)
    model = model.cuda().train()
    inp, tgt = torch.from_numpy(z["input"]).cuda(), torch.from_numpy(z["target"]).cuda()
    attrs = _user_attrs(z)
    close(model(InputSequence(inp, inp.ne(0), attrs)), z["logits"], msg="logits")
    module = UserNextItemPredictionTrainingModule(model)
    out = module.training_step({"item": inp, "item.target": tgt, **attrs}, 0)
    close(out["loss"], z["loss"], rtol=1e-5, atol=1e-5, msg="loss")
    out["loss"].backward()
    _check_grads(model, z)
    model.eval()
    close(module.predict_step({"item": inp, **attrs}, 0), z["eval_logits"], msg="eval rows (reference row choice)")
    et = torch.randint(3, int(z["V"]), (inp.shape[0],), generator=torch.Generator().manual_seed(0))
    res = model.evaluate_rank(inp, inp.ne(0), attrs, et.cuda(), k=5, rows=module._target_rows(inp, inp.ne(0)))
    assert np.array_equal(res["rank"].cpu().numpy(), O.target_rank(z["eval_logits"], et.numpy()))


@pytest.mark.parametrize("name", ["kebert4rec_basket_max.npz", "kebert4rec_basket_sum.npz", "kebert4rec_basket_mean.npz", "sasrec_basket_mean.npz"])
def test_basket_inputs_vs_reference_fixture(golden_dir, name):
    """basket inputs (N,S,BS) with embedding_pooling_type max / sum / mean (models/common/layers/sequence_embedding.py:9-45, :83-93; padding
    mask from the step maximum, modules/util/module_util.py:25-29): logits, loss and every gradient -- incl. the item table through the
    pooling (max: the slot that held the maximum) -- vs the unmodified reference"""
    from asme_b200.data import InputSequence
    from asme_b200.modules import get_padding_mask
    z, w, model = build_from_fixture(golden_dir, name)
    model.load_state_dict(w)
    model = model.cuda().train()
    inp, tgt = torch.from_numpy(z["input"]).cuda(), torch.from_numpy(z["target"]).cuda()
    assert inp.dim() == 3
    pm = get_padding_mask(inp, 0)
    attrs = {"category": torch.from_numpy(z["category"]).cuda()} if "category" in z.files else {}
    close(model(InputSequence(inp, pm, attrs)), z["logits"], msg="logits")
    loss, ctx = model.loss_ce(inp, pm, attrs, tgt)
    close(loss, z["loss"], msg="loss")
    model.loss_ce_backward(ctx)
    _check_grads(model, z)


def test_usasrec_first_item_full_vs_reference_fixture(golden_dir):
    """UserSASRec with replace_first_item=True (the user token overwrites position 0, models/ubert4rec/components.py:117-121) and a
    ``user_linear_upscale`` user attribute (:12-44: Linear over the multi-hot of a list of ids, id 0 included); the module's
    ``first_item=True`` keeps all S logit rows (user_next_item_prediction_training_module.py:57-60).  Logits, loss, every gradient
    (incl. the transposed Linear of the id list and its bias) and the evaluation rows vs the unmodified reference."""
    from asme_b200.data import InputSequence
    from asme_b200.modules import UserNextItemPredictionTrainingModule
    z, w, model = build_from_fixture(golden_dir, "usasrec_first_item_full.npz")
    model.load_state_dict(w)
    model = model.cuda().train()
    inp, tgt = torch.from_numpy(z["input"]).cuda(), torch.from_numpy(z["target"]).cuda()
    attrs = _user_attrs(z)
    logits = model(InputSequence(inp, inp.ne(0), attrs))
    assert tuple(logits.shape) == tuple(z["logits"].shape)          # S positions, not S + 1
    close(logits, z["logits"], msg="logits")
    with pytest.raises(ValueError):
        UserNextItemPredictionTrainingModule(model, first_item=False)
    module = UserNextItemPredictionTrainingModule(model, first_item=True)
    out = module.training_step({"item": inp, "item.target": tgt, **attrs}, 0)
    close(out["loss"], z["loss"], rtol=1e-5, atol=1e-5, msg="loss")
    out["loss"].backward()
    _check_grads(model, z)
    model.eval()
    close(module.predict_step({"item": inp, **attrs}, 0), z["eval_logits"], msg="eval rows")
    et = torch.randint(3, int(z["V"]), (inp.shape[0],), generator=torch.Generator().manual_seed(0))
    res = model.evaluate_rank(inp, inp.ne(0), attrs, et.cuda(), k=5, rows=module._target_rows(inp, inp.ne(0)))
    assert np.array_equal(res["rank"].cpu().numpy(), O.target_rank(z["eval_logits"], et.numpy()))
    assert np.array_equal(res["topk_idx"].cpu().numpy(), O.topk_ids(z["eval_logits"], 5))


def test_usasrec_first_item_neg_sampling_vs_reference_fixture(golden_dir):
    """UserSASRecModel(mode="neg_sampling") -- the reference's default mode: products with the positive / negative item embeddings
    (models/user_sasrec/components.py:10-60), BCE loss, item-subset scores at the last position"""
    from asme_b200.data import InputSequence
    z, w, model = build_from_fixture(golden_dir, "usasrec_first_item_neg.npz")
    model.load_state_dict(w)
    model = model.cuda().train()
    inp = torch.from_numpy(z["input"]).cuda()
    attrs = _user_attrs(z)
    pos, neg = torch.from_numpy(z["positive"]).cuda(), torch.from_numpy(z["negative"]).cuda()
    pl, nl = model(InputSequence(inp, inp.ne(0), dict(attrs, positive_samples=pos, negative_samples=neg)))
    close(pl, z["pos_logits"], msg="positive logits")
    close(nl, z["neg_logits"], msg="negative logits")
    loss, ctx = model.loss_bce(inp, inp.ne(0), attrs, pos, neg, inp.ne(0))
    close(loss, z["loss"], rtol=1e-5, atol=1e-5, msg="loss")
    model.loss_bce_backward(ctx)
    _check_grads(model, z)
    model.eval()
    items = torch.from_numpy(z["eval_items"]).cuda()
    close(model(InputSequence(inp, inp.ne(0), dict(attrs, positive_samples=items))), z["eval_logits"], msg="item-subset scores")
    from asme_b200.models import UserSASRecModel
    with pytest.raises(ValueError):      # a PREPENDED user token cannot be combined with the sampled projection (shapes, as in the reference)
        UserSASRecModel(16, 2, 1, 59, 9, 0.0, user_attributes={"gender": {"embedding_type": "content_embedding"}},
                        attribute_vocab_sizes={"gender": 7}, mode="neg_sampling")


# ------------------------------------------------------------------------------------------------------------
# fresh seeded inputs vs the oracle at the BASELINE.json shapes (reduced batch so that the CPU oracle takes seconds)
# ------------------------------------------------------------------------------------------------------------
def _random_batch(gen, B, S, V, p_mask=0.2):
    seq = torch.randint(3, V, (B, S), generator=gen)
    lengths = torch.randint(max(2, S // 10), S + 1, (B,), generator=gen)
    target = torch.zeros_like(seq)
    for i in range(B):
        n = int(lengths[i])
        seq[i, n:] = 0
        m = torch.rand(n, generator=gen) < p_mask
        m[n - 1] = True
        target[i, :n][m] = seq[i, :n][m]
        seq[i, :n][m] = 1
    return seq, target, lengths


def _cpu_weights(model, drop=("_projection_layer.embedding",)):
    return {k: v.detach().cpu().clone() for k, v in model.state_dict().items() if not k.startswith(drop)}


def test_bert4rec_c2_shape_vs_oracle():
    """C2: V=3709, S=200, H=64, L=2, heads=2 (batch 16 instead of 256 for the CPU oracle)"""
    from asme_b200.models import BERT4RecModel
    torch.manual_seed(0)
    V, S, H, B = 3709, 200, 64, 16
    model = BERT4RecModel(H, 2, 2, V, S, 0.0, initializer_range=0.1).cuda().train()
    w = _cpu_weights(model)
    seq, target, _ = _random_batch(torch.Generator().manual_seed(1235), B, S, V)
    loss, ctx = model.loss_ce(seq.cuda(), seq.cuda().ne(0), {}, target.cuda())
    model.loss_ce_backward(ctx)
    leaves = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    ref = O.cross_entropy_ignore_pad(O.bert4rec_logits(leaves, seq, 2, 2), target)
    ref.backward()
    close(loss, ref.detach(), msg="C2 loss")
    got = _grads(model)
    for name, leaf in leaves.items():
        close(got[name], leaf.grad, rtol=1e-3, atol=1e-6, msg=f"C2 grad {name}")


def test_kebert4rec_c3_shape_vs_oracle():
    """C3: V=12104, S=50, H=64, attributes category (content_embedding, 256) + tags (linear_upscale, 512, 4 ids)"""
    from asme_b200.models import KeBERT4RecModel
    torch.manual_seed(0)
    V, S, H, B = 12104, 50, 64, 32
    gen = torch.Generator().manual_seed(1236)
    model = KeBERT4RecModel(H, 2, 2, V, S, 0.0, initializer_range=0.1,
                            prefusion_attributes={"category": {"embedding_type": "content_embedding"},
                                                  "tags": {"embedding_type": "linear_upscale"}},
                            attribute_vocab_sizes={"category": 256, "tags": 512}).cuda().train()
    w = _cpu_weights(model)
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
    close(loss, ref.detach(), msg="C3 loss")
    got = _grads(model)
    for name, leaf in leaves.items():
        close(got[name], leaf.grad, rtol=1e-3, atol=1e-6, msg=f"C3 grad {name}")


def test_sasrec_neg_c4_shape_vs_oracle():
    """C4: V=13047, S=50, H=64, neg_sampling BCE (batch 64 instead of 1024)"""
    from asme_b200.models import SASRecModel
    torch.manual_seed(0)
    V, S, H, B = 13047, 50, 64, 64
    gen = torch.Generator().manual_seed(1237)
    model = SASRecModel(H, 2, 2, V, S, 0.0, mode="neg_sampling").cuda().train()
    w = _cpu_weights(model, drop=("_projection_layer",))
    seq, _, lengths = _random_batch(gen, B, S, V, p_mask=0.0)
    seq[seq == 1] = 7
    pos, neg = torch.randint(3, V, (B, S), generator=gen), torch.randint(3, V, (B, S), generator=gen)
    pos[seq == 0] = 0
    neg[seq == 0] = 0
    loss, ctx = model.loss_bce(seq.cuda(), seq.cuda().ne(0), {}, pos.cuda(), neg.cuda(), seq.cuda().ne(0))
    model.loss_bce_backward(ctx)
    leaves = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    p, n = O.sasrec_neg_logits(leaves, seq, pos, neg, 2, 2)
    ref = O.sasrec_bce(p, n, seq.ne(0))
    ref.backward()
    close(loss, ref.detach(), msg="C4 loss")
    got = _grads(model)
    for name, leaf in leaves.items():
        close(got[name], leaf.grad, rtol=1e-3, atol=1e-6, msg=f"C4 grad {name}")


def test_training_modules_and_metrics_end_to_end():
    """MaskedTrainingModule drives the fused path through the Lightning-style hooks; metric values equal the oracle's
    dense computation on the materialised logits of the same model."""
    from asme_b200.models import BERT4RecModel
    from asme_b200.modules import MaskedTrainingModule
    from asme_b200.metrics import build_metrics
    from asme_b200.data import InputSequence
    torch.manual_seed(0)
    V, S, H, B = 500, 20, 32, 48
    model = BERT4RecModel(H, 2, 1, V, S, 0.1, initializer_range=0.3)
    module = MaskedTrainingModule(model, metrics=build_metrics({"recall": [1, 5, 10], "ndcg": [5, 10], "mrr": [10], "precision": [5],
                                                               "f1": [5], "rank": [], "mrr_full": []}),
                                  num_warmup_steps=2).cuda()
    gen = torch.Generator().manual_seed(5)
    seq, target, lengths = _random_batch(gen, B, S - 1, V)
    seq = torch.nn.functional.pad(seq, (0, 1))
    target = torch.nn.functional.pad(target, (0, 1))
    batch = {"item": seq.cuda(), "item.target": target.cuda()}
    (optimizer,), (sched,) = module.configure_optimizers()
    module.train()
    losses = []
    for step in range(8):
        optimizer.zero_grad()
        out = module.training_step(batch, step)
        out["loss"].backward()
        optimizer.step()
        sched["scheduler"].step()
        losses.append(float(out["loss"]))
    assert losses[-1] < losses[0], losses            # dropout 0.1 on, loss still goes down on a fixed batch
    # evaluation: one MASK appended per row
    module.eval()
    ev = seq.clone()
    ev[ev == 1] = 9
    for i in range(B):
        ev[i, lengths[i]] = 1
    tgt = torch.randint(3, V, (B,), generator=gen)
    ebatch = {"item": ev.cuda(), "item.target": tgt.cuda()}
    step_values = module.validation_step_end(module.validation_step(ebatch, 0))
    result = module.validation_epoch_end(None)
    logits = model(InputSequence(ev.cuda(), ev.cuda().ne(0), {}))[ev.cuda().eq(1)].cpu().numpy()
    pm = O.multi_hot(logits.shape, tgt.numpy())
    want = {"recall@1": O.recall_at_k(logits, pm, 1), "recall@5": O.recall_at_k(logits, pm, 5), "recall@10": O.recall_at_k(logits, pm, 10),
            "NDCG@5": O.ndcg_at_k(logits, pm, 5), "NDCG@10": O.ndcg_at_k(logits, pm, 10), "MRR@10": O.mrr_at_k(logits, pm, 10),
            "precision@5": O.precision_at_k(logits, pm, 5), "F1@5": O.f1_at_k(logits, pm, 5),
            "rank": O.rank_full(logits, pm).astype(np.float32), "MRR": 1.0 / O.rank_full(logits, pm)}
    assert set(result) == set(want)
    for name, v in want.items():
        assert abs(result[name].item() - float(np.mean(v))) < 1e-5, name
        assert abs(step_values[name].item() - float(np.mean(v))) < 1e-5, name
    assert "val_loss" in module.logged


def test_dense_metric_signature_runs_reference_golden_vectors(golden_dir):
    """the reference's own metric tests (tests/test_{recall,ndcg,dcg,mrr,precision,f1}.py), replayed against the
    drop-in metric classes through their dense update(predictions, positive_item_mask) signature"""
    import json
    from asme_b200 import metrics as M
    cls = {"recall": M.RecallMetric, "ndcg": M.NormalizedDiscountedCumulativeGainMetric, "dcg": M.DiscountedCumulativeGainMetric,
           "mrr": M.MRRMetric, "precision": M.PrecisionMetric, "f1": M.F1Metric}
    with open(os.path.join(golden_dir, "metric_vectors.json")) as f:
        vectors = json.load(f)
    n = 0
    for name, samples in vectors.items():
        for s in samples:
            metric = cls[name](k=s["k"])
            metric.update(torch.tensor(s["predictions"], dtype=torch.float32).cuda(), torch.tensor(s["positive_mask"]).cuda())
            assert abs(metric.compute().item() - s["expected"]) < 10e-4, (name, s)
            n += 1
    assert n == 87
    z = np.load(os.path.join(golden_dir, "metrics_dense_small.npz"))
    pred, tg = torch.from_numpy(z["predictions"]).cuda(), torch.from_numpy(z["targets"]).cuda()
    spec = {"recall": [1, 3, 5, 10], "ndcg": [1, 3, 5, 10], "mrr": [1, 3, 5, 10], "precision": [1, 3, 5, 10], "f1": [1, 3, 5, 10],
            "dcg": [1, 3, 5, 10], "mrr_full": [], "rank": []}
    container = M.build_metrics(spec)
    step = container.update(None, tg, pred)
    final = container.compute()
    for key in z.files:
        if key.startswith("final::"):
            assert abs(final[key[7:]].item() - float(z[key])) < 1e-6, key
        if key.startswith("step::"):
            assert abs(step[key[6:]].item() - float(z[key])) < 1e-6, key
