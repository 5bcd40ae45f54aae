"""GPU input pipeline against the oracle's restatements of the reference processors
(data/datasets/processors/cloze_mask.py:50-92, pos_neg_sampler.py:41-114): same invariants on every sample, same distribution over
many samples (the reference's draws come from torch's global generator and cannot be reproduced on the device)."""
import numpy as np
import pytest
import torch

from oracle import asme_oracle as O

pytestmark = pytest.mark.gpu


def make_batch(gen, B, S, V, min_len=2):
    seq = torch.randint(3, V, (B, S), generator=gen)
    lengths = torch.randint(min_len, S + 1, (B,), generator=gen)
    seq[torch.arange(S).unsqueeze(0) >= lengths.unsqueeze(1)] = 0
    return seq, lengths


def test_cloze_mask_invariants_and_distribution():
    from asme_b200 import input_pipeline as ip
    gen = torch.Generator().manual_seed(0)
    B, S, V, VC, p, p_last = 4096, 60, 500, 40, 0.3, 0.15
    seq, lengths = make_batch(gen, B, S, V)
    cat = torch.randint(3, VC, (B, S), generator=gen)
    cat[seq == 0] = 0
    batch = {"item": seq.cuda(), "cat": cat.cuda()}
    out = ip.cloze_mask(batch, {"item": V, "cat": VC}, p, p_last, seed=7, masking_targets=["item", "cat"])
    x, c, t = out["item"].cpu(), out["cat"].cpu(), out["item.target"].cpu()
    valid = seq != 0
    sel = t != 0
    # invariants of cloze_mask.py:50-92 on every sample
    assert not (sel & ~valid).any()                                   # padding is never selected
    assert torch.equal(t[sel], seq[sel])                               # the target is the original item
    assert torch.equal(x[~sel], seq[~sel]) and torch.equal(c[~sel], cat[~sel])          # nothing else changes
    masked = sel & (x == 1) & (c == 1)                                 # all masking targets are masked at the same positions
    # (a random replacement may hit the MASK id in one feature only: 1 / (V - 1) of the 10 % branch)
    assert float(((sel & (x == 1)) ^ (sel & (c == 1))).float().sum()) < 0.01 * float(masked.float().sum())
    rnd_x, rnd_c = sel & (x != seq) & (x != 1), sel & (c != cat) & (c != 1)
    assert int(x[rnd_x].max()) <= V - 2 and int(c[rnd_c].max()) <= VC - 2   # random ids come from [0, len - 2]
    # sequences in "last item only" mode: exactly one selected position, the last real one, and it is a MASK
    n_sel = sel.sum(1)
    last_only = (n_sel == 1) & sel[torch.arange(B), lengths - 1] & (x[torch.arange(B), lengths - 1] == 1)
    # distribution against the oracle's per-sample restatement (torch CPU generator) on the same sequences
    torch.manual_seed(123)
    o_sel = o_mask = o_keep = o_valid = o_last = 0
    for b in range(1024):
        n = int(lengths[b])
        s_o, t_o = O.cloze_mask_sequence(seq[b, :n].tolist(), p, p_last, V)
        sel_o = [v != 0 for v in t_o]
        one_last = sum(sel_o) == 1 and sel_o[-1] and s_o[-1] == 1
        o_last += one_last
        if not one_last:
            o_sel += sum(sel_o)
            o_valid += n
            o_mask += sum(1 for v, f in zip(s_o, sel_o) if f and v == 1)
            o_keep += sum(1 for v, w, f in zip(s_o, seq[b, :n].tolist(), sel_o) if f and v == w)
    normal = ~last_only
    g_valid = int(valid[normal].sum())
    g_sel = int(sel[normal].sum())
    g_mask = int(masked[normal].sum())
    g_keep = int((sel & (x == seq))[normal].sum())
    assert abs(float(last_only.float().mean()) - o_last / 1024) < 0.04          # ~ p_last (+ the single-selection coincidences)
    assert abs(g_sel / g_valid - o_sel / o_valid) < 0.02                         # ~ mask_prob
    assert abs(g_mask / g_sel - o_mask / o_sel) < 0.02                           # ~ 0.8
    assert abs(g_keep / g_sel - o_keep / o_sel) < 0.02                           # ~ 0.1 (+ random ids that hit the original)
    assert abs(g_sel / g_valid - p) < 0.01 and abs(g_mask / g_sel - 0.8) < 0.01
    # pure function of the seed
    again = ip.cloze_mask(batch, {"item": V, "cat": VC}, p, p_last, seed=7, masking_targets=["item", "cat"])
    assert torch.equal(again["item"], out["item"]) and torch.equal(again["item.target"], out["item.target"])
    other = ip.cloze_mask(batch, {"item": V, "cat": VC}, p, p_last, seed=8, masking_targets=["item", "cat"])
    assert not torch.equal(other["item"], out["item"])


def test_cloze_mask_last_item_only_matches_reference_golden_vector():
    """tests/test_cloze_mask.py:9-21 of the reference: probability 1 of masking only the last item"""
    from asme_b200 import input_pipeline as ip
    seq = torch.tensor([[5, 8, 9, 7, 3, 4, 0, 0], [3, 0, 0, 0, 0, 0, 0, 0]])
    out = ip.cloze_mask({"item": seq.cuda()}, {"item": 13}, 1.0, 1.0, seed=1)
    assert out["item"].cpu().tolist() == [[5, 8, 9, 7, 3, 1, 0, 0], [1, 0, 0, 0, 0, 0, 0, 0]]
    assert out["item.target"].cpu().tolist() == [[0, 0, 0, 0, 0, 4, 0, 0], [3, 0, 0, 0, 0, 0, 0, 0]]


def test_pos_neg_sample_invariants_and_distribution():
    from asme_b200 import input_pipeline as ip
    gen = torch.Generator().manual_seed(1)
    B, S1, V = 2048, 51, 300
    seq, lengths = make_batch(gen, B, S1, V)
    out = ip.pos_neg_sample({"item": seq.cuda()}, V, seed=3)
    x, pos, neg = out["item"].cpu(), out["positive_samples"].cpu(), out["negative_samples"].cpu()
    S = S1 - 1
    inside = torch.arange(S).unsqueeze(0) < (lengths - 1).unsqueeze(1)
    # pos_neg_sampler.py:96-101: x = seq[:-1], pos = seq[1:], padded to the batch shape
    assert torch.equal(x[inside], seq[:, :-1][inside]) and torch.equal(pos[inside], seq[:, 1:][inside])
    assert (x[~inside] == 0).all() and (pos[~inside] == 0).all() and (neg[~inside] == 0).all()
    # :44-54: negatives are never special tokens and never a token of their own sequence
    assert int(neg[inside].min()) >= 3 and int(neg[inside].max()) < V
    clash = (neg.unsqueeze(2) == seq.unsqueeze(1)).any(dim=2) & inside
    assert not clash.any()
    # uniform over the allowed ids: compare the histogram with the oracle's multinomial on the same sequences
    torch.manual_seed(5)
    ref = []
    for b in range(512):
        n = int(lengths[b])
        ref += O.pos_neg_sequence(seq[b, :n].tolist(), V)[2]
    g = neg[:512][inside[:512]].numpy()
    hist_g = np.bincount(g, minlength=V) / len(g)
    hist_o = np.bincount(np.array(ref), minlength=V) / len(ref)
    assert len(g) == len(ref)
    assert np.abs(hist_g - hist_o).max() < 0.004 and abs(g.mean() - np.mean(ref)) < 3.0
    assert torch.equal(ip.pos_neg_sample({"item": seq.cuda()}, V, seed=3)["negative_samples"].cpu(), neg)


def test_pipeline_feeds_the_training_modules():
    """GPU cloze masking -> MaskedTrainingModule.training_step; GPU negative sampling -> SequenceNextItemPredictionTrainingModule"""
    from asme_b200 import input_pipeline as ip
    from asme_b200.models import BERT4RecModel, SASRecModel
    from asme_b200.modules import MaskedTrainingModule, SequenceNextItemPredictionTrainingModule
    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(2)
    V, S, B = 400, 32, 64
    seq, _ = make_batch(gen, B, S, V)
    batch = ip.cloze_mask({"item": seq.cuda()}, {"item": V}, 0.2, 0.1, seed=11)
    module = MaskedTrainingModule(BERT4RecModel(64, 2, 1, V, S, 0.1).cuda().train(), num_warmup_steps=0)
    loss = module.training_step(batch, 0)["loss"]
    loss.backward()
    assert torch.isfinite(loss) and 4.0 < float(loss) < 8.0
    seq1, _ = make_batch(gen, B, S + 1, V)
    nb = ip.pos_neg_sample({"item": seq1.cuda()}, V, seed=12)
    module = SequenceNextItemPredictionTrainingModule(SASRecModel(64, 2, 1, V, S, 0.1, mode="neg_sampling").cuda().train())
    loss = module.training_step(nb, 0)["loss"]
    loss.backward()
    assert torch.isfinite(loss)


def test_weighted_negatives_invariants_and_distribution():
    """metrics_sampler.py:140-204: negatives follow the item weights, are distinct, never the target, never an input item"""
    from asme_b200 import ops
    gen = torch.Generator().manual_seed(3)
    V, B, S, n = 60, 4000, 6, 10
    weights = torch.rand(V, generator=gen) + 0.05
    weights[[0, 1, 2, 7]] = 0.0                                     # special tokens and one unpopular item: never drawn
    seq = torch.tensor([[5, 9, 11, 0, 0, 0]]).repeat(B, 1)          # every user has the same history and target ...
    tgt = torch.full((B,), 13)
    cdf = torch.cumsum(weights.double(), 0).cuda()
    neg, failed = ops.weighted_negatives(cdf, seq.cuda(), tgt.cuda(), n, seed=5)
    neg = neg.cpu()
    assert int(failed.item()) == 0
    assert (neg.sort(dim=1).values.diff(dim=1) != 0).all()          # distinct within a user
    banned = torch.tensor([0, 1, 2, 7, 5, 9, 11, 13])
    assert not torch.isin(neg, banned).any()
    # ... so the inclusion frequencies can be compared with torch.multinomial (without replacement) on the same renormalised weights
    w = weights.clone()
    w[[5, 9, 11, 13, 0]] = 0.0
    torch.manual_seed(1)
    ref = torch.multinomial(w.unsqueeze(0).repeat(B, 1), n)
    f_g = np.bincount(neg.reshape(-1).numpy(), minlength=V) / B
    f_o = np.bincount(ref.reshape(-1).numpy(), minlength=V) / B
    assert np.abs(f_g - f_o).max() < 0.035                          # inclusion probabilities are O(0.2): ~4 sigma at B = 4000
    # the FIRST draw alone follows the renormalised weights exactly
    first = np.bincount(neg[:, 0].numpy(), minlength=V) / B
    assert np.abs(first - (w / w.sum()).numpy()).max() < 0.015
    # too few admissible items: the flag is raised instead of torch.multinomial's exception
    few = torch.zeros(V)
    few[[20, 21, 22]] = 1.0
    _, failed = ops.weighted_negatives(torch.cumsum(few.double(), 0).cuda(), seq[:4].cuda(), tgt[:4].cuda(), 5, seed=1)
    assert int(failed.item()) == 1


def test_sampled_and_fixed_subset_metrics_from_fused_predictions():
    """FixedItemsSampler / NegativeMetricsSampler on the fused evaluation output (no dense logits): identical to the oracle's
    metrics on the dense logits gathered at the same items"""
    from asme_b200.data import InputSequence
    from asme_b200.metrics import (FixedItemsSampler, FusedPredictions, NegativeMetricsSampler, NormalizedDiscountedCumulativeGainMetric,
                                   RankingMetricsContainer, RecallMetric)
    from asme_b200.models import BERT4RecModel, mask_position_rows
    from asme_b200 import models
    old = models.DEFAULT_PRECISION
    models.set_default_precision("fp32")
    try:
        torch.manual_seed(0)
        gen = torch.Generator().manual_seed(4)
        V, S, B = 300, 20, 64
        model = BERT4RecModel(32, 2, 1, V, S, 0.0, initializer_range=0.3).cuda().eval()
        seq, lengths = make_batch(gen, B, S, V)
        seq[torch.arange(B), lengths - 1] = 1
        tgt = torch.randint(3, V, (B,), generator=gen)
        seq_d, tgt_d = seq.cuda(), tgt.cuda()
        out = model.evaluate_rank(seq_d, seq_d.ne(0), {}, tgt_d, k=10)
        pred = FusedPredictions(out["rank"], out["topk_idx"], out["topk_val"], out["target_score"], V, scorer=out["scorer"])
        dense = model(InputSequence(seq_d, seq_d.ne(0), {})).reshape(B * S, V)[mask_position_rows(seq_d, 1)].cpu()
        fixed = list(range(3, 120, 2))
        cont = RankingMetricsContainer([RecallMetric(5), NormalizedDiscountedCumulativeGainMetric(5)], FixedItemsSampler(fixed))
        got = cont.update(seq_d, tgt_d, pred)
        items = torch.tensor(fixed).unsqueeze(0).repeat(B, 1)
        sub = dense.gather(1, items).numpy()
        pos = items.eq(tgt.unsqueeze(1)).numpy()
        assert abs(float(got["recall@5_fixed"]) - O.recall_at_k(sub, pos, 5).mean()) < 1e-6
        assert abs(float(got["NDCG@5_fixed"]) - O.ndcg_at_k(sub, pos, 5).mean()) < 1e-6
        sampler = NegativeMetricsSampler([0.0, 0.0, 0.0] + [1.0] * (V - 3), 50, "_sampled(50)", seed=9)
        cont = RankingMetricsContainer([RecallMetric(5), NormalizedDiscountedCumulativeGainMetric(5)], sampler)
        sample = sampler.sample(seq_d, tgt_d, pred)
        assert tuple(sample.sampled_predictions.shape) == (B, 51)
        sampler._calls -= 1                                             # the container's call below repeats exactly this draw
        got = cont.update(seq_d, tgt_d, pred)
        sp = sample.sampled_predictions.cpu().numpy()
        pm = sample.positive_item_mask.cpu().numpy()
        assert abs(float(got["recall@5_sampled(50)"]) - O.recall_at_k(sp, pm, 5).mean()) < 1e-6
        assert abs(float(got["NDCG@5_sampled(50)"]) - O.ndcg_at_k(sp, pm, 5).mean()) < 1e-6
        # the gathered scores are the dense logits at those items
        np.testing.assert_allclose(sub, pred.gather(1, items.cuda()).cpu().numpy(), rtol=1e-5, atol=1e-5)
        # the all-items container keeps its k values on the device: a tensor built from the Python list every step would be a pageable
        # host-to-device copy, i.e. a stream synchronisation in front of the metric kernels of every evaluation step
        from asme_b200.metrics import AllItemsSampler
        allc = RankingMetricsContainer([RecallMetric(5), NormalizedDiscountedCumulativeGainMetric(5)], AllItemsSampler())
        v1 = allc.update(seq_d, tgt_d, pred)
        ks_first = next(iter(allc._ks_device.values()))
        v2 = allc.update(seq_d, tgt_d, pred)
        assert len(allc._ks_device) == 1 and next(iter(allc._ks_device.values())) is ks_first
        assert float(v1["recall@5"]) == float(v2["recall@5"]) and float(v1["NDCG@5"]) == float(v2["NDCG@5"])
    finally:
        models.set_default_precision(old)


def test_files_to_predictions_csv_end_to_end(golden_dir):
    """either side of the hot path in one go: the reference's data files (session CSV + session index + position index + vocabulary)
    -> pre-tokenised store -> collated batch -> GPU -> SASRec next-item top-n on the fused path -> the predict evaluators -> CSV text"""
    import csv
    import io
    import os
    from asme_b200 import evaluation as E
    from asme_b200 import formats as F
    from asme_b200.models import SASRecModel
    from asme_b200.modules import NextItemPredictionTrainingModule
    d = os.path.join(golden_dir, "formats")
    vocab = F.read_vocabulary(os.path.join(d, "sessions.vocabulary.item_id.txt"))
    store = F.TokenisedSessions.from_csv(os.path.join(d, "sessions.csv"), F.read_session_index(os.path.join(d, "sessions.session.idx")), vocab)
    positions = F.read_position_index(os.path.join(d, "sessions.nextitem.idx"))
    S, n = 12, 5
    batch = {k: v.cuda() for k, v in store.batch(positions[:48, 0], positions[:48, 1], max_seq_length=S).items()}

    class Vocab:
        def tokens(self):
            return list(vocab)

        def ids(self):
            return list(vocab.values())

    class Tok:
        pad_token_id, mask_token_id, vocabulary = 0, 1, Vocab()

        def get_special_token_ids(self):
            return [0, 1, 2]

    torch.manual_seed(0)
    module = NextItemPredictionTrainingModule(SASRecModel(64, 2, 1, len(vocab), S, 0.1, mode="full"), item_tokenizer=Tok()).cuda().eval()
    preds = module.predict_topn(batch, n, with_rank=True)
    writer = E.CSVSingleLineWriter([E.ExtractSampleIdEvaluator(), E.LogInputEvaluator(Tok()), E.TrueTargetEvaluator(Tok()),
                                    E.ExtractRecommendationEvaluator(Tok(), n), E.ExtractScoresEvaluator(Tok(), n)])
    out = io.StringIO()
    writer.init_file(out)
    writer.write_evaluation(0, batch, preds)
    rows = list(csv.reader(io.StringIO(out.getvalue())))
    assert rows[0] == ["SID", "input", "target", "recommendation", "score"] and len(rows) == 49
    assert rows[1][0] == f"{int(positions[0, 0])}_{int(positions[0, 1])}"
    # the written recommendations are the tokens of the fused top-n ids, best first, and agree with the dense logits' ordering
    dense = module.predict_step(batch, 0).float()
    want = torch.topk(dense, n, dim=1).indices.cpu().numpy()
    tokens = list(vocab)
    agree = 0
    for i, row in enumerate(rows[1:]):
        rec = eval(row[3])
        assert len(rec) == n and all(t in vocab for t in rec)
        agree += rec == [tokens[j] for j in want[i]]
    assert agree >= 44                                   # bf16 near-ties may swap neighbours in a few rows
    assert ((preds.rank >= 1) & (preds.rank <= len(vocab))).all()
