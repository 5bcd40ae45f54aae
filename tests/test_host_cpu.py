"""CPU-side tests of the host logic: state-dict compatibility with the reference checkpoints (fixture keys and
shapes), arena packing, plug-in registration table, metric bookkeeping that needs no kernel."""
import os

import numpy as np
import pytest
import torch

from asme_b200.models import BERT4RecModel, KeBERT4RecModel, SASRecModel
from asme_b200 import plugin


def _fixture_weights(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    return z, {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w::")}


def build_from_fixture(golden_dir, name):
    z, w = _fixture_weights(golden_dir, name)
    kw = dict(transformer_hidden_size=int(z["H"]), num_transformer_heads=int(z["heads"]), num_transformer_layers=int(z["L"]),
              item_vocab_size=int(z["V"]), max_seq_length=int(z["S"]), transformer_dropout=0.0)
    post = "_sequence_representation_modifier_layer.postfusion_attribute_embeddings"
    if name.startswith("bert4rec"):
        model = BERT4RecModel(**kw)
    elif name.startswith("kebert4rec_postfusion") or name.startswith("sasrec_postfusion"):
        merge = name.split("_")[2].split(".")[0]
        if name.startswith("kebert4rec"):
            model = KeBERT4RecModel(prefusion_attributes={"category": {"embedding_type": "content_embedding"}},
                                    postfusion_attributes={"category": {"embedding_type": "content_embedding"},
                                                           "tags": {"embedding_type": "linear_upscale"}},
                                    postfusion_merge_function=merge,
                                    attribute_vocab_sizes={"category": w[f"{post}.category.weight"].shape[0],
                                                           "tags": w[f"{post}.tags.linear.weight"].shape[1]}, **kw)
        else:
            model = SASRecModel(mode="full", postfusion_attributes={"category": {"embedding_type": "content_embedding"}},
                                postfusion_merge_function=merge,
                                attribute_vocab_sizes={"category": w[f"{post}.category.weight"].shape[0]}, **kw)
    elif name.startswith("kebert4rec_basket"):
        model = KeBERT4RecModel(prefusion_attributes={"category": {"embedding_type": "content_embedding"}},
                                embedding_pooling_type=name.split("_")[2].split(".")[0],
                                attribute_vocab_sizes={"category": w["_sequence_embedding_layer.prefusion_attribute_embeddings.category.weight"].shape[0]},
                                **kw)
    elif name.startswith("sasrec_basket"):
        model = SASRecModel(mode="full", embedding_pooling_type="mean", **kw)
    elif name.startswith("kebert4rec"):
        model = KeBERT4RecModel(prefusion_attributes={"category": {"embedding_type": "content_embedding"},
                                                      "tags": {"embedding_type": "linear_upscale"}},
                                attribute_vocab_sizes={"category": w["_sequence_embedding_layer.prefusion_attribute_embeddings.category.weight"].shape[0],
                                                       "tags": w["_sequence_embedding_layer.prefusion_attribute_embeddings.tags.linear.weight"].shape[1]},
                                **kw)
    elif name.startswith("usasrec_first_item"):
        from asme_b200.models import UserSASRecModel
        emb = "_sequence_embedding_layer"
        sizes = {"user_id": w[f"{emb}.user_attribute_embeddings.user_id.linear.weight"].shape[1],
                 "gender": w[f"{emb}.user_attribute_embeddings.gender.weight"].shape[0],
                 "category": w[f"{emb}.additional_attribute_embeddings.category.weight"].shape[0]}
        model = UserSASRecModel(additional_attributes={"category": {"embedding_type": "content_embedding"}},
                                user_attributes={"user_id": {"embedding_type": "user_linear_upscale"},
                                                 "gender": {"embedding_type": "content_embedding"}},
                                attribute_vocab_sizes=sizes, replace_first_item=True,
                                mode="full" if name.endswith("full.npz") else "neg_sampling", **kw)
    elif name.startswith("ubert4rec") or name.startswith("usasrec"):
        from asme_b200.models import UBERT4RecModel, UserSASRecModel
        emb = "_sequence_embedding_layer"
        sizes = {"user_id": w[f"{emb}.user_attribute_embeddings.user_id.weight"].shape[0],
                 "gender": w[f"{emb}.user_attribute_embeddings.gender.weight"].shape[0],
                 "category": w[f"{emb}.additional_attribute_embeddings.category.weight"].shape[0]}
        ukw = dict(additional_attributes={"category": {"embedding_type": "content_embedding"}},
                   user_attributes={"user_id": {"embedding_type": "user_embedding"}, "gender": {"embedding_type": "content_embedding"}},
                   attribute_vocab_sizes=sizes, **kw)
        model = UBERT4RecModel(segment_embedding=True, **ukw) if name.startswith("ubert4rec") else UserSASRecModel(mode="full", **ukw)
    elif name.startswith("sasrec_full"):
        model = SASRecModel(mode="full", **kw)
    else:
        model = SASRecModel(mode="neg_sampling", **kw)
    return z, w, model


@pytest.mark.parametrize("name", ["bert4rec_small.npz", "kebert4rec_small.npz", "sasrec_full_small.npz", "sasrec_neg_small.npz",
                                  "ubert4rec_small.npz", "usasrec_full_small.npz", "usasrec_first_item_full.npz", "usasrec_first_item_neg.npz",
                                  "kebert4rec_postfusion_add.npz",
                                  "kebert4rec_postfusion_multiply.npz", "sasrec_postfusion_add.npz", "sasrec_postfusion_multiply.npz",
                                  "kebert4rec_basket_max.npz", "sasrec_basket_mean.npz"])
def test_state_dict_is_checkpoint_compatible(golden_dir, name):
    """every key of the reference's state_dict exists with the same shape, and loads strictly"""
    z, w, model = build_from_fixture(golden_dir, name)
    sd = model.state_dict()
    for k, v in w.items():
        assert k in sd, f"missing reference key {k}"
        assert tuple(sd[k].shape) == tuple(v.shape), k
    if not name.startswith("sasrec_neg"):      # that fixture omits the aliased _projection_layer.* duplicates
        assert set(sd) == set(w)
        model.load_state_dict(w, strict=True)
    else:
        model.load_state_dict(w, strict=False)
    for k, v in w.items():
        assert torch.equal(model.state_dict()[k], v), k
    assert model.arena_is_intact()


def test_arena_layout_glues_qkv_and_layernorm_pairs():
    m = BERT4RecModel(16, 2, 2, 50, 12, 0.0)
    pre = "_sequence_representation_layer.transformer_layer.transformer_blocks.0"
    wqkv = m.weights_span(f"{pre}.attention.linear_layers.0.weight", f"{pre}.attention.linear_layers.2.weight", (48, 16))
    sd = m.state_dict()
    assert torch.equal(wqkv, torch.cat([sd[f"{pre}.attention.linear_layers.{i}.weight"] for i in range(3)]))
    gb = m.weights_span(f"{pre}.input_sublayer.norm.weight", f"{pre}.input_sublayer.norm.bias", (2, 16))
    assert torch.equal(gb[0], sd[f"{pre}.input_sublayer.norm.weight"]) and torch.equal(gb[1], sd[f"{pre}.input_sublayer.norm.bias"])
    # tied projection: same storage, both keys present
    assert sd["_projection_layer.embedding.weight"].data_ptr() == sd["_sequence_embedding_layer.item_embedding.embedding.weight"].data_ptr()
    # every parameter is 256-byte aligned inside the flat arena unless glued
    assert m._arena.flat.numel() % 64 == 0


def test_arena_survives_module_apply():
    m = SASRecModel(16, 2, 1, 50, 12, 0.0, mode="full")
    before = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.double().float()          # _apply replaces .data -> arena must be re-packed
    assert m.arena_is_intact()
    for k, v in before.items():
        assert torch.equal(m.state_dict()[k], v)
    opt_params = list(m.parameters())
    m.attach_grads()
    assert all(p.grad is not None and p.grad.shape == p.shape for p in opt_params)


def test_linear_upscale_weight_is_stored_transposed():
    m = KeBERT4RecModel(16, 2, 1, 50, 12, 0.0, prefusion_attributes={"tags": {"embedding_type": "linear_upscale"}},
                        attribute_vocab_sizes={"tags": 17})
    p = dict(m.named_parameters())["_sequence_embedding_layer.prefusion_attribute_embeddings.tags.linear.weight"]
    assert tuple(p.shape) == (16, 17) and not p.is_contiguous()          # reference layout, (Va,H) storage underneath
    assert m.weight("_sequence_embedding_layer.prefusion_attribute_embeddings.tags.linear.weight").shape == (17, 16)
    assert m.required_metadata_keys() == ["tags"]


def test_constructor_errors_match_reference():
    with pytest.raises(KeyError):
        BERT4RecModel(16, 2, 1, 50, 12, 0.0, project_layer_type="nope")
    with pytest.raises(Exception, match="unknown projection mode"):
        SASRecModel(16, 2, 1, 50, 12, 0.0, mode="nope")
    with pytest.raises(AssertionError):
        BERT4RecModel(10, 3, 1, 50, 12, 0.0)


def test_plugin_registration_table():
    assert set(plugin.REGISTRATIONS) == {"bert4rec", "kebert4rec", "sasrec-cross", "sasrec-neg", "ubert4rec", "user-sasrec-full"}
    mod_cls, model_cls = plugin.REGISTRATIONS["ubert4rec"]
    assert mod_cls.__name__ == "UBERTMaskedTrainingModule" and model_cls.__name__ == "UBERT4RecModel"
    mod_cls, model_cls = plugin.REGISTRATIONS["sasrec-neg"]
    assert mod_cls.__name__ == "SequenceNextItemPredictionTrainingModule" and model_cls.__name__ == "SASRecModel"


def test_missing_metadata_raises_like_reference():
    from asme_b200.modules import get_additional_meta_data
    m = KeBERT4RecModel(16, 2, 1, 50, 12, 0.0, prefusion_attributes={"cat": {"embedding_type": "content_embedding"}},
                        attribute_vocab_sizes={"cat": 9})
    with pytest.raises(Exception, match="does not contain the following additional metadata: cat"):
        get_additional_meta_data(m, {"item": torch.zeros(1, 2, dtype=torch.long)})


def test_model_forward_without_gpu_fails_loudly():
    from asme_b200.data import InputSequence
    m = BERT4RecModel(16, 2, 1, 50, 12, 0.0)
    seq = torch.randint(3, 50, (2, 12))
    with pytest.raises(RuntimeError, match="CUDA"):
        m(InputSequence(seq, seq.ne(0), {}))


def test_user_models_expose_user_keys_and_extra_position():
    """models/ubert4rec/ubert4rec_model.py:36-50, :88-92: user attributes are required AND optional metadata, and reserve one more
    position; the item-side tables keep the reference's names"""
    from asme_b200.models import UBERT4RecModel, UserSASRecModel
    kw = dict(additional_attributes={"cat": {"embedding_type": "content_embedding"}},
              user_attributes={"uid": {"embedding_type": "user_embedding"}}, attribute_vocab_sizes={"cat": 9, "uid": 5})
    m = UBERT4RecModel(16, 2, 1, 50, 12, 0.0, segment_embedding=True, **kw)
    assert m.required_metadata_keys() == ["uid", "cat"] and m.optional_metadata_keys() == ["uid"] and m.user_prefix == 1
    sd = m.state_dict()
    assert tuple(sd["_sequence_embedding_layer.item_embedding_layer.position_embedding.weight"].shape) == (13, 16)
    assert tuple(sd["_sequence_embedding_layer.segment_embedding.weight"].shape) == (2, 16)
    assert "_sequence_representation_layer.transformer_encoder.transformer_blocks.0.attention.output_linear.weight" in sd
    assert not m.cfg.bidirectional                      # the reference builds UBERT4Rec's encoder with bidirectional=False
    with pytest.raises(ValueError):       # sampled projection + PREPENDED user token: (N,S,H) x (N,S+1,H), fails in the reference too
        UserSASRecModel(16, 2, 1, 50, 12, 0.0, mode="neg_sampling", **kw)
    n = UserSASRecModel(16, 2, 1, 50, 12, 0.0, mode="neg_sampling", replace_first_item=True, **kw)
    assert n.user_prefix == 0 and n.projection_kind == "sasrec_neg" and "_projection_layer.embedding.item_embedding.embedding.weight" in n.state_dict()
    with pytest.raises(ValueError):       # S+1 segment ids for S positions
        UserSASRecModel(16, 2, 1, 50, 12, 0.0, mode="full", replace_first_item=True, segment_embedding=True,
                        user_attributes={"uid": {"embedding_type": "user_embedding"}, "g": {"embedding_type": "user_embedding"}},
                        attribute_vocab_sizes={"uid": 5, "g": 4})
    u = UBERT4RecModel(16, 2, 1, 50, 12, 0.0, user_attributes={"uid": {"embedding_type": "user_linear_upscale"}},
                       attribute_vocab_sizes={"uid": 5})
    sd = u.state_dict()          # models/ubert4rec/components.py:12-44: nn.Linear(vocab, hidden)
    assert tuple(sd["_sequence_embedding_layer.user_attribute_embeddings.uid.linear.weight"].shape) == (16, 5)
    assert tuple(sd["_sequence_embedding_layer.user_attribute_embeddings.uid.linear.bias"].shape) == (16,)
    with pytest.raises(NotImplementedError):
        UBERT4RecModel(16, 2, 1, 50, 12, 0.0, user_attributes={"uid": {"embedding_type": "linear_upscale"}}, attribute_vocab_sizes={"uid": 5})


def test_fixed_items_sampler_on_dense_predictions():
    """metrics_sampler.py:74-107 -- the sampler works on dense (N,I) predictions without a GPU (pure index plumbing)"""
    from asme_b200.metrics import FixedItemsSampler
    pred = torch.arange(24, dtype=torch.float32).view(3, 8)
    targets = torch.tensor([2, 5, 7])
    s = FixedItemsSampler([1, 2, 5]).sample(None, targets, pred)
    assert s.sampled_predictions.tolist() == [[1.0, 2.0, 5.0], [9.0, 10.0, 13.0], [17.0, 18.0, 21.0]]
    assert s.positive_item_mask.tolist() == [[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [0.0, 0.0, 0.0]]
    assert FixedItemsSampler([1]).suffix_metric_name() == "_fixed"


class _Vocab:
    def __init__(self, n):
        self._tokens = ["<PAD>", "<MASK>", "<UNK>"] + [f"item_{i}" for i in range(3, n)]

    def tokens(self):
        return list(self._tokens)

    def ids(self):
        return list(range(len(self._tokens)))


class _Tok:
    def __init__(self, n):
        self.vocabulary = _Vocab(n)

    def get_special_token_ids(self):
        return [0, 1, 2]


def test_predict_evaluators_dense_and_fused_agree():
    """evaluation/evaluation.py:43-236 -- headers, sample-wise flags and outputs; the fused input (top-n list + log-sum-exp) gives
    the same recommendations and softmax scores as softmax + sort on the dense logits (the reference arithmetic)"""
    from asme_b200 import evaluation as E
    from asme_b200.metrics import FusedPredictions
    gen = torch.Generator().manual_seed(3)
    N, V, n = 6, 40, 5
    logits = torch.randn(N, V, generator=gen) * 2
    tok = _Tok(V)
    batch = {"item": torch.tensor([[5, 7, 1, 0], [9, 1, 0, 0], [3, 4, 6, 1], [8, 1, 0, 0], [11, 12, 1, 0], [2, 30, 1, 0]]),
             "item.target": torch.tensor([4, 8, 15, 16, 23, 39]), "sample_ids": torch.arange(10, 16), "pos": torch.arange(6)}
    val, idx = torch.topk(logits, n, dim=1)
    fused = FusedPredictions(None, idx.to(torch.int32), val, None, V, lse=torch.logsumexp(logits, dim=1))
    assert fused.size() == torch.Size([N, V])
    rec, sc = E.ExtractRecommendationEvaluator(tok, n), E.ExtractScoresEvaluator(tok, n)
    assert rec.get_header() == ["recommendation"] and sc.get_header() == ["score"] and not rec.eval_samplewise()
    want_scores, want_idx = torch.sort(torch.softmax(logits, dim=-1), dim=-1, descending=True)
    want_rec = [[tok.vocabulary.tokens()[j] for j in row[:n].tolist()] for row in want_idx]
    assert rec.evaluate(0, batch, logits) == want_rec and rec.evaluate(0, batch, fused) == want_rec
    np.testing.assert_allclose(np.asarray(sc.evaluate(0, batch, logits)), want_scores[:, :n].numpy(), rtol=1e-6)
    np.testing.assert_allclose(np.asarray(sc.evaluate(0, batch, fused)), want_scores[:, :n].numpy(), rtol=1e-5)
    # selected items filter what is left AFTER the cut to n predictions (:181-185): rows may become shorter than n
    selected = want_idx[0, :2].tolist() + want_idx[1, 3:9].tolist()
    recf = E.ExtractRecommendationEvaluator(tok, n, selected_items=selected)
    got_dense, got_fused = recf.evaluate(0, batch, logits), recf.evaluate(0, batch, fused)
    assert got_dense == got_fused and got_dense[0][:2] == want_rec[0][:2]
    assert all(len(r) <= n for r in got_dense)
    scf = E.ExtractScoresEvaluator(tok, n, selected_items=selected)
    assert [len(r) for r in scf.evaluate(0, batch, fused)] == [len(r) for r in got_dense]
    # the bookkeeping evaluators
    assert E.LogInputEvaluator(tok).evaluate(0, batch, fused)[0] == ["item_5", "item_7"]
    assert E.TrueTargetEvaluator(tok).evaluate(0, batch, fused)[2] == ["item_15"]
    assert E.ExtractSampleIdEvaluator().evaluate(0, batch, fused) == [f"{10 + i}_{i}" for i in range(N)]
    assert E.ExtractSampleIdEvaluator().evaluate(0, {"sample_ids": torch.arange(6)}, logits) == list(range(6))
    with pytest.raises(RuntimeError, match="holds 5 items"):
        E.ExtractScoresEvaluator(tok, 8).evaluate(0, batch, fused)
    with pytest.raises(RuntimeError, match="log-sum-exp"):
        E.ExtractScoresEvaluator(tok, 3).evaluate(0, batch, FusedPredictions(None, idx.to(torch.int32), val, None, V))


def test_predict_writers_reproduce_the_reference_csv(golden_dir):
    """asme/core/writer/prediction/batch_prediction_writer.py: CSV text of the UNMODIFIED reference evaluators + writers on seeded
    logits (tests/golden/make_predict_golden.py).  Dense input: identical text.  Fused input (top-n list + log-sum-exp): identical
    except for the last bits of the softmax scores."""
    import csv
    import io
    import json
    from asme_b200 import evaluation as E
    from asme_b200.metrics import FusedPredictions
    with open(os.path.join(golden_dir, "predict_small.json")) as f:
        g = json.load(f)

    class Tok(_Tok):
        def __init__(self, tokens):
            self.vocabulary = _Vocab(0)
            self.vocabulary._tokens = list(tokens)

    tok = Tok(g["tokens"])
    n = g["num_predictions"]
    logits = torch.tensor(g["logits"])
    batch = {k: torch.tensor(v) for k, v in g["batch"].items()}
    val, idx = torch.topk(logits, n, dim=1)
    fused = FusedPredictions(None, idx.to(torch.int32), val, None, logits.shape[1], lse=torch.logsumexp(logits, dim=1))
    assert fused.shape[0] == logits.shape[0]

    def run(writer_cls, preds):
        evaluators = [E.ExtractSampleIdEvaluator(), E.LogInputEvaluator(tok), E.TrueTargetEvaluator(tok),
                      E.ExtractRecommendationEvaluator(tok, n), E.ExtractScoresEvaluator(tok, n)]
        out = io.StringIO()
        w = writer_cls(evaluators)
        w.init_file(out)
        w.write_evaluation(0, batch, preds)
        return out.getvalue()

    for name, cls in (("multi_line", E.CSVMultiLineWriter), ("single_line", E.CSVSingleLineWriter)):
        assert run(cls, logits) == g[name]
        got, want = list(csv.reader(io.StringIO(run(cls, fused)))), list(csv.reader(io.StringIO(g[name])))
        assert len(got) == len(want) and got[0] == want[0]
        for a, b in zip(got[1:], want[1:]):
            assert a[:-1] == b[:-1]
            fa, fb = np.asarray(json.loads(a[-1]), dtype=np.float64), np.asarray(json.loads(b[-1]), dtype=np.float64)
            np.testing.assert_allclose(fa, fb, rtol=1e-5)


def test_per_sample_metrics_evaluator_on_fused_ranks():
    """evaluation.py:239-287 on the fused input: per-sample metric values are closed forms of the target's exact rank"""
    from types import SimpleNamespace
    from asme_b200 import evaluation as E
    from asme_b200.metrics import FusedPredictions, build_metrics
    gen = torch.Generator().manual_seed(5)
    N, V = 7, 50
    module = SimpleNamespace(metrics=build_metrics({"recall": [1, 5], "ndcg": [5], "mrr": [5], "rank": []}))
    ev = E.PerSampleMetricsEvaluator(_Tok(V), None, module)
    assert ev.eval_samplewise() and ev.get_header() == module.metrics.get_metric_names()
    names = ev.get_header()
    for batch_index in range(2):                      # raw_metric_values()[batch_index]: one entry per evaluated batch
        logits = torch.randn(N, V, generator=gen)
        targets = torch.randint(3, V, (N,), generator=gen)
        rank = (logits > logits.gather(1, targets.unsqueeze(1))).sum(dim=1) + 1
        fused = FusedPredictions(rank.to(torch.int32), None, None, None, V)
        rows = np.asarray(ev.evaluate(batch_index, {"item.target": targets}, fused))
        assert rows.shape == (N, len(names))
        r = rank.numpy().astype(np.float64)
        want = {"recall@1": r <= 1, "recall@5": r <= 5, "NDCG@5": (r <= 5) / np.log2(r + 1), "MRR@5": (r <= 5) / r, "rank": r}
        for j, name in enumerate(names):
            np.testing.assert_allclose(rows[:, j], np.asarray(want[name], dtype=np.float64), rtol=1e-6, err_msg=name)
    with pytest.raises(RuntimeError, match="no target rank"):
        ev.evaluate(2, {}, FusedPredictions(None, torch.zeros(N, 3, dtype=torch.int32), torch.zeros(N, 3), None, V))
    with pytest.raises(RuntimeError, match="selected items"):
        E.PerSampleMetricsEvaluator(_Tok(V), [3, 4], module).evaluate(0, {}, fused)


def test_result_writers(tmp_path):
    """writer/results/results_writer.py: JSON object / CSV lines of the overall metrics, chosen by file extension"""
    import json
    from asme_b200 import evaluation as E
    metrics = {"recall@10": torch.tensor(0.25), "NDCG@10": torch.tensor(0.125), "MRR": 0.5}
    with open(tmp_path / "r.json", "w") as f:
        E.build_result_writer(f).write_overall_results("MaskedTrainingModule", metrics)
    assert json.loads((tmp_path / "r.json").read_text()) == {"recommender_id": "MaskedTrainingModule",
                                                            "metrics": {"recall@10": 0.25, "NDCG@10": 0.125, "MRR": 0.5}}
    with open(tmp_path / "r.csv", "w", newline="") as f:
        E.build_result_writer(f).write_overall_results("MaskedTrainingModule", metrics)
    assert (tmp_path / "r.csv").read_text().splitlines() == ["metric name,value", "recall@10,0.25", "NDCG@10,0.125", "MRR,0.5",
                                                             "recommender_id,MaskedTrainingModule"]
    assert E.check_file_format_supported("x.json") and not E.check_file_format_supported("x.txt")
    with open(tmp_path / "r.txt", "w") as f, pytest.raises(KeyError, match="not a supported format"):
        E.build_result_writer(f)


# ---- row a19: initialisers ---------------------------------------------------------------------------------------------------------
def _init_models():
    import json
    from asme_b200.models import UBERT4RecModel, UserSASRecModel
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "init_stats.json")) as f:
        ref = json.load(f)
    sh = ref["shape"]
    kw = dict(transformer_hidden_size=sh["H"], num_transformer_heads=sh["heads"], num_transformer_layers=sh["L"], item_vocab_size=sh["V"],
              max_seq_length=sh["S"], transformer_dropout=0.1)
    sizes = {"category": sh["VA"], "tags": sh["VT"], "user_id": sh["VU"]}
    pre = {"category": {"embedding_type": "content_embedding"}, "tags": {"embedding_type": "linear_upscale"}}
    ukw = dict(additional_attributes={"category": {"embedding_type": "content_embedding"}},
               user_attributes={"user_id": {"embedding_type": "user_embedding"}}, attribute_vocab_sizes=sizes, **kw)
    torch.manual_seed(11)
    ours = {
        "bert4rec": BERT4RecModel(**kw),
        "bert4rec_range_0.1": BERT4RecModel(initializer_range=0.1, **kw),
        "kebert4rec": KeBERT4RecModel(prefusion_attributes=pre, postfusion_attributes={"category": {"embedding_type": "content_embedding"}},
                                      attribute_vocab_sizes=sizes, **kw),
        "sasrec_full": SASRecModel(mode="full", **kw),
        "sasrec_neg": SASRecModel(mode="neg_sampling", **kw),
        "ubert4rec": UBERT4RecModel(segment_embedding=True, **ukw),
        "usasrec_full": UserSASRecModel(mode="full", **ukw),
    }
    return ref["models"], ours


@pytest.mark.parametrize("name", ["bert4rec", "bert4rec_range_0.1", "kebert4rec", "sasrec_full", "sasrec_neg", "ubert4rec", "usasrec_full"])
def test_initialisers_match_the_reference_statistics(name):
    """freshly constructed models: every parameter the reference creates exists with the same shape, constants (LayerNorm 1 / 0, zero
    biases) are exact, random tensors have the reference initialiser's distribution -- N(0, initializer_range) after
    normal_initialize_weights (bert4rec_model.py:59-68), xavier-normal for the SASRec family (transformer_encoder_model.py:63-73),
    U(+-1/sqrt(V)) for ``output_bias`` (layers.py:134-136).  The reference's statistics come from tests/golden/init_stats.json."""
    import math
    ref_all, ours_all = _init_models()
    ref, model = ref_all[name], ours_all[name]
    params = dict(model.named_parameters())
    sd = model.state_dict()
    for pname, st in ref.items():
        assert pname in sd, f"{name}: missing parameter {pname}"
        p = sd[pname].detach().double()
        assert list(p.shape) == st["shape"], pname
        n = p.numel()
        if st["std"] == 0.0:                                   # constants: LayerNorm weight 1, every bias 0
            assert float(p.min()) == float(p.max()) == st["mean"], pname
            continue
        std, mean = float(p.std(unbiased=False)), float(p.mean())
        tol = 4.0 / math.sqrt(2 * n) + 0.01                     # sampling error of a standard deviation (both sides) + slack
        assert abs(std / st["std"] - 1.0) < 2 * tol, f"{name}.{pname}: std {std} vs reference {st['std']}"
        assert abs(mean - st["mean"]) < 6 * st["std"] / math.sqrt(n) * 1.5, f"{name}.{pname}: mean {mean} vs {st['mean']}"
        if pname.endswith("output_bias"):                       # uniform, not normal: hard bounds
            bound = 1.0 / math.sqrt(n)
            assert float(p.min()) >= -bound and float(p.max()) <= bound
            assert abs(std - bound / math.sqrt(3)) < 0.05 * bound
        else:                                                   # normal: about 0.27 % beyond three standard deviations
            frac = float((p.abs() > 3 * st["std"]).double().mean())
            assert frac < 0.01 + 8.0 / n, f"{name}.{pname}: {frac} of the entries beyond 3 sigma"
    assert params, name
