"""Generate the committed golden fixtures by running the UNMODIFIED reference
(``/root/reference``) in the build container under the import shims of ``_shims/ref_shims.py``.

    python tests/golden/make_golden.py

Writes, next to this file:
  metric_vectors.json    every (predictions, positive_mask, k, expected) tuple of the reference's
                         own metric tests (tests/test_{recall,ndcg,dcg,mrr,precision,f1}.py)
  bert4rec_small.npz, kebert4rec_small.npz, sasrec_full_small.npz, sasrec_neg_small.npz,
  ubert4rec_small.npz, usasrec_full_small.npz
                         weights (reference state-dict names), inputs, and the reference's
                         outputs: logits / hidden stages, loss, per-parameter gradients,
                         eval rows, metric values.
  metrics_dense_small.npz  dense-signature metric values of the reference classes on random
                         (B,V) scores incl. ties, with the stable tie order of SURVEY.md 8c.

This script is the ONLY code that touches /root/reference; it cannot run on the GPU box.
"""
import importlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_shims"))
import ref_shims  # noqa: E402

ref_shims.install()


def to_np(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def grads_of(model, prefix="grad::"):
    return {prefix + n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}


def weights_of(model, prefix="w::"):
    return {prefix + n: v.detach().clone() for n, v in model.state_dict().items()}


def randomize(model, gen):
    """Replace the trivial LayerNorm(1,0) / zero-bias init by random values so that every
    parameter influences the fixture."""
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.2 + (1.0 if "norm" in n and n.endswith("weight") else 0.0))
            else:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.3)


def make_sequences(gen, b, s, v, min_len=1):
    seq = torch.randint(3, v, (b, s), generator=gen)
    lengths = torch.randint(min_len, s + 1, (b,), generator=gen)
    lengths[0] = s                      # at least one full row
    for i in range(b):
        seq[i, lengths[i]:] = 0
    return seq, lengths


def export_metric_vectors():
    sys.path.insert(0, "/root/reference/tests")
    out = {}
    for name in ["recall", "ndcg", "dcg", "mrr", "precision", "f1"]:
        mod = importlib.import_module(f"test_{name}")
        samples = []
        for fn in ("get_single_item_recommendation_samples", "get_multiple_item_recommendation_samples"):
            if hasattr(mod, fn):
                for pred, mask, k, val in getattr(mod, fn)():
                    samples.append({"predictions": pred.tolist(), "positive_mask": mask.tolist(), "k": int(k),
                                    "expected": float(val), "group": fn})
        out[name] = samples
    with open(os.path.join(HERE, "metric_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("metric vectors:", {k: len(v) for k, v in out.items()})


def cloze(seq, lengths, gen, p=0.3):
    inp = seq.clone()
    tgt = torch.zeros_like(seq)
    for i in range(seq.shape[0]):
        n = int(lengths[i])
        m = torch.rand(n, generator=gen) < p
        if not m.any():
            m[n - 1] = True
        tgt[i, :n][m] = seq[i, :n][m]
        inp[i, :n][m] = 1
    return inp, tgt


def bert4rec_fixture():
    from asme.core.models.bert4rec.bert4rec_model import BERT4RecModel
    from asme.core.models.common.layers.data.sequence import InputSequence
    gen = torch.Generator().manual_seed(101)
    V, S, H, L, heads, B = 61, 12, 16, 2, 2, 6
    ref_shims.set_injection_context({"item": ref_shims.make_tokenizer(V)})
    model = BERT4RecModel(transformer_hidden_size=H, num_transformer_heads=heads, num_transformer_layers=L,
                          max_seq_length=S, transformer_dropout=0.0)
    randomize(model, gen)
    seq, lengths = make_sequences(gen, B, S, V)
    inp, tgt = cloze(seq, lengths, gen)
    logits = model(InputSequence(inp, inp.ne(0), {}))
    loss = torch.nn.CrossEntropyLoss(ignore_index=0)(logits.view(-1, V), tgt.view(-1))
    loss.backward()
    # eval: append a MASK at the end of each (shortened) sequence, one MASK per row
    ev = seq.clone()
    ev_t = torch.randint(3, V, (B,), generator=gen)
    for i in range(B):
        n = min(int(lengths[i]), S - 1)
        ev[i, n] = 1
        ev[i, n + 1:] = 0
    with torch.no_grad():
        ev_logits = model(InputSequence(ev, ev.ne(0), {}))[ev.eq(1)]
    data = {"V": V, "S": S, "H": H, "L": L, "heads": heads, "input": inp, "target": tgt, "logits": logits,
            "loss": loss, "eval_input": ev, "eval_target": ev_t, "eval_logits": ev_logits}
    data.update(weights_of(model))
    data.update(grads_of(model))
    np.savez_compressed(os.path.join(HERE, "bert4rec_small.npz"), **to_np(data))
    print("bert4rec loss", float(loss))


def kebert4rec_fixture():
    from asme.core.models.kebert4rec.kebert4rec_model import KeBERT4RecModel
    from asme.core.models.common.layers.data.sequence import InputSequence
    gen = torch.Generator().manual_seed(202)
    V, S, H, L, heads, B, VA, VT, A = 83, 10, 16, 1, 2, 5, 13, 19, 3
    ref_shims.set_injection_context({"item": ref_shims.make_tokenizer(V), "category": ref_shims.make_tokenizer(VA, "c"),
                                     "tags": ref_shims.make_tokenizer(VT, "t")})
    model = KeBERT4RecModel(transformer_hidden_size=H, num_transformer_heads=heads, num_transformer_layers=L,
                            max_seq_length=S, transformer_dropout=0.0,
                            prefusion_attributes={"category": {"embedding_type": "content_embedding"},
                                                  "tags": {"embedding_type": "linear_upscale"}})
    randomize(model, gen)
    seq, lengths = make_sequences(gen, B, S, V)
    inp, tgt = cloze(seq, lengths, gen)
    cat = torch.randint(3, VA, (B, S), generator=gen)
    tags = torch.randint(0, VT, (B, S, A), generator=gen)       # includes pad id 0 and duplicates
    cat[inp == 0] = 0
    cat[inp == 1] = 1
    tags[inp == 0] = 0
    logits = model(InputSequence(inp, inp.ne(0), {"category": cat, "tags": tags}))
    loss = torch.nn.CrossEntropyLoss(ignore_index=0)(logits.view(-1, V), tgt.view(-1))
    loss.backward()
    data = {"V": V, "S": S, "H": H, "L": L, "heads": heads, "input": inp, "target": tgt, "category": cat,
            "tags": tags, "logits": logits, "loss": loss}
    data.update(weights_of(model))
    data.update(grads_of(model))
    np.savez_compressed(os.path.join(HERE, "kebert4rec_small.npz"), **to_np(data))
    print("kebert4rec loss", float(loss))


def sasrec_fixtures():
    from asme.core.models.sasrec.sasrec_model import SASRecModel
    from asme.core.models.common.layers.data.sequence import InputSequence
    from asme.core.losses.sasrec.sas_rec_losses import SASRecBinaryCrossEntropyLoss
    gen = torch.Generator().manual_seed(303)
    V, S, H, L, heads, B = 71, 9, 16, 2, 2, 7
    ref_shims.set_injection_context({"item": ref_shims.make_tokenizer(V)})
    # ---- mode="full" (sasrec-cross): per-position CE + last-position eval rows
    model = SASRecModel(transformer_hidden_size=H, num_transformer_heads=heads, num_transformer_layers=L,
                        max_seq_length=S, transformer_dropout=0.0, mode="full")
    randomize(model, gen)
    seq, lengths = make_sequences(gen, B, S, V)
    tgt = torch.zeros_like(seq)
    for i in range(B):
        n = int(lengths[i])
        tgt[i, :n] = torch.randint(3, V, (n,), generator=gen)
    logits = model(InputSequence(seq, seq.ne(0), {}))
    loss = torch.nn.CrossEntropyLoss(ignore_index=0)(logits.reshape(-1, V), tgt.reshape(-1))
    loss.backward()
    last = logits.detach()[torch.arange(B), seq.ne(0).sum(-1) - 1]
    data = {"V": V, "S": S, "H": H, "L": L, "heads": heads, "input": seq, "target": tgt, "logits": logits,
            "loss": loss, "eval_logits": last}
    data.update(weights_of(model))
    data.update(grads_of(model))
    np.savez_compressed(os.path.join(HERE, "sasrec_full_small.npz"), **to_np(data))
    print("sasrec full loss", float(loss))
    # ---- mode="neg_sampling" (sasrec-neg): pos/neg logits + BCE, and the all-items eval branch
    model = SASRecModel(transformer_hidden_size=H, num_transformer_heads=heads, num_transformer_layers=L,
                        max_seq_length=S, transformer_dropout=0.0, mode="neg_sampling")
    randomize(model, gen)
    pos = torch.zeros_like(seq)
    neg = torch.zeros_like(seq)
    for i in range(B):
        n = int(lengths[i])
        pos[i, :n] = torch.randint(3, V, (n,), generator=gen)
        neg[i, :n] = torch.randint(3, V, (n,), generator=gen)
    p, n_ = model(InputSequence(seq, seq.ne(0), {"positive_samples": pos, "negative_samples": neg}))
    loss = SASRecBinaryCrossEntropyLoss()(p, n_, mask=seq.ne(0))
    loss.backward()
    items = torch.arange(V).repeat(B, 1)
    with torch.no_grad():
        ev = model(InputSequence(seq, seq.ne(0), {"positive_samples": items}))
    data = {"V": V, "S": S, "H": H, "L": L, "heads": heads, "input": seq, "positive_samples": pos,
            "negative_samples": neg, "pos_logits": p, "neg_logits": n_, "loss": loss, "eval_logits": ev}
    data.update({k: v for k, v in weights_of(model).items() if "_projection_layer" not in k})
    data.update({k: v for k, v in grads_of(model).items() if "_projection_layer" not in k})
    np.savez_compressed(os.path.join(HERE, "sasrec_neg_small.npz"), **to_np(data))
    print("sasrec neg loss", float(loss))


def user_fixtures():
    """UBERT4Rec (cloze, user token prepended, segment embedding) and UserSASRec (mode="full") -- SURVEY.md 8f row 1."""
    from asme.core.models.ubert4rec.ubert4rec_model import UBERT4RecModel
    from asme.core.models.user_sasrec.user_sasrec_model import UserSASRecModel
    from asme.core.models.common.layers.data.sequence import InputSequence
    gen = torch.Generator().manual_seed(505)
    V, S, H, L, heads, B, VU, VG, VC = 67, 11, 16, 2, 2, 6, 9, 7, 13
    ref_shims.set_injection_context({"item": ref_shims.make_tokenizer(V), "user_id": ref_shims.make_tokenizer(VU, "u"),
                                     "gender": ref_shims.make_tokenizer(VG, "g"), "category": ref_shims.make_tokenizer(VC, "c")})
    toks = {"tokenizers.user_id": ref_shims.make_tokenizer(VU, "u"), "tokenizers.gender": ref_shims.make_tokenizer(VG, "g"),
            "tokenizers.category": ref_shims.make_tokenizer(VC, "c")}      # the factory resolves these annotations, not @inject
    user_attributes = {"user_id": {"embedding_type": "user_embedding"}, "gender": {"embedding_type": "content_embedding"}}
    additional = {"category": {"embedding_type": "content_embedding"}}
    seq, lengths = make_sequences(gen, B, S, V)
    # user attributes arrive as (B,S) sequence features of which only column 0 is read (ubert4rec/components.py:112-113)
    uid = torch.randint(3, VU, (B, 1), generator=gen).repeat(1, S)
    gender = torch.randint(3, VG, (B, 1), generator=gen).repeat(1, S)
    cat = torch.randint(3, VC, (B, S), generator=gen)
    # ---- UBERT4Rec
    model = UBERT4RecModel(transformer_hidden_size=H, num_transformer_heads=heads, num_transformer_layers=L,
                           item_vocab_size=V, additional_tokenizers=toks,
                           max_seq_length=S, transformer_dropout=0.0, additional_attributes=additional,
                           user_attributes=user_attributes, positional_embedding=True, segment_embedding=True)
    randomize(model, gen)
    inp, tgt = cloze(seq, lengths, gen)
    c = cat.clone()
    c[inp == 0] = 0
    c[inp == 1] = 1
    attrs = {"user_id": uid, "gender": gender, "category": c}
    logits = model(InputSequence(inp, inp.ne(0), attrs))                       # (B, S+1, V)
    tgt1 = torch.cat([torch.zeros(B, 1, dtype=tgt.dtype), tgt], dim=1)         # ubert_masked_training_module.py:75-77
    loss = torch.nn.CrossEntropyLoss(ignore_index=0)(logits.reshape(-1, V), tgt1.reshape(-1))
    loss.backward()
    ev = seq.clone()
    for i in range(B):
        n = min(int(lengths[i]), S - 1)
        ev[i, n] = 1
        ev[i, n + 1:] = 0
    ce = cat.clone()
    ce[ev == 0] = 0
    ce[ev == 1] = 1
    with torch.no_grad():
        full = model(InputSequence(ev, ev.ne(0), {"user_id": uid, "gender": gender, "category": ce}))
        mask1 = torch.cat([torch.zeros(B, 1, dtype=torch.bool), ev.eq(1)], dim=1)   # ubert_masked_training_module.py:93-97
        ev_logits = full[mask1]
    data = {"V": V, "S": S, "H": H, "L": L, "heads": heads, "input": inp, "target": tgt, "user_id": uid, "gender": gender,
            "category": c, "logits": logits, "loss": loss, "eval_input": ev, "eval_category": ce, "eval_logits": ev_logits}
    data.update(weights_of(model))
    data.update(grads_of(model))
    np.savez_compressed(os.path.join(HERE, "ubert4rec_small.npz"), **to_np(data))
    print("ubert4rec loss", float(loss))
    # ---- UserSASRec, mode="full"
    model = UserSASRecModel(transformer_hidden_size=H, num_transformer_heads=heads, num_transformer_layers=L,
                            item_vocab_size=V, additional_tokenizers=toks,
                            max_seq_length=S, transformer_dropout=0.0, additional_attributes=additional,
                            user_attributes=user_attributes, segment_embedding=False, mode="full")
    randomize(model, gen)
    tgt = torch.zeros_like(seq)
    for i in range(B):
        n = int(lengths[i])
        tgt[i, :n] = torch.randint(3, V, (n,), generator=gen)
    c = cat.clone()
    c[seq == 0] = 0
    attrs = {"user_id": uid, "gender": gender, "category": c}
    logits = model(InputSequence(seq, seq.ne(0), attrs))                        # (B, S+1, V)
    item_logits = logits[:, 1:, :]                                              # user_next_item_prediction_training_module.py:151-155
    loss = torch.nn.CrossEntropyLoss(ignore_index=0)(item_logits.reshape(-1, V), tgt.reshape(-1))
    loss.backward()
    # evaluation rows exactly as the reference module picks them: index (length - 1) of the S+1 positions (:124-135)
    last = logits.detach()[torch.arange(B), seq.ne(0).sum(-1) - 1]
    data = {"V": V, "S": S, "H": H, "L": L, "heads": heads, "input": seq, "target": tgt, "user_id": uid, "gender": gender,
            "category": c, "logits": logits, "loss": loss, "eval_logits": last}
    data.update(weights_of(model))
    data.update(grads_of(model))
    np.savez_compressed(os.path.join(HERE, "usasrec_full_small.npz"), **to_np(data))
    print("usasrec full loss", float(loss))


def user_fixtures_first_item():
    """UserSASRec whose user token REPLACES the first item (replace_first_item=True, the example configs'
    local_usersasrec_config_first_item.jsonnet), with a user_linear_upscale user attribute (a LIST of ids per user,
    models/ubert4rec/components.py:12-44), in both projection modes: "full" (Linear over the catalog, the module's first_item=True
    loss over all S positions) and "neg_sampling" (user_sasrec/components.py:10-60) -- SURVEY.md 8f row 1 leftovers."""
    from asme.core.models.user_sasrec.user_sasrec_model import UserSASRecModel
    from asme.core.models.common.layers.data.sequence import InputSequence
    from asme.core.losses.sasrec.sas_rec_losses import SASRecBinaryCrossEntropyLoss
    gen = torch.Generator().manual_seed(606)
    V, S, H, L, heads, B, VU, VG, VC, A = 59, 9, 16, 2, 2, 7, 11, 7, 13, 3
    toks = {"tokenizers.user_id": ref_shims.make_tokenizer(VU, "u"), "tokenizers.gender": ref_shims.make_tokenizer(VG, "g"),
            "tokenizers.category": ref_shims.make_tokenizer(VC, "c")}
    ref_shims.set_injection_context({"item": ref_shims.make_tokenizer(V), "user_id": toks["tokenizers.user_id"],
                                     "gender": toks["tokenizers.gender"], "category": toks["tokenizers.category"]})
    user_attributes = {"user_id": {"embedding_type": "user_linear_upscale"}, "gender": {"embedding_type": "content_embedding"}}
    additional = {"category": {"embedding_type": "content_embedding"}}
    seq, lengths = make_sequences(gen, B, S, V, min_len=2)
    # a list-valued user feature: (B, S, A) ids, 0-padded lists, the same list at every step; only step 0 is read
    uid = torch.randint(1, VU, (B, 1, A), generator=gen)
    uid[torch.rand(B, 1, A, generator=gen) < 0.3] = 0
    uid = uid.repeat(1, S, 1)
    gender = torch.randint(3, VG, (B, 1), generator=gen).repeat(1, S)
    cat = torch.randint(3, VC, (B, S), generator=gen)
    cat[seq == 0] = 0
    attrs = {"user_id": uid, "gender": gender, "category": cat}
    tgt = torch.zeros_like(seq)
    for i in range(B):
        n = int(lengths[i])
        tgt[i, :n] = torch.randint(3, V, (n,), generator=gen)
    common = dict(transformer_hidden_size=H, num_transformer_heads=heads, num_transformer_layers=L, item_vocab_size=V,
                  additional_tokenizers=toks, max_seq_length=S, transformer_dropout=0.0, additional_attributes=additional,
                  user_attributes=user_attributes, segment_embedding=False, replace_first_item=True)
    base = {"V": V, "S": S, "H": H, "L": L, "heads": heads, "input": seq, "target": tgt, "user_id": uid, "gender": gender, "category": cat}
    # ---- mode="full": first_item=True keeps all S logit rows (user_next_item_prediction_training_module.py:57-60)
    model = UserSASRecModel(mode="full", **common)
    randomize(model, gen)
    logits = model(InputSequence(seq, seq.ne(0), attrs))                        # (B, S, V)
    assert tuple(logits.shape) == (B, S, V)
    loss = torch.nn.CrossEntropyLoss(ignore_index=0)(logits.reshape(-1, V), tgt.reshape(-1))
    loss.backward()
    last = logits.detach()[torch.arange(B), seq.ne(0).sum(-1) - 1]             # :124-135
    data = dict(base, logits=logits, loss=loss, eval_logits=last)
    data.update(weights_of(model))
    data.update(grads_of(model))
    np.savez_compressed(os.path.join(HERE, "usasrec_first_item_full.npz"), **to_np(data))
    print("usasrec first-item full loss", float(loss))
    # ---- mode="neg_sampling"
    model = UserSASRecModel(mode="neg_sampling", **common)
    randomize(model, gen)
    pos = tgt.clone()
    neg = torch.randint(3, V, (B, S), generator=gen)
    neg[seq == 0] = 0
    pl, nl = model(InputSequence(seq, seq.ne(0), dict(attrs, positive_samples=pos, negative_samples=neg)))
    loss = SASRecBinaryCrossEntropyLoss()(pl, nl, seq.ne(0))
    loss.backward()
    items = torch.randint(0, V, (B, 12), generator=gen)
    with torch.no_grad():
        ev = model(InputSequence(seq, seq.ne(0), dict(attrs, positive_samples=items)))         # (B, 12)
    data = dict(base, positive=pos, negative=neg, pos_logits=pl, neg_logits=nl, loss=loss, eval_items=items, eval_logits=ev)
    data.update(weights_of(model))
    data.update(grads_of(model))
    np.savez_compressed(os.path.join(HERE, "usasrec_first_item_neg.npz"), **to_np(data))
    print("usasrec first-item neg loss", float(loss))


def basket_fixtures():
    """basket inputs (N,S,BS) pooled per step (models/common/layers/sequence_embedding.py:9-45, :83-93): KeBERT4Rec with max / sum /
    mean pooling and one attribute table, SASRec (mode="full") with mean pooling -- SURVEY.md 2.1 #3 (dense fallback)."""
    from asme.core.models.kebert4rec.kebert4rec_model import KeBERT4RecModel
    from asme.core.models.sasrec.sasrec_model import SASRecModel
    from asme.core.models.common.layers.data.sequence import InputSequence
    gen = torch.Generator().manual_seed(707)
    V, S, H, L, heads, B, VA, BS = 71, 8, 16, 1, 2, 6, 11, 3
    ref_shims.set_injection_context({"item": ref_shims.make_tokenizer(V), "category": ref_shims.make_tokenizer(VA, "c")})
    seq, lengths = make_sequences(gen, B, S, V)
    basket = torch.zeros(B, S, BS, dtype=torch.int64)
    basket[:, :, 0] = seq
    for j in range(1, BS):               # further items of a step, 0-padded baskets of different sizes
        extra = torch.randint(3, V, (B, S), generator=gen)
        extra[torch.rand(B, S, generator=gen) < 0.4] = 0
        extra[seq == 0] = 0
        basket[:, :, j] = extra
    pm = basket.max(dim=2).values.ne(0)                                  # modules/util/module_util.py:25-29
    cat = torch.randint(3, VA, (B, S), generator=gen)
    cat[seq == 0] = 0
    tgt = torch.zeros_like(seq)
    for i in range(B):
        n = int(lengths[i])
        tgt[i, :n] = torch.randint(3, V, (n,), generator=gen)
    for pooling in ("max", "sum", "mean"):
        model = KeBERT4RecModel(transformer_hidden_size=H, num_transformer_heads=heads, num_transformer_layers=L, max_seq_length=S,
                                transformer_dropout=0.0, embedding_pooling_type=pooling,
                                prefusion_attributes={"category": {"embedding_type": "content_embedding"}})
        randomize(model, gen)
        logits = model(InputSequence(basket, pm, {"category": cat}))
        loss = torch.nn.CrossEntropyLoss(ignore_index=0)(logits.view(-1, V), tgt.view(-1))
        loss.backward()
        data = {"V": V, "S": S, "H": H, "L": L, "heads": heads, "input": basket, "target": tgt, "category": cat, "logits": logits, "loss": loss}
        data.update(weights_of(model))
        data.update(grads_of(model))
        np.savez_compressed(os.path.join(HERE, f"kebert4rec_basket_{pooling}.npz"), **to_np(data))
        print("kebert4rec basket", pooling, float(loss))
    model = SASRecModel(transformer_hidden_size=H, num_transformer_heads=heads, num_transformer_layers=L, max_seq_length=S,
                        transformer_dropout=0.0, embedding_pooling_type="mean", mode="full")
    randomize(model, gen)
    logits = model(InputSequence(basket, pm, {}))
    loss = torch.nn.CrossEntropyLoss(ignore_index=0)(logits.view(-1, V), tgt.view(-1))
    loss.backward()
    data = {"V": V, "S": S, "H": H, "L": L, "heads": heads, "input": basket, "target": tgt, "logits": logits, "loss": loss}
    data.update(weights_of(model))
    data.update(grads_of(model))
    np.savez_compressed(os.path.join(HERE, "sasrec_basket_mean.npz"), **to_np(data))
    print("sasrec basket mean", float(loss))


def postfusion_fixtures():
    """post-fusion attributes: the attribute embeddings are merged into the ENCODED sequence, by ``add`` or
    ``multiply``, before the modifier transform (KeBERT4Rec, models/kebert4rec/components.py:97-116) or instead of it (SASRec,
    models/sasrec/components.py:88-106)."""
    from asme.core.models.kebert4rec.kebert4rec_model import KeBERT4RecModel
    from asme.core.models.sasrec.sasrec_model import SASRecModel
    from asme.core.models.common.layers.data.sequence import InputSequence
    V, S, H, L, heads, B, VA, VT, A = 59, 8, 16, 1, 2, 5, 11, 17, 3
    ref_shims.set_injection_context({"item": ref_shims.make_tokenizer(V), "category": ref_shims.make_tokenizer(VA, "c"),
                                     "tags": ref_shims.make_tokenizer(VT, "t")})
    for merge in ("add", "multiply"):
        gen = torch.Generator().manual_seed(606 + len(merge))
        # ---- KeBERT4Rec: one pre-fused and two post-fused attributes (a table and an id-bag Linear)
        model = KeBERT4RecModel(transformer_hidden_size=H, num_transformer_heads=heads, num_transformer_layers=L,
                                max_seq_length=S, transformer_dropout=0.0,
                                prefusion_attributes={"category": {"embedding_type": "content_embedding"}},
                                postfusion_attributes={"category": {"embedding_type": "content_embedding"},
                                                       "tags": {"embedding_type": "linear_upscale"}},
                                postfusion_merge_function=merge)
        randomize(model, gen)
        seq, lengths = make_sequences(gen, B, S, V)
        inp, tgt = cloze(seq, lengths, gen)
        cat = torch.randint(3, VA, (B, S), generator=gen)
        tags = torch.randint(0, VT, (B, S, A), generator=gen)
        cat[inp == 0] = 0
        cat[inp == 1] = 1
        tags[inp == 0] = 0
        logits = model(InputSequence(inp, inp.ne(0), {"category": cat, "tags": tags}))
        loss = torch.nn.CrossEntropyLoss(ignore_index=0)(logits.view(-1, V), tgt.view(-1))
        loss.backward()
        data = {"V": V, "S": S, "H": H, "L": L, "heads": heads, "input": inp, "target": tgt, "category": cat, "tags": tags,
                "logits": logits, "loss": loss}
        data.update(weights_of(model))
        data.update(grads_of(model))
        np.savez_compressed(os.path.join(HERE, f"kebert4rec_postfusion_{merge}.npz"), **to_np(data))
        print(f"kebert4rec postfusion {merge} loss", float(loss))
        # ---- SASRec (mode="full"): identity modifier with one post-fused table
        model = SASRecModel(transformer_hidden_size=H, num_transformer_heads=heads, num_transformer_layers=L,
                            max_seq_length=S, transformer_dropout=0.0, mode="full",
                            postfusion_attributes={"category": {"embedding_type": "content_embedding"}},
                            postfusion_merge_function=merge)
        randomize(model, gen)
        tgt = torch.zeros_like(seq)
        for i in range(B):
            n = int(lengths[i])
            tgt[i, :n] = torch.randint(3, V, (n,), generator=gen)
        c = torch.randint(3, VA, (B, S), generator=gen)
        c[seq == 0] = 0
        logits = model(InputSequence(seq, seq.ne(0), {"category": c}))
        loss = torch.nn.CrossEntropyLoss(ignore_index=0)(logits.reshape(-1, V), tgt.reshape(-1))
        loss.backward()
        data = {"V": V, "S": S, "H": H, "L": L, "heads": heads, "input": seq, "target": tgt, "category": c, "logits": logits, "loss": loss}
        data.update(weights_of(model))
        data.update(grads_of(model))
        np.savez_compressed(os.path.join(HERE, f"sasrec_postfusion_{merge}.npz"), **to_np(data))
        print(f"sasrec postfusion {merge} loss", float(loss))


def init_stats_fixture():
    """row a19: per-parameter statistics of the reference's initialisers (models/bert4rec/bert4rec_model.py:59-68 N(0, range) applied
    AFTER models/transformer/transformer_encoder_model.py:63-73 xavier-normal; layers.py:134-136 U(+-1/sqrt(V)) for ``output_bias``)
    for every model family at a shape large enough for a statistical comparison."""
    from asme.core.models.bert4rec.bert4rec_model import BERT4RecModel
    from asme.core.models.kebert4rec.kebert4rec_model import KeBERT4RecModel
    from asme.core.models.sasrec.sasrec_model import SASRecModel
    from asme.core.models.ubert4rec.ubert4rec_model import UBERT4RecModel
    from asme.core.models.user_sasrec.user_sasrec_model import UserSASRecModel
    V, S, H, L, heads, VA, VT, VU = 2003, 32, 64, 1, 2, 301, 157, 211
    ref_shims.set_injection_context({"item": ref_shims.make_tokenizer(V), "category": ref_shims.make_tokenizer(VA, "c"),
                                     "tags": ref_shims.make_tokenizer(VT, "t"), "user_id": ref_shims.make_tokenizer(VU, "u")})
    toks = {"tokenizers.user_id": ref_shims.make_tokenizer(VU, "u"), "tokenizers.category": ref_shims.make_tokenizer(VA, "c")}
    kw = dict(transformer_hidden_size=H, num_transformer_heads=heads, num_transformer_layers=L, max_seq_length=S, transformer_dropout=0.1)
    pre = {"category": {"embedding_type": "content_embedding"}, "tags": {"embedding_type": "linear_upscale"}}
    ukw = dict(item_vocab_size=V, additional_tokenizers=toks, additional_attributes={"category": {"embedding_type": "content_embedding"}},
               user_attributes={"user_id": {"embedding_type": "user_embedding"}}, **kw)
    torch.manual_seed(7)
    models = {
        "bert4rec": BERT4RecModel(**kw),
        "bert4rec_range_0.1": BERT4RecModel(initializer_range=0.1, **kw),
        "kebert4rec": KeBERT4RecModel(prefusion_attributes=pre, postfusion_attributes={"category": {"embedding_type": "content_embedding"}}, **kw),
        "sasrec_full": SASRecModel(mode="full", **kw),
        "sasrec_neg": SASRecModel(mode="neg_sampling", **kw),
        "ubert4rec": UBERT4RecModel(positional_embedding=True, segment_embedding=True, **ukw),
        "usasrec_full": UserSASRecModel(segment_embedding=False, mode="full", **ukw),
    }
    out = {"shape": dict(V=V, S=S, H=H, L=L, heads=heads, VA=VA, VT=VT, VU=VU), "models": {}}
    for name, model in models.items():
        stats = {}
        for pname, p in model.named_parameters():
            x = p.detach().double()
            stats[pname] = {"shape": list(p.shape), "mean": float(x.mean()), "std": float(x.std(unbiased=False)), "min": float(x.min()),
                            "max": float(x.max())}
        out["models"][name] = stats
    with open(os.path.join(HERE, "init_stats.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("init stats:", {k: len(v) for k, v in out["models"].items()})


def metrics_fixture():
    """Reference metric classes on random scores; ties are made deterministic by patching
    torch.argsort to a stable sort inside the reference call (the reference's own argsort is
    unstable, SURVEY.md 8c fixes 'ties -> lowest item id')."""
    import asme.core.metrics.common as common
    from asme.core.metrics.container.metrics_container import RankingMetricsContainer
    from asme.core.metrics.container.metrics_sampler import AllItemsSampler
    from asme.core.metrics.recall import RecallMetric
    from asme.core.metrics.ndcg import NormalizedDiscountedCumulativeGainMetric
    from asme.core.metrics.mrr import MRRMetric
    from asme.core.metrics.precision import PrecisionMetric
    from asme.core.metrics.f1 import F1Metric
    from asme.core.metrics.dcg import DiscountedCumulativeGainMetric
    from asme.core.metrics.mrr_full import MRRFullMetric
    from asme.core.metrics.rank import Rank

    class _StableTorch:
        def __getattr__(self, name):
            return getattr(torch, name)

        @staticmethod
        def argsort(x, descending=False):
            return torch.sort(x, descending=descending, stable=True).indices

    common.torch = _StableTorch()
    gen = torch.Generator().manual_seed(404)
    B, V = 32, 57
    pred = torch.round(torch.randn(B, V, generator=gen) * 4) / 4          # coarse grid -> many ties
    targets = torch.randint(0, V, (B,), generator=gen)
    ks = [1, 3, 5, 10]
    metrics = []
    for k in ks:
        metrics += [RecallMetric(k), NormalizedDiscountedCumulativeGainMetric(k), MRRMetric(k), PrecisionMetric(k),
                    F1Metric(k), DiscountedCumulativeGainMetric(k)]
    metrics += [MRRFullMetric(), Rank()]
    container = RankingMetricsContainer(metrics, AllItemsSampler())
    step = container.update(None, targets, pred)
    final = container.compute()
    data = {"predictions": pred, "targets": targets, "ks": np.array(ks)}
    for name, v in step.items():
        data["step::" + name] = v
    for name, v in final.items():
        data["final::" + name] = v
    np.savez_compressed(os.path.join(HERE, "metrics_dense_small.npz"), **to_np(data))
    common.torch = torch
    print("metrics:", {k: float(v) for k, v in final.items()})


if __name__ == "__main__":
    torch.manual_seed(0)
    only = sys.argv[1:]
    steps = {"metric_vectors": export_metric_vectors, "bert4rec": bert4rec_fixture, "kebert4rec": kebert4rec_fixture,
             "sasrec": sasrec_fixtures, "user": user_fixtures, "user_first_item": user_fixtures_first_item, "basket": basket_fixtures, "postfusion": postfusion_fixtures, "init": init_stats_fixture, "metrics": metrics_fixture}
    for name, fn in steps.items():
        if not only or name in only:
            fn()
