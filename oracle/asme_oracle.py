"""CPU oracle for the ASME sequential-recommender hot path.

TEST INFRASTRUCTURE ONLY. This file is a functional restatement (plain torch CPU ops for the
floating-point stages, numpy for the integer / index stages) of the reference algorithm
``LSX-UniWue/recsys-22-user-attributes-recommender`` for the path SURVEY.md section 8 names.  It
is imported only by ``tests/``, by ``__graft_entry__.smoke()`` and by ``bench.py``'s CPU
baseline / ``--impl reference`` arm -- never by the product package ``asme_b200``.

Pinning:
  * metrics: every golden vector of the reference's own tests (``tests/test_{recall,ndcg,dcg,
    mrr,precision,f1}.py``), exported to ``tests/golden/metric_vectors.json``;
  * everything else ("parity unpinned" by reference tests): outputs of the UNMODIFIED reference
    classes, run in the build container under import shims by ``tests/golden/make_golden.py``
    and committed as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` compares.

All citations are relative to ``/root/reference/src/asme/core``. Weights are addressed by the
reference's state-dict key names, so a reference checkpoint feeds the oracle
unchanged.
"""
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

PAD_ID = 0   # tests/example_dataset/example.vocabulary.item_id.txt: <PAD>=0, <MASK>=1, <UNK>=2
MASK_ID = 1
UNK_ID = 2
ATTENTION_FILL = -1e9            # models/common/layers/transformer_layers.py:148 (quirk Q4)
LN_EPS = 1e-5                    # torch.nn.LayerNorm default used everywhere in the reference
BCE_EPS = 1e-24                  # losses/sasrec/sas_rec_losses.py:48

Weights = Dict[str, torch.Tensor]

# Training-mode dropout probability of every nn.Dropout site of the reference path (transformer_layers.py:40,78,
# 116,130,152,171,215,249,258; kebert4rec/components.py:51,62).  0 = eval mode / parity mode.  Only the CPU baseline
# timing sets it (the masks are torch's, parity tests always run with 0).
DROPOUT_P = 0.0


def _drop(x: torch.Tensor) -> torch.Tensor:
    return F.dropout(x, DROPOUT_P, training=True) if DROPOUT_P > 0.0 else x


# --------------------------------------------------------------------------------------------
# a1  padding mask                                  modules/util/module_util.py:13-30
# --------------------------------------------------------------------------------------------
def padding_mask(sequence: torch.Tensor, pad_id: int = PAD_ID) -> torch.Tensor:
    if sequence.dim() > 2:                       # basket case reduces over the basket first
        sequence = sequence.max(dim=2).values
    return sequence.ne(pad_id)


# --------------------------------------------------------------------------------------------
# a2/a3  item (+position) embedding, LayerNorm       models/common/layers/sequence_embedding.py:83-93
#                                                    models/common/layers/transformer_layers.py:55-80
# --------------------------------------------------------------------------------------------
def layer_norm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)     # biased variance, as nn.LayerNorm
    return (x - mean) / torch.sqrt(var + LN_EPS) * gamma + beta


def transformer_embedding(seq: torch.Tensor, item_table: torch.Tensor,
                          position_table: Optional[torch.Tensor],
                          norm: Optional[Tuple[torch.Tensor, torch.Tensor]]) -> torch.Tensor:
    """E[ids] (+ P[arange(S)]) -> LayerNorm?   (dropout is identity in eval / p=0).

    Positions count from 0 at the LEFT of the padded tensor (quirk Q3); PAD and MASK rows are
    ordinary table rows (quirk Q9)."""
    x = item_table[seq]
    if position_table is not None:
        s = seq.shape[1]
        x = x + position_table[torch.arange(s)].unsqueeze(0)
    if norm is not None:
        x = _drop(layer_norm(x, norm[0], norm[1]))
    return x


# --------------------------------------------------------------------------------------------
# a4  attribute pre-fusion                           models/kebert4rec/components.py:54-63
#     LinearUpscaler = embedding-bag                 models/kebert4rec/layers.py:15-27
# --------------------------------------------------------------------------------------------
def linear_upscale(ids: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """multi_hot(ids).float() @ W.T + b with column 0 zeroed; duplicates are COUNTED and the bias
    is always added.  ids: (B,S,A); weight: (H,Va) (nn.Linear layout); bias: (H)."""
    va = weight.shape[1]
    multi_hot = F.one_hot(ids, va).sum(2).to(weight.dtype)
    multi_hot[:, :, 0] = 0
    return multi_hot @ weight.t() + bias


def attribute_sum(attrs: Dict[str, torch.Tensor], w: Weights, prefix: str,
                  names: Sequence[str]) -> Optional[torch.Tensor]:
    """Sum over attribute modules ``<prefix>.<name>`` in ModuleDict order."""
    total = None
    for name in names:
        if f"{prefix}.{name}.weight" in w:                       # content_embedding
            part = w[f"{prefix}.{name}.weight"][attrs[name]]
        else:                                                    # linear_upscale
            part = linear_upscale(attrs[name], w[f"{prefix}.{name}.linear.weight"],
                                  w[f"{prefix}.{name}.linear.bias"])
        total = part if total is None else total + part
    return total


# --------------------------------------------------------------------------------------------
# a5  attention mask                                 models/transformer/sequence_representation.py:17-51
# --------------------------------------------------------------------------------------------
def attention_mask(pad_mask: Optional[torch.Tensor], batch: int, seq_len: int,
                   bidirectional: bool) -> Optional[torch.Tensor]:
    """(B,1,S,S) float/bool mask, 0 = masked; None when bidirectional without padding mask."""
    if bidirectional:
        if pad_mask is None:
            return None
        return pad_mask.unsqueeze(1).repeat(1, seq_len, 1).unsqueeze(1)
    causal = torch.tril(torch.ones(seq_len, seq_len)).unsqueeze(0).repeat(batch, 1, 1).unsqueeze(1)
    if pad_mask is not None:
        causal = causal * pad_mask.unsqueeze(1).repeat(1, seq_len, 1).unsqueeze(1)
    return causal


# --------------------------------------------------------------------------------------------
# a6-a8  transformer block                           models/common/layers/transformer_layers.py:120-258
# --------------------------------------------------------------------------------------------
def gelu_erf(x: torch.Tensor) -> torch.Tensor:
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))          # nn.GELU() = exact erf form


def multi_head_attention(x: torch.Tensor, w: Weights, prefix: str, heads: int,
                         mask: Optional[torch.Tensor]) -> torch.Tensor:
    b, s, h = x.shape
    d = h // heads
    q, k, v = [(x @ w[f"{prefix}.linear_layers.{i}.weight"].t() + w[f"{prefix}.linear_layers.{i}.bias"])
               .view(b, s, heads, d).transpose(1, 2) for i in range(3)]
    scores = (q @ k.transpose(-2, -1)) / math.sqrt(d)
    if mask is not None:
        scores = scores.masked_fill(mask == 0, ATTENTION_FILL)       # -1e9, not -inf (Q4)
    p = _drop(torch.softmax(scores, dim=-1))
    ctx = (p @ v).transpose(1, 2).contiguous().view(b, s, h)
    return ctx @ w[f"{prefix}.output_linear.weight"].t() + w[f"{prefix}.output_linear.bias"]


def feed_forward(x: torch.Tensor, w: Weights, prefix: str) -> torch.Tensor:
    inner = _drop(gelu_erf(x @ w[f"{prefix}.w_1.weight"].t() + w[f"{prefix}.w_1.bias"]))
    return inner @ w[f"{prefix}.w_2.weight"].t() + w[f"{prefix}.w_2.bias"]


def transformer_block(x: torch.Tensor, w: Weights, prefix: str, heads: int,
                      mask: Optional[torch.Tensor]) -> torch.Tensor:
    """pre-LN residual sublayers, NO final LayerNorm (transformer_layers.py:251-258)."""
    y = layer_norm(x, w[f"{prefix}.input_sublayer.norm.weight"], w[f"{prefix}.input_sublayer.norm.bias"])
    x = x + _drop(multi_head_attention(y, w, f"{prefix}.attention", heads, mask))
    y = layer_norm(x, w[f"{prefix}.output_sublayer.norm.weight"], w[f"{prefix}.output_sublayer.norm.bias"])
    x = x + _drop(feed_forward(y, w, f"{prefix}.feed_forward"))
    return _drop(x)


def transformer_encoder(x: torch.Tensor, w: Weights, heads: int, layers: int,
                        mask: Optional[torch.Tensor],
                        prefix: str = "_sequence_representation_layer.transformer_layer") -> torch.Tensor:
    for l in range(layers):
        x = transformer_block(x, w, f"{prefix}.transformer_blocks.{l}", heads, mask)
    return x


# --------------------------------------------------------------------------------------------
# a9  modifiers                                      models/common/components/representation_modifier/ffn_modifier.py:24-26
# --------------------------------------------------------------------------------------------
def ffn_modifier(x: torch.Tensor, w: Weights, prefix: str = "_sequence_representation_modifier_layer") -> torch.Tensor:
    y = gelu_erf(x @ w[f"{prefix}.transform.0.weight"].t() + w[f"{prefix}.transform.0.bias"])
    return layer_norm(y, w[f"{prefix}.transform.2.weight"], w[f"{prefix}.transform.2.bias"])


def postfusion_merge(x: torch.Tensor, context: torch.Tensor, merge: str) -> torch.Tensor:
    """models/kebert4rec/components.py:100-113, models/sasrec/components.py:88-106."""
    if merge == "add":
        return x + context
    if merge == "multiply":
        return x * context
    return x


# --------------------------------------------------------------------------------------------
# a10/a11  projections                               models/common/layers/layers.py:105-143
#                                                    models/sasrec/components.py:21-61
# --------------------------------------------------------------------------------------------
def project(h: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """h @ W.T + b, W = tied item table (BERT4Rec) or the untied Linear weight."""
    return h @ weight.t() + bias


def sasrec_pos_neg(h: torch.Tensor, item_table: torch.Tensor, pos: torch.Tensor,
                   neg: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    return (item_table[pos] * h).sum(-1), (item_table[neg] * h).sum(-1)


def sasrec_score_items(h: torch.Tensor, pad_mask: torch.Tensor, item_table: torch.Tensor,
                       items: torch.Tensor) -> torch.Tensor:
    """eval branch: E[items] (B,I,H) @ h[b, len_b-1]  (no bias)."""
    b = h.shape[0]
    last = h[torch.arange(b), pad_mask.sum(-1) - 1]
    return (item_table[items] @ last.unsqueeze(-1)).squeeze(-1)


# --------------------------------------------------------------------------------------------
# a12/a13  losses                                    modules/masked_training_module.py:93-111
#                                                    losses/sasrec/sas_rec_losses.py:15-75
# --------------------------------------------------------------------------------------------
def cross_entropy_ignore_pad(logits: torch.Tensor, target: torch.Tensor, pad_id: int = PAD_ID) -> torch.Tensor:
    """mean over rows with target != pad of (logsumexp(row) - row[target]);  NaN if no such row."""
    v = logits.shape[-1]
    logits = logits.reshape(-1, v)
    target = target.reshape(-1)
    keep = target.ne(pad_id)
    lse = torch.logsumexp(logits[keep], dim=-1)
    picked = logits[keep].gather(1, target[keep].unsqueeze(1)).squeeze(1)
    return (lse - picked).sum() / keep.sum()


def sasrec_bce(pos_logits: torch.Tensor, neg_logits: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    pos = torch.log(torch.sigmoid(pos_logits) + BCE_EPS) * mask
    neg = torch.log(1 - torch.sigmoid(neg_logits) + BCE_EPS) * mask      # 1-sigmoid(n), not sigmoid(-n)
    return torch.sum(-pos - neg) / torch.sum(mask)


# --------------------------------------------------------------------------------------------
# a14  row selection                                 modules/masked_training_module.py:80-91
#                                                    modules/next_item_prediction_training_module.py:226-244
# --------------------------------------------------------------------------------------------
def select_masked_rows(x: torch.Tensor, seq: torch.Tensor, mask_id: int = MASK_ID) -> torch.Tensor:
    return x[seq.eq(mask_id)]


def select_last_rows(x: torch.Tensor, seq: torch.Tensor, pad_id: int = PAD_ID) -> torch.Tensor:
    lengths = padding_mask(seq, pad_id).sum(dim=-1) - 1
    return x[torch.arange(seq.shape[0]), lengths]


# --------------------------------------------------------------------------------------------
# whole models (embed -> encode -> modify -> project), models/sequence_recommendation_model.py:35-53
# --------------------------------------------------------------------------------------------
_ITEM = "_sequence_embedding_layer.item_embedding.embedding.weight"
_PRE = "_sequence_embedding_layer.item_embedding_layer"
_ATTR = "_sequence_embedding_layer.prefusion_attribute_embeddings"


def bert4rec_hidden(w: Weights, seq: torch.Tensor, heads: int, layers: int) -> torch.Tensor:
    """models/bert4rec/bert4rec_model.py:24-56. No positional embedding (quirk Q1)."""
    pm = padding_mask(seq)
    x = transformer_embedding(seq, w[_ITEM], None,
                              (w["_sequence_embedding_layer.embedding_norm.weight"],
                               w["_sequence_embedding_layer.embedding_norm.bias"]))
    x = transformer_encoder(x, w, heads, layers, attention_mask(pm, seq.shape[0], seq.shape[1], True))
    return ffn_modifier(x, w)


def bert4rec_logits(w: Weights, seq: torch.Tensor, heads: int, layers: int) -> torch.Tensor:
    h = bert4rec_hidden(w, seq, heads, layers)
    if "_projection_layer.output_bias" in w:                        # transpose_embedding (default)
        return project(h, w[_ITEM], w["_projection_layer.output_bias"])
    return project(h, w["_projection_layer.linear.weight"], w["_projection_layer.linear.bias"])


def kebert4rec_hidden(w: Weights, seq: torch.Tensor, attrs: Dict[str, torch.Tensor], heads: int,
                      layers: int, prefusion: Sequence[str] = (), postfusion: Sequence[str] = (),
                      merge: str = "add") -> torch.Tensor:
    """models/kebert4rec/kebert4rec_model.py:24-84."""
    pm = padding_mask(seq)
    pos = w.get(f"{_PRE}.position_embedding.weight")
    x = transformer_embedding(seq, w[f"{_PRE}.item_embedding.embedding.weight"], pos, None)
    ctx = attribute_sum(attrs, w, _ATTR, prefusion)
    if ctx is not None:
        x = x + ctx
    x = _drop(layer_norm(x, w["_sequence_embedding_layer.norm_embedding.weight"],
                         w["_sequence_embedding_layer.norm_embedding.bias"]))
    x = transformer_encoder(x, w, heads, layers, attention_mask(pm, seq.shape[0], seq.shape[1], True))
    if postfusion:
        ctx = attribute_sum(attrs, w, "_sequence_representation_modifier_layer.postfusion_attribute_embeddings",
                            postfusion)
        x = postfusion_merge(x, ctx, merge)
    return ffn_modifier(x, w)


def kebert4rec_logits(w: Weights, seq, attrs, heads, layers, prefusion=(), postfusion=(), merge="add"):
    h = kebert4rec_hidden(w, seq, attrs, heads, layers, prefusion, postfusion, merge)
    return project(h, w["_projection_layer.linear.weight"], w["_projection_layer.linear.bias"])


def sasrec_hidden(w: Weights, seq: torch.Tensor, attrs: Dict[str, torch.Tensor], heads: int, layers: int,
                  prefusion: Sequence[str] = (), postfusion: Sequence[str] = (), merge: str = "add") -> torch.Tensor:
    """models/sasrec/sasrec_model.py:28-95: causal; LayerNorm applied TWICE to the embedding (Q2)."""
    pm = padding_mask(seq)
    x = transformer_embedding(seq, w[f"{_PRE}.item_embedding.embedding.weight"],
                              w[f"{_PRE}.position_embedding.weight"],
                              (w[f"{_PRE}.embedding_norm.weight"], w[f"{_PRE}.embedding_norm.bias"]))
    ctx = attribute_sum(attrs, w, _ATTR, prefusion)
    if ctx is not None:
        x = x + ctx
    x = _drop(layer_norm(x, w["_sequence_embedding_layer.norm_embedding.weight"],
                         w["_sequence_embedding_layer.norm_embedding.bias"]))
    x = transformer_encoder(x, w, heads, layers, attention_mask(pm, seq.shape[0], seq.shape[1], False))
    if postfusion:
        ctx = attribute_sum(attrs, w, "_sequence_representation_modifier_layer.postfusion_attribute_embeddings",
                            postfusion)
        x = postfusion_merge(x, ctx, merge)
    return x


def sasrec_full_logits(w: Weights, seq, attrs, heads, layers, **kw) -> torch.Tensor:
    h = sasrec_hidden(w, seq, attrs, heads, layers, **kw)
    return project(h, w["_projection_layer.linear.weight"], w["_projection_layer.linear.bias"])


def sasrec_neg_logits(w: Weights, seq, pos, neg, heads, layers, **kw):
    h = sasrec_hidden(w, seq, {}, heads, layers, **kw)
    return sasrec_pos_neg(h, w[f"{_PRE}.item_embedding.embedding.weight"], pos, neg)


# --------------------------------------------------------------------------------------------
# 8f row 1  user-attribute models                    models/ubert4rec/components.py:96-133, :170-200
#           UBERT4Rec                                models/ubert4rec/ubert4rec_model.py:19-92
#           UserSASRec                               models/user_sasrec/user_sasrec_model.py:21-107
# --------------------------------------------------------------------------------------------
_ADD_ATTR = "_sequence_embedding_layer.additional_attribute_embeddings"
_USER_ATTR = "_sequence_embedding_layer.user_attribute_embeddings"
_USER_ENC = "_sequence_representation_layer.transformer_encoder"


def user_prefixed_embedding(w: Weights, seq: torch.Tensor, attrs: Dict[str, torch.Tensor], additional: Sequence[str],
                            user: Sequence[str], item_norm: bool) -> torch.Tensor:
    """(B,S) -> (B,S+1,H): the user token (sum of the user-attribute embeddings of column 0 of each user feature) is
    prepended to the item embeddings (+ item attributes); an optional segment embedding (row 0: user, row 1: items) is added,
    then LayerNorm (+ dropout).  The item embedding itself is E[ids] + P[0..S-1] and, for UserSASRec, already normed once."""
    norm1 = (w[f"{_PRE}.embedding_norm.weight"], w[f"{_PRE}.embedding_norm.bias"]) if item_norm else None
    x = transformer_embedding(seq, w[f"{_PRE}.item_embedding.embedding.weight"], w.get(f"{_PRE}.position_embedding.weight"), norm1)
    ctx = attribute_sum(attrs, w, _ADD_ATTR, additional)
    if ctx is not None:
        x = x + ctx
    u = None
    for name in user:
        part = w[f"{_USER_ATTR}.{name}.weight"][attrs[name][:, 0:1]]
        u = part if u is None else u + part
    if u is not None:
        x = torch.cat([u, x], dim=1)
    seg = w.get("_sequence_embedding_layer.segment_embedding.weight")
    if seg is not None:
        segments = torch.ones(seq.shape, dtype=torch.int64)
        if u is not None:
            segments = torch.cat([torch.zeros(seq.shape[0], 1, dtype=torch.int64), segments], dim=1)
        x = x + seg[segments]
    return _drop(layer_norm(x, w["_sequence_embedding_layer.norm_embedding.weight"],
                            w["_sequence_embedding_layer.norm_embedding.bias"]))


def user_encoder(w: Weights, x: torch.Tensor, seq: torch.Tensor, heads: int, layers: int, has_user: bool) -> torch.Tensor:
    """causal mask over the S+1 positions; the user position is never padding (components.py:170-174, bidirectional=False)."""
    pm = padding_mask(seq)
    if has_user:
        pm = torch.cat([torch.ones(pm.shape[0], 1, dtype=pm.dtype), pm], dim=1)
    return transformer_encoder(x, w, heads, layers, attention_mask(pm, x.shape[0], x.shape[1], False), prefix=_USER_ENC)


def ubert4rec_logits(w: Weights, seq, attrs, heads, layers, additional=(), user=()) -> torch.Tensor:
    """(B,S+1,V) when user attributes are configured"""
    x = user_prefixed_embedding(w, seq, attrs, additional, user, item_norm=False)
    h = ffn_modifier(user_encoder(w, x, seq, heads, layers, bool(user)), w)
    return project(h, w["_projection_layer.linear.weight"], w["_projection_layer.linear.bias"])


def usasrec_full_logits(w: Weights, seq, attrs, heads, layers, additional=(), user=()) -> torch.Tensor:
    x = user_prefixed_embedding(w, seq, attrs, additional, user, item_norm=True)
    h = user_encoder(w, x, seq, heads, layers, bool(user))
    return project(h, w["_projection_layer.linear.weight"], w["_projection_layer.linear.bias"])


# --------------------------------------------------------------------------------------------
# 8f row 2  input pipeline                           data/datasets/processors/cloze_mask.py:50-92
#                                                    data/datasets/processors/pos_neg_sampler.py:41-114
# Per-sample restatements that draw from torch's global CPU generator exactly as the reference does
# (data/datasets/processors/utils.py:29-50), so the reference's own seeded golden vectors pin them.
# --------------------------------------------------------------------------------------------
def _random_uniform() -> float:
    return torch.empty((), dtype=torch.float, device="cpu").uniform_(0.0, 1.0).item()


def _random_int(start: int, end: int) -> int:
    """utils.py:41-50 -- documented as inclusive, but Tensor.random_(from, to) excludes `to`"""
    return torch.empty((), dtype=torch.int, device="cpu").random_(start, end).item()


def cloze_mask_sequence(sequence: Sequence[int], mask_prob: float, only_last_item_mask_prob: float, vocab_size: int,
                        mask_id: int = MASK_ID, pad_id: int = PAD_ID) -> Tuple[List[int], List[int]]:
    """(masked sequence, target) of ONE un-padded item sequence, cloze_mask.py:50-92"""
    seq = list(sequence)
    target = list(sequence)
    if _random_uniform() <= only_last_item_mask_prob:
        last = len(seq) - 1
        seq[last] = mask_id
        target[:last] = [pad_id] * last
        return seq, target
    for index in range(len(seq)):
        prob = _random_uniform()
        if prob < mask_prob:
            prob = prob / mask_prob
            if prob < 0.8:
                seq[index] = mask_id
            elif prob < 0.9:
                seq[index] = _random_int(0, vocab_size - 1)
        else:
            target[index] = pad_id
    return seq, target


def pos_neg_sequence(sequence: Sequence[int], vocab_size: int, special_ids: Sequence[int] = (PAD_ID, MASK_ID, 2)):
    """(x, pos, neg) of ONE un-padded item sequence, pos_neg_sampler.py:41-63, :86-101"""
    weights = torch.ones([vocab_size])
    weights[list(special_ids)] = 0.0
    weights[list(set(sequence))] = 0.0
    neg = torch.multinomial(weights, num_samples=len(sequence) - 1, replacement=True).tolist()
    return list(sequence[:-1]), list(sequence[1:]), neg


# --------------------------------------------------------------------------------------------
# a15-a18  ranking metrics (integer / index work: numpy)
#          metrics/container/metrics_sampler.py:45-71, metrics/common.py:4-175, metrics/mrr.py:21-37
# --------------------------------------------------------------------------------------------
FLOAT32_MIN = float(np.finfo(np.float32).min)


def multi_hot(shape: Tuple[int, int], targets: np.ndarray) -> np.ndarray:
    out = np.zeros(shape, dtype=np.int64)
    targets = np.asarray(targets)
    if targets.ndim == 1:
        targets = targets[:, None]
    np.put_along_axis(out, targets, 1, axis=1)
    return out


def sorted_item_ids(pred: np.ndarray, metric_mask: Optional[np.ndarray] = None) -> np.ndarray:
    """Descending score, ties -> ascending item id.  The reference uses an unstable argsort
    (metrics/common.py:21); the build contract fixes the tie order."""
    pred = np.asarray(pred, dtype=np.float32)
    if metric_mask is not None:
        pred = np.where(np.asarray(metric_mask) == 0, np.float32(FLOAT32_MIN), pred)
    return np.argsort(-pred.astype(np.float64), axis=1, kind="stable")


def topk_ids(pred: np.ndarray, k: int, metric_mask: Optional[np.ndarray] = None) -> np.ndarray:
    return sorted_item_ids(pred, metric_mask)[:, :k]


def true_positives(pred, positive_mask, k, metric_mask=None) -> np.ndarray:
    ids = topk_ids(pred, k, metric_mask)
    return np.take_along_axis(np.asarray(positive_mask), ids, axis=1)


def target_rank(pred: np.ndarray, target: np.ndarray) -> np.ndarray:
    """1-based rank of a single target per row: 1 + #{s_j > s_t} + #{j < t: s_j == s_t}."""
    pred = np.asarray(pred, dtype=np.float32)
    t = np.asarray(target).reshape(-1)
    st = pred[np.arange(pred.shape[0]), t][:, None]
    greater = (pred > st).sum(axis=1)
    ties_below = ((pred == st) & (np.arange(pred.shape[1])[None, :] < t[:, None])).sum(axis=1)
    return (1 + greater + ties_below).astype(np.int64)


def _dcg_weights(n: int) -> np.ndarray:
    return (1.0 / np.log2(np.arange(2, n + 2).astype(np.float32))).astype(np.float32)


def recall_at_k(pred, positive_mask, k, metric_mask=None) -> np.ndarray:
    tp = true_positives(pred, positive_mask, k, metric_mask).sum(1).astype(np.float32)
    rel = np.asarray(positive_mask).sum(1).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = tp / rel
    r[np.isnan(r)] = 0
    return r


def precision_at_k(pred, positive_mask, k, metric_mask=None) -> np.ndarray:
    return true_positives(pred, positive_mask, k, metric_mask).sum(1).astype(np.float32) / np.float32(k)


def dcg_at_k(pred, positive_mask, k, metric_mask=None) -> np.ndarray:
    tp = true_positives(pred, positive_mask, k, metric_mask).astype(np.float32)
    n = min(k, np.asarray(pred).shape[1])
    return (tp * _dcg_weights(n)[None, :]).sum(1)


def ndcg_at_k(pred, positive_mask, k, metric_mask=None) -> np.ndarray:
    dcg = dcg_at_k(pred, positive_mask, k, metric_mask)
    rel = np.minimum(np.asarray(positive_mask).sum(1), k)
    w = _dcg_weights(k)
    idcg = np.array([w[:int(r)].sum() for r in rel], dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = dcg / idcg
    out[np.isnan(out)] = 0
    return out


def mrr_at_k(pred, positive_mask, k, metric_mask=None) -> np.ndarray:
    """metrics/mrr.py:21-37: reciprocal of the LARGEST hit position within the top k (max(rank*tp))."""
    tp = true_positives(pred, positive_mask, k, metric_mask)
    n = min(np.asarray(pred).shape[1], k)
    rank = (np.arange(1, n + 1)[None, :] * tp).max(axis=1).astype(np.float32)
    with np.errstate(divide="ignore"):
        out = 1.0 / rank
    out[np.isinf(out)] = 0
    return out


def f1_at_k(pred, positive_mask, k, metric_mask=None) -> np.ndarray:
    p = precision_at_k(pred, positive_mask, k, metric_mask)
    r = recall_at_k(pred, positive_mask, k, metric_mask)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = 2 * r * p / (r + p)
    out[np.isnan(out)] = 0
    return out


def rank_full(pred, positive_mask, metric_mask=None) -> np.ndarray:
    """metrics/common.py:30-46: largest rank of a relevant item over the full sort."""
    n = np.asarray(pred).shape[1]
    tp = true_positives(pred, positive_mask, n, metric_mask)
    return (np.arange(1, n + 1)[None, :] * tp).max(axis=1)


def metrics_from_rank(rank: np.ndarray, ks: Sequence[int]) -> Dict[str, float]:
    """Single-target next-item evaluation: every @k metric is a function of the target's rank.
    Returns per-batch SUMS (the metric state of metrics/metric.py:57-96 is sum + count)."""
    rank = np.asarray(rank, dtype=np.int64)
    out = {}
    for k in ks:
        hit = rank <= k
        out[f"recall@{k}"] = float(hit.astype(np.float32).sum())
        out[f"precision@{k}"] = float((hit.astype(np.float32) / np.float32(k)).sum())
        out[f"NDCG@{k}"] = float(np.where(hit, 1.0 / np.log2(rank.astype(np.float32) + 1.0), 0.0)
                                 .astype(np.float32).sum())
        out[f"MRR@{k}"] = float(np.where(hit, 1.0 / rank.astype(np.float32), 0.0).astype(np.float32).sum())
    return out


# --------------------------------------------------------------------------------------------
# a19  initialisers (only matter for from-scratch training)
#      models/transformer/transformer_encoder_model.py:63-73, models/bert4rec/bert4rec_model.py:59-68,
#      models/common/layers/layers.py:134-136
# --------------------------------------------------------------------------------------------
def init_normal_(w: Weights, initializer_range: float, generator: torch.Generator) -> None:
    """BERT4Rec / KeBERT4Rec final init: Linear+Embedding weights N(0, range), LayerNorm 1/0,
    Linear bias 0; ``output_bias`` keeps U(+-1/sqrt(V))."""
    for name, t in w.items():
        if name.endswith("output_bias"):
            continue
        if "norm" in name or name.endswith("transform.2.weight") or name.endswith("transform.2.bias"):
            t.fill_(1.0) if name.endswith("weight") else t.zero_()
        elif name.endswith("bias"):
            t.zero_()
        else:
            t.normal_(0.0, initializer_range, generator=generator)


# --------------------------------------------------------------------------------------------
# full steps used by the CPU baseline
# --------------------------------------------------------------------------------------------
def adam_step(params: List[torch.Tensor], grads: List[torch.Tensor], m: List[torch.Tensor],
              v: List[torch.Tensor], step: int, lr: float, beta1: float = 0.99, beta2: float = 0.998,
              eps: float = 1e-8, weight_decay: float = 0.0) -> None:
    """torch.optim.Adam semantics (L2 decay added to the gradient, not AdamW); reference betas
    default to (0.99, 0.998) (modules/masked_training_module.py:37-38,165-168)."""
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    for p, g, mi, vi in zip(params, grads, m, v):
        if weight_decay != 0.0:
            g = g + weight_decay * p
        mi.mul_(beta1).add_(g, alpha=1 - beta1)
        vi.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (vi.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(mi, denom, value=-lr / bc1)
