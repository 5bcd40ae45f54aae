"""Training modules with the reference's class names, constructor keywords, batch keys and hook protocol
(modules/masked_training_module.py:21-189, modules/next_item_prediction_training_module.py:141-255,
modules/sequence_next_item_prediction_training_module.py:22-185, modules/metrics_trait.py:11-56):

    training_step(batch, idx) -> {"loss": t}        loss.backward() runs the fused CUDA backward
    validation_step / test_step -> {sequence, predictions, targets[, mask]}
    validation_step_end / *_epoch_end                metrics update / compute / reset
    configure_optimizers()

They do not need pytorch_lightning: ``asme_b200.trainer.Trainer`` (or Lightning itself, if present) drives the
hooks.  ``predictions`` is a :class:`~asme_b200.metrics.FusedPredictions` (rank + top-k from the fused scoring
kernel) instead of a dense (N,V) tensor; ``fused_eval=False`` restores the dense reference behaviour.
"""
from typing import Any, Dict, List, Optional

import torch
from torch import nn

from . import ops
from .data import InputSequence
from .inject import resolve_tokenizer
from .metrics import FusedPredictions, MetricsContainer, adopt_metrics
from .models import MASK_TOKEN_ID, PAD_TOKEN_ID, TransformerRecommenderModel, last_position_rows, mask_position_rows, score_rows

ITEM_SEQ_ENTRY_NAME = "item"                 # asme/data/datasets/__init__.py
TARGET_ENTRY_NAME = "item.target"
POSITIVE_SAMPLES_ENTRY_NAME = "positive_samples"
NEGATIVE_SAMPLES_ENTRY_NAME = "negative_samples"
LOG_KEY_VALIDATION_LOSS, LOG_KEY_TEST_LOSS, LOG_KEY_TRAINING_LOSS = "val_loss", "test_loss", "train_loss"
RETURN_KEY_SEQUENCE, RETURN_KEY_PREDICTIONS, RETURN_KEY_TARGETS, RETURN_KEY_MASK = "sequence", "predictions", "targets", "mask"


class _FusedLoss(torch.autograd.Function):
    """Bridges the fused forward/backward into autograd: ``loss.backward()`` launches the CUDA backward, which writes
    straight into the flat gradient arena (param.grad views); autograd itself sees a single node."""

    @staticmethod
    def forward(ctx, anchor, loss_value, backward_fn):
        ctx.backward_fn = backward_fn
        return loss_value.clone()

    @staticmethod
    def backward(ctx, grad_out):
        ctx.backward_fn()
        return None, None, None


def fused_loss(model: TransformerRecommenderModel, loss_value: torch.Tensor, backward_fn) -> torch.Tensor:
    if not torch.is_grad_enabled():
        # direct mode (asme_b200.graphs.GraphedTrainStep captures the step under no_grad and runs the fused backward itself:
        # the autograd engine's worker thread cannot take part in a CUDA-graph capture of these ctypes launches)
        model._pending_backward = backward_fn
        return loss_value
    anchor = next(model.parameters())
    return _FusedLoss.apply(anchor, loss_value, backward_fn)


def get_padding_mask(sequence: torch.Tensor, pad_token_id: int) -> torch.Tensor:
    """modules/util/module_util.py:13-30"""
    if sequence.dim() > 2:
        sequence = sequence.max(dim=2).values
    return sequence.ne(pad_token_id)


def get_additional_meta_data(model, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """modules/util/module_util.py:106-130 (same error behaviour)"""
    metadata = {}
    for key in model.required_metadata_keys():
        if key not in batch:
            raise Exception(f"The batch does not contain the following additional metadata: {key}. "
                            f"Found the following batch entries: {', '.join(batch.keys())}")
        metadata[key] = batch[key]
    for key in model.optional_metadata_keys():
        if key in batch:
            metadata[key] = batch[key]
    return metadata


def build_eval_step_return_dict(input_sequence, predictions, targets, mask=None):
    out = {RETURN_KEY_SEQUENCE: input_sequence, RETURN_KEY_PREDICTIONS: predictions, RETURN_KEY_TARGETS: targets}
    if mask is not None:
        out[RETURN_KEY_MASK] = mask
    return out


class _TokenIds:
    """the three ids the modules need from the item tokenizer (a real asme Tokenizer also works)"""

    def __init__(self, pad_token_id=PAD_TOKEN_ID, mask_token_id=MASK_TOKEN_ID):
        self.pad_token_id, self.mask_token_id = pad_token_id, mask_token_id


def _item_tokenizer(given):
    """``@inject(item_tokenizer=InjectTokenizer("item"))`` (modules/masked_training_module.py:28,
    next_item_prediction_training_module.py:34, sequence_next_item_prediction_training_module.py:34): the ASME factory passes None
    and the tokenizer comes from the build context; outside a container the default special-token ids are used"""
    tok = resolve_tokenizer("item", given)
    return tok if tok is not None else _TokenIds()


try:        # inside an ASME process the modules ARE LightningModules (pl.Trainer type-checks what it is given); without
    import pytorch_lightning as _pl          # Lightning (this image) asme_b200.trainer.Trainer drives the same hooks
    _LightningBase = _pl.LightningModule
    if not (isinstance(_LightningBase, type) and issubclass(_LightningBase, nn.Module)):
        _LightningBase = nn.Module
except Exception:
    _LightningBase = nn.Module


class _ModuleBase(_LightningBase):
    """MetricsTrait + the slice of the LightningModule protocol the runner uses."""

    def __init__(self):
        super().__init__()
        self.logged: Dict[str, Any] = {}
        self.fused_eval = True        # eval steps return FusedPredictions (rank + top-k) instead of dense (N,V) logits
        self.eval_graph = False       # replay the model part of a fused evaluation step from a CUDA graph (all-item metrics only)
        self.eval_loss = True         # also log val_loss / test_loss (one extra fused CE pass over the selected rows)
        self.use_fused_adam = True

    def log(self, name, value, *args, **kwargs):
        self.logged[name] = value
        if _LightningBase is not nn.Module and (getattr(self, "_trainer", None) or self.__dict__.get("trainer")) is not None:
            super().log(name, value, *args, **kwargs)          # attached to a pl.Trainer: loggers, checkpoint monitors, early stopping

    def save_hyperparameters(self, *args, **kwargs):
        pass

    def get_metrics(self) -> MetricsContainer:
        return self.metrics

    # ---- MetricsTrait (modules/metrics_trait.py:21-56) ----
    def _eval_step_end(self, outputs):
        metrics = self.get_metrics()
        values = metrics.update(outputs[RETURN_KEY_SEQUENCE], outputs[RETURN_KEY_TARGETS], outputs[RETURN_KEY_PREDICTIONS],
                                mask=outputs.get(RETURN_KEY_MASK))
        for name, v in values.items():
            self.log(name, v, prog_bar=True)
        return dict(values)

    def _eval_epoch_end(self, outputs=None):
        result = self.get_metrics().compute()
        for name, value in result.items():
            self.log(name, value, prog_bar=True)
        self.get_metrics().reset()
        return result

    def validation_step_end(self, outputs):
        return self._eval_step_end(outputs)

    def test_step_end(self, outputs):
        return self._eval_step_end(outputs)

    def validation_epoch_end(self, outputs=None):
        return self._eval_epoch_end(outputs)

    def test_epoch_end(self, outputs=None):
        return self._eval_epoch_end(outputs)

    # ---- fused evaluation, optionally replayed from a CUDA graph
    def _rank(self, seq, pm, meta, targets, rows_fn=None, **kw):
        """``model.evaluate_rank`` for the evaluation hooks.  With ``self.eval_graph = True`` the model part of the step (~40
        launches, host-bound when issued from Python) is captured once per batch signature and replayed
        (:class:`asme_b200.graphs.GraphedEvalStep`); the result then carries no item scorer, i.e. only all-item metrics."""
        def run(b):
            meta_b = {k[5:]: v for k, v in b.items() if k.startswith("meta.")}
            rows = rows_fn(b["seq"], b["pm"]) if rows_fn is not None else None
            extra = dict(rows=rows, rows_one_per_sequence=True) if rows is not None else {}
            if self._catalog_is_sharded():
                return self.model.evaluate_rank_sharded(b["seq"], b["pm"], meta_b, b["target"], **extra, **kw)
            return self.model.evaluate_rank(b["seq"], b["pm"], meta_b, b["target"], **extra, **kw)
        batch = {"seq": seq, "pm": pm, "target": targets, **{f"meta.{k}": v for k, v in meta.items()}}
        if not getattr(self, "eval_graph", False) or not seq.is_cuda:
            return run(batch)
        from .graphs import GraphedEvalStep
        graphs = self.__dict__.setdefault("_eval_graphs", {})
        key = tuple(sorted((k, str(v)) for k, v in kw.items()))
        if key not in graphs:
            graphs[key] = GraphedEvalStep(run)
        return dict(graphs[key](batch))

    def _catalog_is_sharded(self) -> bool:
        """``module.shard_catalog = True`` under torch.distributed: every rank evaluates its own users, the catalog sweep is
        vocab-sharded over the ranks (asme_b200.sharded; the result carries no item scorer, i.e. all-item metrics only)"""
        import torch.distributed as dist
        return bool(getattr(self, "shard_catalog", False)) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    # ---- predict command on the fused path (evaluation/evaluation.py's evaluators consume the result, asme_b200/evaluation.py)
    def _prediction_rows(self, seq, padding_mask):
        """flat rows whose hidden state predicts the next item; None = the last real position of every sequence"""
        return None

    @torch.no_grad()
    def predict_topn(self, batch, num_predictions: int, with_rank: bool = False) -> FusedPredictions:
        """what ``predict_step`` + ``softmax`` + ``sort`` + ``[:, :num_predictions]`` deliver, without the (N, I) logits: the
        top-n list (best first) and the row log-sum-exp, straight from the scoring sweeps.  ``with_rank`` (the batch must carry
        single targets): also the exact rank and score of every target, which is what per-sample metrics are computed from."""
        seq = batch[ITEM_SEQ_ENTRY_NAME]
        pm = get_padding_mask(seq, self.item_tokenizer.pad_token_id)
        rows = self._prediction_rows(seq, pm)
        if with_rank:
            extra = dict(rows=rows, rows_one_per_sequence=True) if rows is not None else dict(select="last")
            out = self.model.evaluate_rank(seq, pm, get_additional_meta_data(self.model, batch), batch[TARGET_ENTRY_NAME],
                                           k=num_predictions, with_loss=True, pad_id=self.item_tokenizer.pad_token_id, full_rank=True,
                                           **extra)
            return FusedPredictions(out["rank"], out["topk_idx"], out["topk_val"], out["target_score"], self.model.item_vocab_size,
                                    lse=out["lse"])
        out = self.model.recommend(seq, pm, get_additional_meta_data(self.model, batch), num_predictions, rows=rows, select="last",
                                   rows_one_per_sequence=rows is not None)
        return FusedPredictions(None, out["topk_idx"], out["topk_val"], None, self.model.item_vocab_size, lse=out["lse"])

    def _eval_k(self) -> int:
        return min(32, max(self.metrics.max_k(), 1)) if hasattr(self.metrics, "max_k") else 10

    def _full_rank(self) -> bool:
        """the exact target rank (a second, count-only sweep) is needed by ``rank`` / full ``MRR`` -- and by every metric whose k
        exceeds the 32 entries the fused list holds: without it a miss would be ranked 33 and counted as a hit @50"""
        fn = getattr(self.metrics, "needs_full_rank", None)
        if fn is None or bool(fn()):
            return True
        return hasattr(self.metrics, "max_k") and self.metrics.max_k() > 32

    def _adam(self, weight_decay: float):
        if self.use_fused_adam:
            from .optim import FusedAdam
            return FusedAdam(self.model, lr=self.learning_rate, betas=(self.beta_1, self.beta_2), weight_decay=weight_decay)
        return torch.optim.Adam(self.parameters(), lr=self.learning_rate, betas=(self.beta_1, self.beta_2),
                                weight_decay=weight_decay)


class MaskedTrainingModule(_ModuleBase):
    """cloze training (BERT4Rec, KeBERT4Rec); evaluation predicts the single MASK appended to each sequence."""

    def __init__(self, model: TransformerRecommenderModel, item_tokenizer=None, metrics: MetricsContainer = None,
                 learning_rate: float = 0.001, beta_1: float = 0.99, beta_2: float = 0.998, weight_decay: float = 0.001,
                 num_warmup_steps: int = 10000):
        super().__init__()
        self.model = model
        self.learning_rate, self.beta_1, self.beta_2 = learning_rate, beta_1, beta_2
        self.weight_decay = weight_decay          # accepted and ignored, exactly like the reference (:165-168)
        self.num_warmup_steps = num_warmup_steps
        self.item_tokenizer = _item_tokenizer(item_tokenizer)
        self.metrics = adopt_metrics(metrics)

    def _input(self, batch):
        seq = batch[ITEM_SEQ_ENTRY_NAME]
        return seq, get_padding_mask(seq, self.item_tokenizer.pad_token_id), get_additional_meta_data(self.model, batch)

    def forward(self, batch, batch_idx=None) -> torch.Tensor:
        seq, pm, meta = self._input(batch)
        return self.model(InputSequence(seq, pm, meta))

    def training_step(self, batch, batch_idx):
        seq, pm, meta = self._input(batch)
        target = batch[TARGET_ENTRY_NAME]
        if target.dim() > 2:
            raise NotImplementedError("basket targets are outside the B200 hot path")
        loss, ctx = self.model.loss_ce(seq, pm, meta, target, self.item_tokenizer.pad_token_id)
        loss = fused_loss(self.model, loss, lambda: self.model.loss_ce_backward(ctx))
        self.log(LOG_KEY_TRAINING_LOSS, loss, prog_bar=False)
        return {"loss": loss}

    def _prediction_rows(self, seq, padding_mask):
        from .models import mask_position_rows
        return mask_position_rows(seq, self.item_tokenizer.mask_token_id)

    def _get_prediction_for_masked_item(self, batch, batch_idx):
        seq = batch[ITEM_SEQ_ENTRY_NAME]
        return self(batch, batch_idx)[seq.eq(self.item_tokenizer.mask_token_id)]

    def _eval_step(self, batch, batch_idx, is_test=False):
        seq, pm, meta = self._input(batch)
        targets = batch[TARGET_ENTRY_NAME]
        if not self.fused_eval or targets.dim() != 1:
            prediction = self._get_prediction_for_masked_item(batch, batch_idx)
            return build_eval_step_return_dict(seq, prediction, targets)
        out = self._rank(seq, pm, meta, targets, k=self._eval_k(), select="mask",
                         mask_id=self.item_tokenizer.mask_token_id, with_loss=self.eval_loss,
                         pad_id=self.item_tokenizer.pad_token_id, full_rank=self._full_rank())
        if self.eval_loss:
            self.log(LOG_KEY_TEST_LOSS if is_test else LOG_KEY_VALIDATION_LOSS, out["loss"], prog_bar=True)
        pred = FusedPredictions(out["rank"], out["topk_idx"], out["topk_val"], out["target_score"], self.model.item_vocab_size,
                                scorer=out.get("scorer"))
        return build_eval_step_return_dict(seq, pred, targets)

    def validation_step(self, batch, batch_idx):
        return self._eval_step(batch, batch_idx)

    def test_step(self, batch, batch_idx):
        return self._eval_step(batch, batch_idx, is_test=True)

    def predict_step(self, batch, batch_idx, dataloader_idx=None):
        return self._get_prediction_for_masked_item(batch, batch_idx)

    def configure_optimizers(self):
        optimizer = self._adam(0.0)
        if self.num_warmup_steps > 0:
            n = self.num_warmup_steps
            scheduler = torch.optim.lr_scheduler.LambdaLR(optimizer, lambda step: min(1.0, step / n))
            return [optimizer], [{"scheduler": scheduler, "interval": "step", "strict": True}]
        return [optimizer]


class NextItemPredictionTrainingModule(_ModuleBase):
    """sasrec-cross: per-position targets, CE over the full catalog; evaluation scores the last real position."""

    def __init__(self, model: TransformerRecommenderModel, item_tokenizer=None, metrics: MetricsContainer = None,
                 learning_rate: float = 0.001, beta_1: float = 0.99, beta_2: float = 0.998, weight_decay: float = 0,
                 loss_function=None):
        super().__init__()
        self.model = model
        self.learning_rate, self.beta_1, self.beta_2, self.weight_decay = learning_rate, beta_1, beta_2, weight_decay
        self.item_tokenizer = _item_tokenizer(item_tokenizer)
        self.metrics = adopt_metrics(metrics)
        self.loss_function = loss_function     # CE with ignore_index=pad is fused into the scoring kernel

    def forward(self, batch, batch_idx=None):
        seq = batch[ITEM_SEQ_ENTRY_NAME]
        pm = get_padding_mask(seq, self.item_tokenizer.pad_token_id)
        return self.model(InputSequence(seq, pm, get_additional_meta_data(self.model, batch)))

    def training_step(self, batch, batch_idx):
        seq = batch[ITEM_SEQ_ENTRY_NAME]
        pm = get_padding_mask(seq, self.item_tokenizer.pad_token_id)
        target = batch[TARGET_ENTRY_NAME]
        meta = get_additional_meta_data(self.model, batch)
        if target.dim() == 1:        # single target per sequence: the loss sees only the last real position
            rows = last_position_rows(seq, pm)
            full = torch.zeros_like(seq).reshape(-1)
            full[rows] = target
            target = full.view_as(seq)
        loss, ctx = self.model.loss_ce(seq, pm, meta, target, self.item_tokenizer.pad_token_id)
        loss = fused_loss(self.model, loss, lambda: self.model.loss_ce_backward(ctx))
        self.log(LOG_KEY_TRAINING_LOSS, loss)
        return {"loss": loss}

    def _extract_target_logits(self, input_seq, logits):
        pm = get_padding_mask(input_seq, self.item_tokenizer.pad_token_id)
        return logits[torch.arange(input_seq.shape[0], device=logits.device), pm.sum(dim=-1) - 1]

    def validation_step(self, batch, batch_idx):
        seq = batch[ITEM_SEQ_ENTRY_NAME]
        target = batch[TARGET_ENTRY_NAME]
        pm = get_padding_mask(seq, self.item_tokenizer.pad_token_id)
        if not self.fused_eval or target.dim() != 1:
            logits = self(batch, batch_idx)
            return build_eval_step_return_dict(seq, self._extract_target_logits(seq, logits), target)
        out = self._rank(seq, pm, get_additional_meta_data(self.model, batch), target, k=self._eval_k(),
                         select="last", with_loss=self.eval_loss, pad_id=self.item_tokenizer.pad_token_id,
                         full_rank=self._full_rank())
        if self.eval_loss:
            self.log(LOG_KEY_VALIDATION_LOSS, out["loss"], prog_bar=True)
        pred = FusedPredictions(out["rank"], out["topk_idx"], out["topk_val"], out["target_score"], self.model.item_vocab_size,
                                scorer=out.get("scorer"))
        return build_eval_step_return_dict(seq, pred, target)

    def test_step(self, batch, batch_idx):
        return self.validation_step(batch, batch_idx)

    def predict_step(self, batch, batch_idx, dataloader_idx=None):
        return self._extract_target_logits(batch[ITEM_SEQ_ENTRY_NAME], self(batch, batch_idx))

    def configure_optimizers(self):
        return self._adam(self.weight_decay)


class SequenceNextItemPredictionTrainingModule(_ModuleBase):
    """sasrec-neg: one positive and one sampled negative per position, BCE; evaluation ranks the whole vocabulary
    with h_last . E (quirk Q5: the reference's validation_step calls a non-existent self.predict -- the intended
    behaviour, predict_step, is implemented)."""

    def __init__(self, model: TransformerRecommenderModel, item_tokenizer=None, metrics: MetricsContainer = None,
                 learning_rate: float = 0.001, beta_1: float = 0.99, beta_2: float = 0.998, weight_decay: float = 1e-3,
                 loss_function=None):
        super().__init__()
        self.model = model
        self.learning_rate, self.beta_1, self.beta_2, self.weight_decay = learning_rate, beta_1, beta_2, weight_decay
        self.item_tokenizer = _item_tokenizer(item_tokenizer)
        self.metrics = adopt_metrics(metrics)
        self.loss_function = loss_function     # SASRecBinaryCrossEntropyLoss is fused into the pos/neg kernel

    def training_step(self, batch, batch_idx):
        seq = batch[ITEM_SEQ_ENTRY_NAME]
        pm = get_padding_mask(seq, self.item_tokenizer.pad_token_id)
        meta = get_additional_meta_data(self.model, batch)
        loss, ctx = self.model.loss_bce(seq, pm, meta, batch[POSITIVE_SAMPLES_ENTRY_NAME], batch[NEGATIVE_SAMPLES_ENTRY_NAME],
                                        seq.ne(self.item_tokenizer.pad_token_id))
        loss = fused_loss(self.model, loss, lambda: self.model.loss_bce_backward(ctx))
        self.log(LOG_KEY_TRAINING_LOSS, loss)
        return {"loss": loss}

    def predict_step(self, batch, batch_idx, dataloader_idx=None):
        seq = batch[ITEM_SEQ_ENTRY_NAME]
        pm = get_padding_mask(seq, self.item_tokenizer.pad_token_id)
        meta = dict(get_additional_meta_data(self.model, batch))
        return self.model(InputSequence(seq, pm, meta))          # positive_samples=None -> all vocabulary ids

    def validation_step(self, batch, batch_idx):
        seq = batch[ITEM_SEQ_ENTRY_NAME]
        targets = batch[TARGET_ENTRY_NAME]
        pm = get_padding_mask(seq, self.item_tokenizer.pad_token_id)
        if not self.fused_eval or targets.dim() != 1:
            return build_eval_step_return_dict(seq, self.predict_step(batch, batch_idx), targets)
        out = self._rank(seq, pm, get_additional_meta_data(self.model, batch), targets, k=self._eval_k(),
                         select="last", full_rank=self._full_rank())
        pred = FusedPredictions(out["rank"], out["topk_idx"], out["topk_val"], out["target_score"], self.model.item_vocab_size,
                                scorer=out.get("scorer"))
        return build_eval_step_return_dict(seq, pred, targets)

    def test_step(self, batch, batch_idx):
        return self.validation_step(batch, batch_idx)

    def configure_optimizers(self):
        return self._adam(self.weight_decay)


def _prepend_pad_column(target: torch.Tensor, pad_id: int) -> torch.Tensor:
    """targets of the S item positions -> targets of the S+1 encoder positions: the user position never contributes
    (ubert_masked_training_module.py:75-77 prepends a False/0 column, = the pad id = ignore_index)"""
    col = torch.full((target.shape[0], 1), pad_id, dtype=target.dtype, device=target.device)
    return torch.cat([col, target], dim=1)


class UBERTMaskedTrainingModule(MaskedTrainingModule):
    """modules/ubert_masked_training_module.py: cloze training of UBERT4Rec.  The model's outputs have one more position than
    the item sequence (the user token at position 0); targets and the MASK selection are shifted accordingly.  Like the
    reference (:183-186) weight decay is accepted and ignored."""

    def __init__(self, model: TransformerRecommenderModel, item_tokenizer=None, metrics: MetricsContainer = None,
                 learning_rate: float = 0.001, beta_1: float = 0.99, beta_2: float = 0.998, weight_decay: float = 0.001,
                 num_warmup_steps: int = 10000):
        super().__init__(model, item_tokenizer, metrics, learning_rate, beta_1, beta_2, weight_decay, num_warmup_steps)
        self.user_key_len = len(model.optional_metadata_keys())

    def training_step(self, batch, batch_idx):
        seq, pm, meta = self._input(batch)
        target = batch[TARGET_ENTRY_NAME]
        if target.dim() > 2:
            raise NotImplementedError("basket targets are outside the B200 hot path")
        if self.user_key_len > 0:
            target = _prepend_pad_column(target, self.item_tokenizer.pad_token_id)
        loss, ctx = self.model.loss_ce(seq, pm, meta, target, self.item_tokenizer.pad_token_id)
        loss = fused_loss(self.model, loss, lambda: self.model.loss_ce_backward(ctx))
        self.log(LOG_KEY_TRAINING_LOSS, loss, prog_bar=False)
        return {"loss": loss}

    def _mask_rows(self, seq: torch.Tensor) -> torch.Tensor:
        """flat row (into the B x (S+1) hidden states) of every sequence's MASK token (:93-101)"""
        B, S = seq.shape
        shift = 1 if self.user_key_len > 0 else 0
        pos = (seq == self.item_tokenizer.mask_token_id).to(torch.int32).argmax(dim=1).to(torch.int64)
        return torch.arange(B, device=seq.device, dtype=torch.int64) * (S + shift) + pos + shift

    def _prediction_rows(self, seq, padding_mask):
        return self._mask_rows(seq)

    def _get_prediction_for_masked_item(self, batch, batch_idx):
        seq = batch[ITEM_SEQ_ENTRY_NAME]
        logits = self(batch, batch_idx)
        return logits.reshape(-1, logits.shape[-1])[self._mask_rows(seq)]

    def _eval_step(self, batch, batch_idx, is_test=False):
        seq, pm, meta = self._input(batch)
        targets = batch[TARGET_ENTRY_NAME]
        if not self.fused_eval or targets.dim() != 1:
            prediction = self._get_prediction_for_masked_item(batch, batch_idx)
            return build_eval_step_return_dict(seq, prediction, targets)
        out = self._rank(seq, pm, meta, targets, rows_fn=lambda s_, p_: self._mask_rows(s_), k=self._eval_k(),
                         with_loss=self.eval_loss, pad_id=self.item_tokenizer.pad_token_id, full_rank=self._full_rank())
        if self.eval_loss:
            self.log(LOG_KEY_TEST_LOSS if is_test else LOG_KEY_VALIDATION_LOSS, out["loss"], prog_bar=True)
        pred = FusedPredictions(out["rank"], out["topk_idx"], out["topk_val"], out["target_score"], self.model.item_vocab_size,
                                scorer=out.get("scorer"))
        return build_eval_step_return_dict(seq, pred, targets)


class UserNextItemPredictionTrainingModule(NextItemPredictionTrainingModule):
    """modules/user_next_item_prediction_training_module.py: UserSASRec with ``mode="full"``.  Training drops the user
    position from the logits (:151-155) -- here the targets get a pad column instead, which selects the same rows.  Evaluation
    reproduces the reference's row choice exactly: ``logits[b, length_b - 1]`` of the S+1 positions (:124-135), i.e. the
    position BEFORE the last item (position 0 is the user token)."""

    def __init__(self, model: TransformerRecommenderModel, item_tokenizer=None, metrics: MetricsContainer = None,
                 learning_rate: float = 0.001, beta_1: float = 0.99, beta_2: float = 0.998, weight_decay: float = 0,
                 loss_function=None, first_item: bool = False):
        super().__init__(model, item_tokenizer, metrics, learning_rate, beta_1, beta_2, weight_decay, loss_function)
        self.user_key_len = len(model.optional_metadata_keys())
        # first_item (user_next_item_prediction_training_module.py:57-60): the logits are used as they come -- for models whose
        # user token REPLACES the first item (replace_first_item) the S outputs line up with the S targets
        self.first_item = bool(first_item)
        if self.first_item != bool(getattr(model, "replace_first_item", False)) and self.user_key_len > 0:
            raise ValueError("first_item must be set exactly when the model replaces the first item (otherwise the reference compares "
                             "S logit rows with S+1 / S-1 targets)")

    def training_step(self, batch, batch_idx):
        seq = batch[ITEM_SEQ_ENTRY_NAME]
        pm = get_padding_mask(seq, self.item_tokenizer.pad_token_id)
        target = batch[TARGET_ENTRY_NAME]
        if target.dim() != 2:
            raise NotImplementedError("UserNextItemPredictionTrainingModule: per-position targets (N,S) expected")
        if self.user_key_len > 0 and not self.first_item:
            target = _prepend_pad_column(target, self.item_tokenizer.pad_token_id)
        meta = get_additional_meta_data(self.model, batch)
        loss, ctx = self.model.loss_ce(seq, pm, meta, target, self.item_tokenizer.pad_token_id)
        loss = fused_loss(self.model, loss, lambda: self.model.loss_ce_backward(ctx))
        self.log(LOG_KEY_TRAINING_LOSS, loss)
        return {"loss": loss}

    def _target_rows(self, seq: torch.Tensor, pm: torch.Tensor) -> torch.Tensor:
        B, S = seq.shape
        S1 = S + self.model.user_prefix                          # positions of the encoder output
        last = pm.sum(dim=-1).to(torch.int64) - 1
        last = torch.where(last < 0, last + S1, last)            # advanced indexing wraps -1 around
        return torch.arange(B, device=seq.device, dtype=torch.int64) * S1 + last

    def _prediction_rows(self, seq, padding_mask):
        return self._target_rows(seq, padding_mask)

    def _extract_target_logits(self, input_seq, logits):
        pm = get_padding_mask(input_seq, self.item_tokenizer.pad_token_id)
        return logits.reshape(-1, logits.shape[-1])[self._target_rows(input_seq, pm)]

    def validation_step(self, batch, batch_idx):
        seq = batch[ITEM_SEQ_ENTRY_NAME]
        target = batch[TARGET_ENTRY_NAME]
        pm = get_padding_mask(seq, self.item_tokenizer.pad_token_id)
        if not self.fused_eval or target.dim() != 1:
            logits = self(batch, batch_idx)
            return build_eval_step_return_dict(seq, self._extract_target_logits(seq, logits), target)
        out = self._rank(seq, pm, get_additional_meta_data(self.model, batch), target, rows_fn=self._target_rows, k=self._eval_k(),
                         with_loss=self.eval_loss, pad_id=self.item_tokenizer.pad_token_id, full_rank=self._full_rank())
        if self.eval_loss:
            self.log(LOG_KEY_VALIDATION_LOSS, out["loss"], prog_bar=True)
        pred = FusedPredictions(out["rank"], out["topk_idx"], out["topk_val"], out["target_score"], self.model.item_vocab_size,
                                scorer=out.get("scorer"))
        return build_eval_step_return_dict(seq, pred, target)


REGISTRY = {
    # key -> (module class, model class name, model kwargs) ; mirrors modules/config.py:30-55
    "bert4rec": (MaskedTrainingModule, "BERT4RecModel", {}),
    "kebert4rec": (MaskedTrainingModule, "KeBERT4RecModel", {}),
    "sasrec-cross": (NextItemPredictionTrainingModule, "SASRecModel", {}),
    "sasrec-neg": (SequenceNextItemPredictionTrainingModule, "SASRecModel", {}),
    "ubert4rec": (UBERTMaskedTrainingModule, "UBERT4RecModel", {}),
    "user-sasrec-full": (UserNextItemPredictionTrainingModule, "UserSASRecModel", {"mode": "full"}),
}
