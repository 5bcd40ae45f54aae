"""Batch evaluators of the ``predict`` command (asme/core/evaluation/evaluation.py) on the fused evaluation output.

The reference evaluators receive the dense ``(N, I)`` logits of a batch and run ``softmax`` + a full ``sort`` over the
catalog to keep ``num_predictions`` items (evaluation.py:176-178, :222-224).  Here ``logits`` may also be a
:class:`asme_b200.metrics.FusedPredictions` carrying the top-n list and the row log-sum-exp straight from the scoring
sweeps (``model.recommend`` / ``module.predict_topn``): the softmax score of a listed item is ``exp(logit - lse)`` and the
list is already ordered best-first (ties -> lowest item id, where ``torch.sort`` leaves the order of ties open).
Same class names, constructor arguments, headers and return shapes as the reference; dense tensors still work (CPU
included) and follow the reference arithmetic literally.
"""
import csv
import itertools
from typing import IO, Any, Dict, List, Tuple

import numpy as np
import torch

from .metrics import FusedPredictions

ITEM_SEQ_ENTRY_NAME = "item"                 # asme/data/datasets/__init__.py
TARGET_ENTRY_NAME = "item.target"
SAMPLE_IDS = "sample_ids"
SESSION_IDENTIFIER = "session_identifier"


def _vocab_lookup(item_tokenizer) -> np.ndarray:
    """id -> token array (evaluation.py:48-52)"""
    tokens = np.asarray(item_tokenizer.vocabulary.tokens())
    ids = np.asarray(item_tokenizer.vocabulary.ids())
    lookup = np.empty(max(ids) + 1, dtype=tokens.dtype)
    lookup[ids] = tokens
    return lookup


def top_predictions(logits, num_predictions: int) -> Tuple[np.ndarray, np.ndarray]:
    """(softmax scores, item ids) of the ``num_predictions`` best items per sample, best first, as numpy arrays"""
    if isinstance(logits, FusedPredictions):
        if logits.topk_idx is None or logits.topk_idx.shape[1] < num_predictions:
            have = 0 if logits.topk_idx is None else logits.topk_idx.shape[1]
            raise RuntimeError(f"asme_b200: the fused prediction holds {have} items per sample, {num_predictions} requested")
        if logits.lse is None:
            raise RuntimeError("asme_b200: this fused prediction carries no log-sum-exp (use model.recommend / module.predict_topn)")
        val = logits.topk_val[:, :num_predictions].float()
        scores = torch.exp(val - logits.lse.float().unsqueeze(1))
        return scores.cpu().numpy(), logits.topk_idx[:, :num_predictions].to(torch.int64).cpu().numpy()
    softmax = torch.softmax(logits, dim=-1)
    scores, indices = torch.sort(softmax, dim=-1, descending=True)
    return scores[:, :num_predictions].cpu().numpy(), indices[:, :num_predictions].cpu().numpy()


class BatchEvaluator:
    """evaluation.py:12-40"""

    def evaluate(self, batch_index: int, batch: Dict[str, Any], logits) -> List[Any]:
        raise NotImplementedError

    def get_header(self) -> List[str]:
        return self.header

    def eval_samplewise(self) -> bool:
        raise NotImplementedError


class LogInputEvaluator(BatchEvaluator):
    """the input sequence as tokens, special tokens removed (evaluation.py:43-75)"""

    def __init__(self, item_tokenizer):
        self.item_tokenizer = item_tokenizer
        self.header = ["input"]
        self.vocab_lookup = _vocab_lookup(item_tokenizer)
        self.special_tokens = np.asarray(item_tokenizer.get_special_token_ids())

    def eval_samplewise(self) -> bool:
        return True

    def evaluate(self, batch_index, batch, logits) -> List[Any]:
        ids = batch[ITEM_SEQ_ENTRY_NAME]
        input_ids = np.asarray(ids.cpu() if isinstance(ids, torch.Tensor) else ids)
        tokens = self.vocab_lookup[input_ids]
        keep = np.isin(input_ids, self.special_tokens, invert=True)
        return [tokens[i][keep[i]].tolist() for i in range(tokens.shape[0])]


class ExtractSampleIdEvaluator(BatchEvaluator):
    """sample id, ``<id>_<pos>`` when the batch carries sequence positions (evaluation.py:78-118)"""

    def __init__(self, use_session_id: bool = False):
        self.header = ["SID"]
        self.use_session_id = use_session_id

    def eval_samplewise(self) -> bool:
        return True

    def evaluate(self, batch_index, batch, logits) -> List[Any]:
        sample_ids = batch[SESSION_IDENTIFIER] if self.use_session_id else batch[SAMPLE_IDS].tolist()
        positions = batch["pos"].tolist() if "pos" in batch else None
        n = logits.size()[0] if isinstance(logits, FusedPredictions) else logits.shape[0]
        return [sample_ids[i] if positions is None else f"{sample_ids[i]}_{positions[i]}" for i in range(n)]


class TrueTargetEvaluator(BatchEvaluator):
    """the true target as a one-element token list (evaluation.py:121-146)"""

    def __init__(self, item_tokenizer):
        self.header = ["target"]
        self.item_tokenizer = item_tokenizer
        self.vocab_lookup = _vocab_lookup(item_tokenizer)

    def eval_samplewise(self) -> bool:
        return True

    def evaluate(self, batch_index, batch, logits) -> List[Any]:
        targets = batch[TARGET_ENTRY_NAME].cpu().numpy()
        return [[t] for t in self.vocab_lookup[targets].tolist()]


class _TopNEvaluator(BatchEvaluator):
    def __init__(self, item_tokenizer, num_predictions: int, selected_items=None):
        self.item_tokenizer = item_tokenizer
        self.num_predictions = num_predictions
        self.selected_items = selected_items
        self.filter_items = np.asarray(self.selected_items)

    def eval_samplewise(self) -> bool:
        return False

    def _filtered(self, values: np.ndarray, indices: np.ndarray) -> List[Any]:
        """the reference filters AFTER cutting to num_predictions: fewer than num_predictions entries may remain"""
        if self.selected_items is None:
            return values.tolist()
        keep = np.isin(indices, self.filter_items)
        return [values[i][keep[i]].tolist() for i in range(values.shape[0])]


class ExtractScoresEvaluator(_TopNEvaluator):
    """softmax scores of the recommended items (evaluation.py:149-189)"""

    def __init__(self, item_tokenizer, num_predictions: int, selected_items=None):
        super().__init__(item_tokenizer, num_predictions, selected_items)
        self.header = ["score"]

    def evaluate(self, batch_index, batch, logits) -> List[Any]:
        scores, indices = top_predictions(logits, self.num_predictions)
        return self._filtered(scores, indices)


class ExtractRecommendationEvaluator(_TopNEvaluator):
    """the recommended items as tokens (evaluation.py:192-236)"""

    def __init__(self, item_tokenizer, num_predictions: int, selected_items=None):
        super().__init__(item_tokenizer, num_predictions, selected_items)
        self.header = ["recommendation"]
        self.vocab_lookup = _vocab_lookup(item_tokenizer)

    def evaluate(self, batch_index, batch, logits) -> List[Any]:
        _, indices = top_predictions(logits, self.num_predictions)
        return self._filtered(self.vocab_lookup[indices], indices)


class PerSampleMetricsEvaluator(BatchEvaluator):
    """the module's metrics per sample (evaluation.py:239-287): switches every metric of ``module.metrics`` to per-sample
    storage on first use, feeds it this batch and returns one row of metric values per sample.  Fused input: the values are
    closed forms of the target's exact rank (``FusedPredictions.rank``, from ``module.predict_topn(batch, n, with_rank=True)``);
    a selected-items filter changes the candidate set and therefore needs the dense logits."""

    def __init__(self, item_tokenizer, selected_items, module):
        self.item_tokenizer = item_tokenizer
        self.selected_items = selected_items
        self.module = module
        self.header = module.metrics.get_metric_names()
        self.samplewise_metrics_set = False

    def eval_samplewise(self) -> bool:
        return True

    def evaluate(self, batch_index, batch, logits) -> List[Any]:
        from .metrics import MetricStorageMode
        container = self.module.metrics
        if not self.samplewise_metrics_set:          # only here, so that training is not affected (:268-273)
            for metric in container.get_metrics():
                metric.set_metrics_storage_mode(MetricStorageMode.PER_SAMPLE)
            self.samplewise_metrics_set = True
        metrics = [m for m in container.get_metrics() if m._storage_mode == MetricStorageMode.PER_SAMPLE]
        if isinstance(logits, FusedPredictions):
            if self.selected_items:
                raise RuntimeError("asme_b200: per-sample metrics over selected items need the dense logits (module.predict_step)")
            if logits.rank is None:
                raise RuntimeError("asme_b200: this fused prediction carries no target rank (module.predict_topn(..., with_rank=True))")
            for metric in metrics:
                metric.update_from_ranks(logits.rank)
        else:
            targets = batch[TARGET_ENTRY_NAME].to(logits.device)
            item_mask = torch.nn.functional.one_hot(targets, logits.size()[1])
            if self.selected_items:
                chosen = torch.tensor(self.selected_items, dtype=torch.int32, device=logits.device)
                logits = torch.index_select(logits, 1, chosen)
                item_mask = torch.index_select(item_mask, 1, chosen)
            for metric in metrics:
                metric.update(logits, item_mask)
        values = [m.raw_metric_values()[batch_index].cpu().numpy() for m in metrics]
        return np.asarray(values).T.tolist()


# ---------------------------------------------------------------------------------------------------------------
# the writers of the predict command (asme/core/writer/prediction/batch_prediction_writer.py): one batch at a time into a CSV
# file.  ``logits`` is whatever the evaluators accept (dense tensor or FusedPredictions; both have ``.shape[0]``).
# ---------------------------------------------------------------------------------------------------------------
class BatchEvaluationWriter:
    def __init__(self, evaluators: List[BatchEvaluator]):
        self.evaluators = evaluators

    def init_file(self, file_handle: IO[str]):
        self.file_handle = file_handle
        self.csv_writer = csv.writer(file_handle)
        self.csv_writer.writerow(itertools.chain.from_iterable(self.headers))

    def _results(self, batch_index, batch, logits):
        return [(e.evaluate(batch_index, batch, logits), e.eval_samplewise(), e.get_header()) for e in self.evaluators]

    def write_evaluation(self, batch_index, batch, logits):
        raise NotImplementedError


def _append(row: List[Any], value, header: List[str]) -> None:
    if len(header) > 1:
        row.extend(value)
    else:
        row.append(value)


class CSVMultiLineWriter(BatchEvaluationWriter):
    """one line per (sample, recommendation position), the position in a leading ``order`` column (:33-77).  Like the reference,
    the number of lines per sample is the length of the list-valued evaluators' output for the batch's SECOND sample
    (``eval[1]``, :62) -- all samples have num_predictions entries unless a selected-items filter shortened some."""

    def __init__(self, evaluators: List[BatchEvaluator]):
        super().__init__(evaluators)
        self.headers = [["order"]] + [e.get_header() for e in evaluators]

    def write_evaluation(self, batch_index, batch, logits):
        results = self._results(batch_index, batch, logits)
        rows = []
        for sample in range(logits.shape[0]):
            num_rows = [len(values[1]) for values, samplewise, _ in results if not samplewise]
            for i in range(num_rows[0]):
                row = [str(i + 1)]
                for values, samplewise, header in results:
                    _append(row, values[sample] if samplewise else values[sample][i], header)
                rows.append(row)
        self.csv_writer.writerows(rows)


class CSVSingleLineWriter(BatchEvaluationWriter):
    """one line per sample, list-valued outputs written as lists (:79-117)"""

    def __init__(self, evaluators: List[BatchEvaluator]):
        super().__init__(evaluators)
        self.headers = [e.get_header() for e in evaluators]

    def write_evaluation(self, batch_index, batch, logits):
        results = self._results(batch_index, batch, logits)
        rows = []
        for sample in range(logits.shape[0]):
            row: List[Any] = []
            for values, _, header in results:
                _append(row, values[sample], header)
            rows.append(row)
        self.csv_writer.writerows(rows)


# ---------------------------------------------------------------------------------------------------------------
# the result writers of the evaluate command (asme/core/writer/results/results_writer.py): the metric dictionary of
# ``metrics.compute()`` (0-dim tensors here) as JSON {"recommender_id", "metrics"} or as "metric name,value" CSV lines
# ---------------------------------------------------------------------------------------------------------------
def _plain(metrics: Dict[str, Any]) -> Dict[str, float]:
    return {name: (float(value) if isinstance(value, torch.Tensor) else value) for name, value in metrics.items()}


class ResultWriter:
    def __init__(self, file_handle: IO[str]):
        self.file_handle = file_handle

    def write_overall_results(self, recommender_name: str, metrics: Dict[str, Any]):
        raise NotImplementedError


class JSONResultWriter(ResultWriter):
    def write_overall_results(self, recommender_name: str, metrics: Dict[str, Any]):
        import json
        json.dump({"recommender_id": recommender_name, "metrics": _plain(metrics)}, self.file_handle)


class CSVResultWriter(ResultWriter):
    HEADER = ["metric name", "value"]

    def __init__(self, file_handle: IO[str]):
        super().__init__(file_handle)
        self.csv_writer = csv.writer(file_handle)
        self.csv_writer.writerow(self.HEADER)

    def write_overall_results(self, recommender_name: str, metrics: Dict[str, Any]):
        self.csv_writer.writerows([[name, value] for name, value in _plain(metrics).items()] + [["recommender_id", recommender_name]])


SUPPORTED_RESULT_WRITERS = {".csv": CSVResultWriter, ".json": JSONResultWriter}


def check_file_format_supported(output_file) -> bool:
    import pathlib
    return pathlib.Path(output_file).suffix in SUPPORTED_RESULT_WRITERS


def build_result_writer(file_handle: IO[str]) -> ResultWriter:
    """writer chosen by the extension of the file behind ``file_handle`` (.json / .csv)"""
    import pathlib
    suffix = pathlib.Path(file_handle.name).suffix
    if suffix not in SUPPORTED_RESULT_WRITERS:
        raise KeyError(f"{suffix} is not a supported format to write predictions to file")
    return SUPPORTED_RESULT_WRITERS[suffix](file_handle)
