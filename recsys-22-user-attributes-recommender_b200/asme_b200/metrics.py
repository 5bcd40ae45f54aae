"""Ranking metrics with the reference's class names, names() and state semantics
(metrics/metric.py:17-108, metrics/{recall,ndcg,mrr,precision,f1,dcg,mrr_full,rank}.py,
metrics/container/metrics_container.py:13-152, metrics/container/metrics_sampler.py:27-71).

Two entry points per metric:
  update(predictions (N,I), positive_item_mask (N,I), metric_mask=None)  -- the reference's dense signature, computed by
        the O(I) ``asme_b200_dense_ranking`` kernel (ties: score desc, item id asc);
  update_from_ranks(rank (N) int32)                                      -- the fused path: the scoring kernel already
        produced the exact 1-based target rank, no (N,I) tensor exists.
State = running sum + count; ``compute()`` = sum / count; ``forward`` returns the batch-local value while accumulating
(what torchmetrics' ``forward`` does). States sync across ranks with one all-reduce(SUM) (``sync``), matching
``dist_reduce_fx="sum"``.
"""
from enum import Enum
from typing import Dict, List, Optional

import torch
from torch import nn

from . import ops


class MetricStorageMode(Enum):
    PER_SAMPLE = "per_sample"
    SUM = "sum"


_DENSE_ROW = {"recall": 0, "precision": 1, "dcg": 2, "ndcg": 3, "mrr": 4, "f1": 5, "rank": 6, "mrr_full": 6}
_FUSED_ROW = {"recall": 0, "ndcg": 1, "mrr": 2, "precision": 3}     # rows of asme_b200_ranking_metrics


class RankingMetric(nn.Module):
    def __init__(self, metric_id: str, k: Optional[int] = None, storage_mode: MetricStorageMode = MetricStorageMode.SUM,
                 dist_sync_on_step: bool = False):
        super().__init__()
        self._metric_id = metric_id
        self._k = k
        self._storage_mode = storage_mode if isinstance(storage_mode, MetricStorageMode) else MetricStorageMode.SUM
        self.register_buffer("_sum", torch.tensor(0.0), persistent=False)
        self.register_buffer("count", torch.tensor(0), persistent=False)
        self._per_sample: List[torch.Tensor] = []

    def set_metrics_storage_mode(self, storage_mode: MetricStorageMode):
        self._storage_mode = storage_mode
        self.reset()

    # ---- per-row values ---------------------------------------------------------------------------------
    def _values_dense(self, predictions, positive_item_mask, metric_mask) -> torch.Tensor:
        k = self._k if self._k is not None else 1
        table = ops.dense_ranking(predictions, positive_item_mask, metric_mask, min(k, 32))
        if self._k is not None and self._k > 32:
            raise RuntimeError("asme_b200: k > 32 is not supported by the top-k kernels")
        v = table[_DENSE_ROW[self._metric_id]]
        if self._metric_id == "mrr_full":
            v = torch.where(v > 0, 1.0 / v, torch.zeros_like(v))
        return v

    def _values_from_rank(self, rank: torch.Tensor) -> torch.Tensor:
        """single relevant item per row: every metric is a closed form of the target's 1-based rank."""
        r = rank.to(torch.float32)
        if self._metric_id == "rank":
            return r
        if self._metric_id == "mrr_full":
            return 1.0 / r
        hit = (rank <= self._k).to(torch.float32)
        if self._metric_id == "recall":
            return hit
        if self._metric_id == "precision":
            return hit / self._k
        if self._metric_id in ("ndcg", "dcg"):
            return hit / torch.log2(r + 1.0)
        if self._metric_id == "mrr":
            return hit / r
        if self._metric_id == "f1":
            p = hit / self._k
            return torch.where(hit > 0, 2 * hit * p / (hit + p), torch.zeros_like(hit))
        raise KeyError(self._metric_id)

    # ---- state ----------------------------------------------------------------------------------------------
    def _accumulate(self, values: torch.Tensor):
        if self._sum.device != values.device:
            self._sum = self._sum.to(values.device)
            self.count = self.count.to(values.device)
        if self._storage_mode == MetricStorageMode.PER_SAMPLE:
            self._per_sample.append(values)
        self._sum = self._sum + values.sum()
        self.count = self.count + values.shape[0]

    def update(self, predictions: torch.Tensor, positive_item_mask: torch.Tensor, metric_mask: torch.Tensor = None) -> None:
        self._accumulate(self._values_dense(predictions, positive_item_mask, metric_mask))

    def update_from_ranks(self, rank: torch.Tensor) -> torch.Tensor:
        values = self._values_from_rank(rank)
        self._accumulate(values)
        return values.mean()

    def update_from_sum(self, batch_sum: torch.Tensor, n: int) -> torch.Tensor:
        """accumulate a batch sum computed by ``asme_b200_ranking_metrics`` (fused path, K20)."""
        if self._sum.device != batch_sum.device:
            self._sum = self._sum.to(batch_sum.device)
            self.count = self.count.to(batch_sum.device)
        self._sum = self._sum + batch_sum
        self.count = self.count + n
        return batch_sum / n

    def forward(self, predictions, positive_item_mask, metric_mask=None) -> torch.Tensor:
        values = self._values_dense(predictions, positive_item_mask, metric_mask)
        self._accumulate(values)
        return values.sum() / values.shape[0]

    def compute(self) -> torch.Tensor:
        return self._sum / self.count

    def raw_metric_values(self):
        return self._per_sample if self._storage_mode == MetricStorageMode.PER_SAMPLE else self._sum

    def reset(self):
        self._sum = torch.zeros_like(self._sum)
        self.count = torch.zeros_like(self.count)
        self._per_sample = []

    def sync(self, group=None):
        """all-reduce(SUM) of (sum, count) across ranks -- the reference's dist_reduce_fx="sum"."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            packed = torch.stack([self._sum.to(torch.float64), self.count.to(torch.float64)])
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
            self._sum = packed[0].to(torch.float32)
            self.count = packed[1].to(self.count.dtype)

    def name(self) -> str:
        raise NotImplementedError


def _metric_class(metric_id: str, label: str, has_k: bool = True):
    if has_k:
        class _M(RankingMetric):
            def __init__(self, k: int, dist_sync_on_step: bool = False, storage_mode: MetricStorageMode = MetricStorageMode.SUM):
                super().__init__(metric_id, k, storage_mode, dist_sync_on_step)

            def name(self):
                return f"{label}@{self._k}"
    else:
        class _M(RankingMetric):
            def __init__(self, dist_sync_on_step: bool = False, storage_mode: MetricStorageMode = MetricStorageMode.SUM):
                super().__init__(metric_id, None, storage_mode, dist_sync_on_step)

            def name(self):
                return label
    return _M


RecallMetric = _metric_class("recall", "recall")
NormalizedDiscountedCumulativeGainMetric = _metric_class("ndcg", "NDCG")
DiscountedCumulativeGainMetric = _metric_class("dcg", "DCG")
MRRMetric = _metric_class("mrr", "MRR")
PrecisionMetric = _metric_class("precision", "precision")
F1Metric = _metric_class("f1", "F1")
MRRFullMetric = _metric_class("mrr_full", "MRR", has_k=False)
Rank = _metric_class("rank", "rank", has_k=False)
for _cls, _name in ((RecallMetric, "RecallMetric"), (NormalizedDiscountedCumulativeGainMetric, "NormalizedDiscountedCumulativeGainMetric"),
                    (DiscountedCumulativeGainMetric, "DiscountedCumulativeGainMetric"), (MRRMetric, "MRRMetric"),
                    (PrecisionMetric, "PrecisionMetric"), (F1Metric, "F1Metric"), (MRRFullMetric, "MRRFullMetric"), (Rank, "Rank")):
    _cls.__name__ = _cls.__qualname__ = _name

METRIC_REGISTRY = {"recall": RecallMetric, "ndcg": NormalizedDiscountedCumulativeGainMetric, "dcg": DiscountedCumulativeGainMetric,
                   "mrr": MRRMetric, "precision": PrecisionMetric, "f1": F1Metric, "mrr_full": MRRFullMetric, "rank": Rank}


# ------------------------------------------------------------------------------------------------
# samplers / containers
# ------------------------------------------------------------------------------------------------
class FusedPredictions:
    """What the B200 eval step hands to the metrics instead of a dense (N,I) logits tensor: the exact 1-based
    target rank and the top-k list, produced inside the scoring kernel (the logits never reached HBM)."""

    def __init__(self, rank: torch.Tensor, topk_idx: torch.Tensor, topk_val: torch.Tensor, target_score: torch.Tensor,
                 num_items: int, scorer=None, lse: Optional[torch.Tensor] = None):
        self.rank, self.topk_idx, self.topk_val, self.target_score, self.num_items = rank, topk_idx, topk_val, target_score, num_items
        # log-sum-exp of every row over the whole catalog (predict: softmax score of a listed item = exp(topk_val - lse))
        self.lse = lse
        # scorer(items (N,I) int64) -> (N,I) fp32 scores of chosen items for the same rows: what the item-subset samplers gather
        self.scorer = scorer

    def gather(self, dim: int, items: torch.Tensor) -> torch.Tensor:
        """``predictions.gather(1, items)`` of the dense tensor this object stands for"""
        if dim != 1 or self.scorer is None:
            raise RuntimeError("asme_b200: this evaluation step carries no item scorer (sampled / fixed-subset metrics need one)")
        return self.scorer(items)

    def size(self):
        n = self.rank.shape[0] if self.rank is not None else self.topk_idx.shape[0]
        return torch.Size([n, self.num_items])

    @property
    def shape(self):
        """(N, I) of the dense tensor this object stands for -- the reference's prediction writers read ``logits.shape[0]``"""
        return self.size()


class MetricsSample:
    def __init__(self, sampled_predictions, positive_item_mask, metric_mask):
        self.sampled_predictions, self.positive_item_mask, self.metric_mask = sampled_predictions, positive_item_mask, metric_mask


class AllItemsSampler:
    """metrics_sampler.py:45-71 -- all items are candidates.  For dense predictions the multi-hot target matrix is
    built (compatibility); fused predictions pass straight through."""

    def sample(self, input_seq, targets, predictions, mask=None) -> MetricsSample:
        if isinstance(predictions, FusedPredictions):
            return MetricsSample(predictions, None, None)
        multihot = torch.zeros(predictions.shape, dtype=torch.long, device=predictions.device)
        t = targets.unsqueeze(1) if targets.dim() == 1 else targets
        multihot = multihot.scatter(1, t, torch.ones_like(t))
        return MetricsSample(predictions, multihot, None)

    def suffix_metric_name(self) -> str:
        return ""


class FixedItemsSampler:
    """metrics_sampler.py:74-107 -- only the configured items are candidates (single-target recommendation)"""

    def __init__(self, fixed_items: List[int]):
        self.fixed_items = list(fixed_items)

    def sample(self, input_seq, targets, predictions, mask=None) -> MetricsSample:
        if targets.dim() != 1:
            raise NotImplementedError("basket targets are outside the B200 hot path")
        cache = self.__dict__.setdefault("_items_device", {})       # (built once per device: a per-step pageable H2D copy synchronises)
        row = cache.get(str(targets.device))
        if row is None:
            row = cache[str(targets.device)] = torch.tensor(self.fixed_items, dtype=torch.int64, device=targets.device)
        items = row.unsqueeze(0).repeat(targets.shape[0], 1)
        sampled = predictions.gather(1, items)
        positive = items.eq(targets.unsqueeze(1)).to(dtype=sampled.dtype)
        return MetricsSample(sampled, positive, None)

    def suffix_metric_name(self) -> str:
        return "_fixed"


class NegativeMetricsSampler:
    """metrics_sampler.py:140-204 -- the target plus ``sample_size`` negatives drawn (without replacement) from the item weights,
    never the target nor an item of the input sequence.  The draw runs on the GPU (csrc/pipeline.cu) from the library's
    counter-based generator: same distribution, not the same draws as torch.multinomial."""

    def __init__(self, weights: List[float], sample_size: int, metrics_suffix: str, seed: int = 0):
        self.weights = weights
        self.sample_size = int(sample_size)
        self.metrics_suffix = metrics_suffix
        self.seed = int(seed)
        self._calls = 0
        self._cdf = None

    def _cdf_on(self, device):
        if self._cdf is None or self._cdf.device != device:
            self._cdf = torch.cumsum(torch.as_tensor(self.weights, dtype=torch.float64), 0).to(device)
        return self._cdf

    def sample(self, input_seq, targets, predictions, mask=None) -> MetricsSample:
        if targets.dim() != 1 or input_seq.dim() != 2:
            raise NotImplementedError("basket inputs / targets are outside the B200 hot path")
        self._calls += 1
        negatives, failed = ops.weighted_negatives(self._cdf_on(targets.device), input_seq, targets, self.sample_size,
                                                   (self.seed << 32) + self._calls)
        items = torch.cat([targets.unsqueeze(1), negatives], dim=1)
        sampled = predictions.gather(1, items)
        positive = items.eq(targets.unsqueeze(1)).to(dtype=sampled.dtype)
        self.last_failed = failed          # device flag: a user had fewer admissible items than sample_size (torch.multinomial raises)
        return MetricsSample(sampled, positive, torch.ones_like(items))

    def suffix_metric_name(self) -> str:
        return self.metrics_suffix


class MetricsContainer(nn.Module):
    pass


class RankingMetricsContainer(MetricsContainer):
    def __init__(self, metrics: List[RankingMetric], sampler):
        super().__init__()
        self.metrics = nn.ModuleList(metrics)
        self.sampler = sampler

    def update(self, input_seq, targets, predictions, mask=None) -> Dict[str, torch.Tensor]:
        samples = self.sampler.sample(input_seq, targets, predictions, mask)
        results = {}
        fused = isinstance(samples.sampled_predictions, FusedPredictions)
        table, ks = None, []
        if fused:      # one kernel for every recall / NDCG / MRR / precision @k of this container
            rank = samples.sampled_predictions.rank
            ks = sorted({m._k for m in self.metrics if m._metric_id in _FUSED_ROW and m._storage_mode == MetricStorageMode.SUM})
            if ks:
                if len(ks) > 8:
                    raise RuntimeError("asme_b200: at most 8 distinct k per metrics container")
                # the k values live on the device once per (device, ks): building the tensor from the Python list every step is a
                # pageable host-to-device copy, and CUDA synchronises the stream before one -- the host then sat out the whole
                # replayed model graph before it could queue the metric kernels (~60 us of idle device per evaluation step)
                cache = self.__dict__.setdefault("_ks_device", {})
                ck = (str(rank.device), tuple(ks))
                ks_t = cache.get(ck)
                if ks_t is None:
                    ks_t = cache[ck] = torch.tensor(ks, dtype=torch.int32, device=rank.device)
                table = torch.zeros(4, len(ks), dtype=torch.float32, device=rank.device)
                ops.ranking_metrics(rank, ks_t, table)
        for metric in self.metrics:
            if fused and table is not None and metric._metric_id in _FUSED_ROW and metric._k in ks:
                step_value = metric.update_from_sum(table[_FUSED_ROW[metric._metric_id], ks.index(metric._k)], rank.shape[0])
            elif fused:
                step_value = metric.update_from_ranks(samples.sampled_predictions.rank)
            else:
                step_value = metric(samples.sampled_predictions, samples.positive_item_mask, samples.metric_mask)
            results[f"{metric.name()}{self.sampler.suffix_metric_name()}"] = step_value
        return results

    def compute(self) -> Dict[str, torch.Tensor]:
        return {f"{m.name()}{self.sampler.suffix_metric_name()}": m.compute() for m in self.metrics}

    def reset(self):
        for m in self.metrics:
            m.reset()

    def sync(self, group=None):
        for m in self.metrics:
            m.sync(group)

    def get_metric_names(self) -> List[str]:
        return [f"{m.name()}{self.sampler.suffix_metric_name()}" for m in self.metrics]

    def get_metrics(self):
        return self.metrics

    def max_k(self) -> int:
        return max([m._k for m in self.metrics if m._k is not None] + [1])

    def needs_full_rank(self) -> bool:
        """only ``rank`` and the full ``MRR`` need the target's exact rank beyond the top-k list"""
        return any(m._metric_id in ("rank", "mrr_full") for m in self.metrics)


class AggregateMetricsContainer(MetricsContainer):
    def __init__(self, containers: List[MetricsContainer]):
        super().__init__()
        self.containers = nn.ModuleList(containers)

    def update(self, input_seq, targets, predictions, mask=None):
        results = {}
        for c in self.containers:
            results.update(c.update(input_seq, targets, predictions, mask) or {})
        return results

    def compute(self):
        results = {}
        for c in self.containers:
            results.update(c.compute() or {})
        return results

    def reset(self):
        for c in self.containers:
            c.reset()

    def sync(self, group=None):
        for c in self.containers:
            c.sync(group)

    def get_metric_names(self):
        return [n for c in self.containers for n in c.get_metric_names()]

    def get_metrics(self):
        return [m for c in self.containers for m in c.get_metrics()]

    def max_k(self) -> int:
        return max([c.max_k() for c in self.containers] + [1])

    def needs_full_rank(self) -> bool:
        return any(getattr(c, "needs_full_rank", lambda: True)() for c in self.containers)


def build_metrics(spec: Dict[str, List[int]], sampler=None) -> AggregateMetricsContainer:
    """{"recall": [1,5,10], "ndcg": [1,5,10]} -> container (the shape of the reference's ``metrics.full`` config section)."""
    metrics = []
    for key, ks in spec.items():
        cls = METRIC_REGISTRY[key]
        if key in ("mrr_full", "rank"):
            metrics.append(cls())
        else:
            metrics += [cls(k) for k in ks]
    return AggregateMetricsContainer([RankingMetricsContainer(metrics, sampler or AllItemsSampler())])


# ------------------------------------------------------------------------------------------------
# containers built by the reference's own factories
# ------------------------------------------------------------------------------------------------
_CLASS_TO_ID = {"RecallMetric": "recall", "NormalizedDiscountedCumulativeGainMetric": "ndcg", "DiscountedCumulativeGainMetric": "dcg",
                "MRRMetric": "mrr", "PrecisionMetric": "precision", "F1Metric": "f1", "MRRFullMetric": "mrr_full", "Rank": "rank"}


def _adopt_sampler(sampler):
    kind = type(sampler).__name__
    if kind == "AllItemsSampler":
        return AllItemsSampler()
    if kind == "FixedItemsSampler":
        return FixedItemsSampler(sampler.fixed_items)
    if kind == "NegativeMetricsSampler":
        return NegativeMetricsSampler(sampler.weights, sampler.sample_size, sampler.metrics_suffix)
    raise NotImplementedError(f"asme_b200: metrics sampler {kind} has no B200 counterpart")


def adopt_metrics(container):
    """The ASME module factory builds the ``metrics`` argument with the reference's OWN ``MetricsContainerFactory``
    (init/factories/modules/modules.py:88, init/factories/metrics/metrics_container.py:31-33): an
    ``AggregateMetricsContainer`` of ``RankingMetricsContainer(metrics, sampler)`` whose metrics sort dense (N,I) logits.  The
    B200 evaluation step has no dense logits, so such a container is re-expressed -- same metric names, k values, sampler
    parameters and name suffixes -- with the classes of this module.  Containers of this module pass through unchanged."""
    if container is None or isinstance(container, MetricsContainer):
        return container
    groups = getattr(container, "containers", None)
    if groups is None:
        groups = [container]
    adopted = []
    for group in groups:
        metrics = []
        for metric in group.metrics:
            metric_id = _CLASS_TO_ID.get(type(metric).__name__)
            if metric_id is None:
                raise NotImplementedError(f"asme_b200: metric {type(metric).__name__} has no B200 counterpart")
            cls = METRIC_REGISTRY[metric_id]
            mode = getattr(getattr(metric, "_storage_mode", None), "name", "SUM")
            kwargs = dict(storage_mode=MetricStorageMode[mode])
            metrics.append(cls(**kwargs) if metric_id in ("mrr_full", "rank") else cls(int(metric._k), **kwargs))
        adopted.append(RankingMetricsContainer(metrics, _adopt_sampler(group.sampler)))
    return AggregateMetricsContainer(adopted)
