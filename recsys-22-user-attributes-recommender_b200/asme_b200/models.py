"""Drop-in model classes: ``BERT4RecModel``, ``KeBERT4RecModel``, ``SASRecModel``.

Same class names, constructor keywords and parameter (state-dict) names as the reference
(models/bert4rec/bert4rec_model.py:24-56, models/kebert4rec/kebert4rec_model.py:24-84,
models/sasrec/sasrec_model.py:28-95), including its quirks, but every stage runs
on the sm_100a kernels of ``libasme_b200.so``:

  forward(InputSequence)            -> logits (N,S,V) / (pos,neg) / (N,I)   reference-compatible, inference
  loss(...) / loss_backward(...)    -> fused training step: scoring + CE (or BCE) without materialising logits
  evaluate_rank(...)                -> fused scoring + top-k + exact target rank for the metrics

There is no PyTorch arithmetic on these paths and no CPU fallback.
"""
import math
import os
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch
from torch import nn

from . import ops
from .arena import ArenaModule
from .data import InputSequence
from .inject import resolve_tokenizers, resolve_vocab_size
from .engine import BLOCKS, MODIFIER, EncoderConfig, EncoderEngine, Saved, block_param_specs, modifier_param_specs

PAD_TOKEN_ID = 0
MASK_TOKEN_ID = 1

# Precision policy of newly built models (``model.precision`` can be changed afterwards):
#   "bf16": tensor-core path -- tcgen05 GEMMs / catalog scoring with bf16 operands and fp32 accumulation; residual stream,
#           LayerNorm, softmax statistics, losses and optimizer state stay fp32 (north_star tolerance 1e-3)
#   "fp32": strict-parity SIMT path (1e-5)
DEFAULT_PRECISION = os.environ.get("ASME_B200_PRECISION", "bf16")
# Exact top-k lists under the bf16 policy: the bf16 sweep only proposes candidates, which are re-scored from the fp32 table and
# certified (csrc/rescore.cu) -- lists, target positions and @k metrics are then those of the fp32 path for the same hidden rows.
EXACT_TOPK = os.environ.get("ASME_B200_EXACT_TOPK", "1") == "1"


def set_default_precision(precision: str) -> None:
    global DEFAULT_PRECISION
    if precision not in ("bf16", "fp32"):
        raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
    DEFAULT_PRECISION = precision

_EMB = "_sequence_embedding_layer"
_PRE_ATTR = f"{_EMB}.prefusion_attribute_embeddings"
_POST_ATTR = f"{MODIFIER}.postfusion_attribute_embeddings"
_ADD_ATTR = f"{_EMB}.additional_attribute_embeddings"          # UBERT4Rec / UserSASRec name for the pre-fused item attributes
_USER_ATTR = f"{_EMB}.user_attribute_embeddings"
_SEGMENT = f"{_EMB}.segment_embedding.weight"


def _attr_vocab_size(name, tokenizers, sizes):
    if sizes and name in sizes:
        return int(sizes[name])
    if tokenizers is not None:
        tok = tokenizers.get("tokenizers." + name, tokenizers.get(name))
        if tok is not None:
            return len(tok)
    raise KeyError(f"no vocabulary size for attribute '{name}': pass additional_attributes_tokenizer or attribute_vocab_sizes")


class TransformerRecommenderModel(ArenaModule):
    """Common machinery; subclasses only describe where their parameters live."""

    # ---- filled in by subclasses -------------------------------------------------------------
    item_table_path: str
    pos_table_path: Optional[str] = None
    ln1_paths: Optional[Tuple[str, str]] = None
    ln2_paths: Optional[Tuple[str, str]] = None
    modifier_kind: str = "identity"            # "ffn" | "identity"
    projection_kind: str = "linear"            # "tied" | "linear" | "sasrec_neg"
    embed_dropout_a: bool = True               # dropout after LN1 (TransformerEmbedding.dropout)
    pre_attr_prefix: str = _PRE_ATTR           # module path of the pre-fused item-attribute tables
    blocks_path: str = BLOCKS                  # module path of the encoder blocks (state-dict prefix)

    def _setup(self, cfg: EncoderConfig, item_vocab_size: int, max_seq_length: int, specs, prefusion, postfusion,
               tokenizers, attr_sizes, merge: str, user_attributes=None, segment_rows: int = 0):
        self.cfg = cfg
        self.item_vocab_size = int(item_vocab_size)
        self.max_seq_length = int(max_seq_length)
        self.postfusion_merge_function = merge
        H = cfg.hidden
        self.prefusion: List[Tuple[str, str, int]] = []      # (name, type, vocab)
        self.postfusion: List[Tuple[str, str, int]] = []
        attr_specs = []
        # user attributes (UBERT4Rec / UserSASRec): their sum becomes the token at position 0.  Tables (``user_embedding`` /
        # ``content_embedding``: one id per user) or ``user_linear_upscale`` (models/ubert4rec/components.py:12-44: Linear over the
        # multi-hot of a LIST of ids, id 0 included -- unlike the item-side LinearUpscaler nothing is zeroed)
        self.user_attrs: List[Tuple[str, int, str]] = []
        for name, info in (user_attributes or {}).items():
            kind = info["embedding_type"]
            vu = _attr_vocab_size(name, tokenizers, attr_sizes)
            if kind in ("user_embedding", "content_embedding"):
                attr_specs.append((f"{_USER_ATTR}.{name}.weight", (vu, H)))
            elif kind == "user_linear_upscale":
                attr_specs.append((f"{_USER_ATTR}.{name}.linear.weight", (vu, H), "T"))   # stored transposed (Vu,H)
                attr_specs.append((f"{_USER_ATTR}.{name}.linear.bias", (H,)))
            else:
                raise NotImplementedError(f"user attribute embedding type {kind!r} (user_embedding, content_embedding and "
                                          "user_linear_upscale are built)")
            self.user_attrs.append((name, vu, kind))
        self.segment_rows = int(segment_rows)
        if self.segment_rows:
            attr_specs.append((_SEGMENT, (self.segment_rows, H)))
        for store, attrs, prefix in ((self.prefusion, prefusion, self.pre_attr_prefix), (self.postfusion, postfusion, _POST_ATTR)):
            for name, info in (attrs or {}).items():
                kind = info["embedding_type"]
                va = _attr_vocab_size(name, tokenizers, attr_sizes)
                store.append((name, kind, va))
                if kind == "content_embedding":
                    attr_specs.append((f"{prefix}.{name}.weight", (va, H)))
                elif kind == "linear_upscale":
                    attr_specs.append((f"{prefix}.{name}.linear.weight", (va, H), "T"))   # stored transposed (Va,H)
                    attr_specs.append((f"{prefix}.{name}.linear.bias", (H,)))
                else:
                    raise KeyError(f"{kind} invalid attribute embedding type")
        self.additional_userdata_keys = [n for n, _, _ in self.user_attrs]
        self.additional_metadata_keys = self.additional_userdata_keys + [n for n, _, _ in self.prefusion] + \
            [n for n, _, _ in self.postfusion]
        self._init_arena(list(specs) + attr_specs + [("_ghost_zero_row", (H,))])
        self.engine = EncoderEngine(self, cfg)
        self.precision = DEFAULT_PRECISION
        self.exact_topk = EXACT_TOPK
        self._wb_cache = None
        self._seed = 0
        self._step_counter = 0
        self.loss_scale = 1.0

    # ---- reference API --------------------------------------------------------------------------
    def required_metadata_keys(self) -> List[str]:
        return self.additional_metadata_keys

    def optional_metadata_keys(self) -> List[str]:
        return self.additional_userdata_keys

    replace_first_item: bool = False           # the user token takes the place of the first item instead of being prepended
    embedding_pooling_type: Optional[str] = None   # "max" | "sum" | "mean": basket inputs (N,S,BS) pooled per step

    @property
    def user_prefix(self) -> int:
        """1 when a user token is prepended to every sequence (hidden states then have S+1 positions); with
        ``replace_first_item`` (models/ubert4rec/components.py:117-121) the token overwrites position 0 and the length stays S"""
        return 1 if (self.user_attrs and not self.replace_first_item) else 0

    def _item_side(self, seq: torch.Tensor, attrs: Dict[str, torch.Tensor]):
        """what the embedding kernels see as the item sequence: with ``replace_first_item`` positions 1..S-1 (the embedding of
        item 0, its position and its attributes are computed by the reference and then thrown away)"""
        if not (self.user_attrs and self.replace_first_item):
            return seq, attrs
        sliced = dict(attrs)
        for name, _k, _v in self.prefusion:
            sliced[name] = attrs[name][:, 1:].contiguous()
        return seq[:, 1:].contiguous(), sliced

    # ---- embedding ------------------------------------------------------------------------------
    def _attr_operands(self, attrs: Dict[str, torch.Tensor], which, prefix, T):
        singles, bags = [], []
        for name, kind, _va in which:
            ids = attrs[name]
            if kind == "content_embedding":
                singles.append((ids.reshape(T), self.weight(f"{prefix}.{name}.weight")))
            else:
                width = ids.shape[-1] if ids.dim() == 3 else 1
                bags.append((ids.reshape(T, width), self.weight(f"{prefix}.{name}.linear.weight"),
                             self.weight(f"{prefix}.{name}.linear.bias")))
        return singles, bags

    def _embed_spec(self, seq: torch.Tensor, attrs, training: bool, seed: int) -> ops.EmbedSpec:
        """``seq`` / ``attrs``: the item side (:meth:`_item_side`)"""
        pool = None
        if seq.dim() == 3:       # basket inputs (N,S,BS): dense fallback -- pool first, then the fused kernel reads the pooled rows
            if not self.embedding_pooling_type:
                raise ValueError("basket sequences (N,S,BS) need embedding_pooling_type (max, sum or mean)")
            if self.user_attrs:
                raise NotImplementedError("basket sequences together with user attributes")
            B, S, BS = seq.shape
            basket = seq.reshape(B * S, BS)
            pooled, arg = ops.embed_pool_fwd(basket, self.weight(self.item_table_path), self.embedding_pooling_type)
            pool = (basket, arg, BS)
            seq = torch.arange(B * S, dtype=torch.int64, device=seq.device).view(B, S)
        B, S = seq.shape
        singles, bags = self._attr_operands(attrs, self.prefusion, self.pre_attr_prefix, B * S)
        # only column 0 of a user feature is read (models/ubert4rec/components.py:112-113).  A user_linear_upscale attribute with a
        # list of A ids per user becomes A gathers from the transposed Linear weight plus one from its bias (a one-row table)
        users, user_paths = [], []
        for name, _vu, kind in self.user_attrs:
            first = attrs[name][:, 0] if attrs[name].dim() > 1 else attrs[name]
            if kind == "user_linear_upscale":
                ids = first.reshape(B, -1)
                for a in range(ids.shape[1]):
                    users.append((ids[:, a].contiguous(), self.weight(f"{_USER_ATTR}.{name}.linear.weight")))
                    user_paths.append((f"{_USER_ATTR}.{name}.linear.weight", False))
                users.append((torch.zeros(B, dtype=torch.int64, device=seq.device), self.weight(f"{_USER_ATTR}.{name}.linear.bias").view(1, -1)))
                user_paths.append((f"{_USER_ATTR}.{name}.linear.bias", True))
            else:
                users.append((first.contiguous(), self.weight(f"{_USER_ATTR}.{name}.weight")))
                user_paths.append((f"{_USER_ATTR}.{name}.weight", False))
        seg = self.weight(_SEGMENT) if self.segment_rows else None
        ln1 = None if self.ln1_paths is None else (self.weight(self.ln1_paths[0]), self.weight(self.ln1_paths[1]))
        ln2 = None if self.ln2_paths is None else (self.weight(self.ln2_paths[0]), self.weight(self.ln2_paths[1]))
        pos = None if self.pos_table_path is None else self.weight(self.pos_table_path)
        if pos is not None and self.user_attrs and self.replace_first_item:
            pos = pos[1:]                       # item s of the sliced sequence sits at position s + 1 of the original one
        if pos is not None and S > pos.shape[0]:
            raise RuntimeError(f"sequence length {S} exceeds max_seq_length {pos.shape[0]}")
        spec = ops.EmbedSpec(seq.reshape(-1), pooled if pool is not None else self.weight(self.item_table_path), pos, singles, bags, ln1, ln2,
                             self.cfg.dropout if training else 0.0, seed, users=users, seg_table=seg)
        spec.user_paths = user_paths
        spec.pool = pool
        return spec

    def _embed_backward(self, saved: Saved, d_x: torch.Tensor):
        self.engine.wait_table_grad()           # a catalog gradient still running on the second stream adds into the same table
        spec: ops.EmbedSpec = saved.embed_spec
        B, S, H = saved.B, saved.S, self.cfg.hidden
        g = self._arena.ensure_grad()
        dln = None
        if self.ln1_paths is not None or self.ln2_paths is not None:
            off = self._arena.offsets[self._spec_name(self._dln_first)]     # [gamma1, beta1, gamma2, beta2] are glued
            dln = g[off:off + 4 * H].view(4, H)
        d_item, d_attr = ops.embed_bwd(spec, B, S, d_x, saved.embed_stats, dln)
        if not self.user_attrs:
            pre, ev = saved.extra.pop("sorted_items", (None, None))
            if getattr(spec, "pool", None) is not None:      # baskets: the token's gradient goes to the pooled slots' table rows
                basket, arg, BS = spec.pool
                d_slots = ops.embed_pool_bwd(d_item, arg, BS, self.embedding_pooling_type)
                ops.embgrad_sorted_reduce(basket.reshape(-1), d_slots, self.weight(self.item_table_path, g))
            elif pre is None:
                ops.embgrad_sorted_reduce(spec.item_ids, d_item, self.weight(self.item_table_path, g))
            else:
                if ev is not None:
                    torch.cuda.current_stream().wait_event(ev)
                pre.reduce(d_item, self.weight(self.item_table_path, g))
            if self.pos_table_path is not None:
                ops.posgrad_reduce(d_item, B, S, self.weight(self.pos_table_path, g))
            self._attr_backward(saved.extra["attrs"], self.prefusion, self.pre_attr_prefix, d_attr, B * S)
            return
        # user prefix: S counts the user position.  Rows of the user position carry the gradient of the user-embedding sum;
        # for the item-side tables they are skipped by giving them the id -1.
        Si = S - 1
        dev = d_item.device

        def with_user_column(ids, fill):
            ids = ids.reshape(B, Si, *ids.shape[2:]) if ids.dim() > 2 else ids.reshape(B, Si)
            col = torch.full((B, 1) + tuple(ids.shape[2:]), fill, dtype=torch.int64, device=dev)
            return torch.cat([col, ids], dim=1)

        ops.embgrad_sorted_reduce(with_user_column(spec.item_ids, -1).reshape(-1), d_item, self.weight(self.item_table_path, g))
        if self.pos_table_path is not None:
            dpos = self.weight(self.pos_table_path, g)
            ops.posgrad_reduce(d_item, B, Si, dpos[1:] if self.replace_first_item else dpos, prefix=1)
        attrs = saved.extra["attrs"]
        shifted = {name: with_user_column(attrs[name], -1) for name, _k, _v in self.prefusion}
        self._attr_backward(shifted, self.prefusion, self.pre_attr_prefix, d_attr, B * S)
        user_rows = torch.arange(B, device=dev, dtype=torch.int64) * S
        d_user = ops.gather_rows(d_attr, user_rows)
        for (ids, _tab), (path, is_bias) in zip(spec.users, spec.user_paths):
            dst = self.weight(path, g)
            ops.embgrad_sorted_reduce(ids, d_user, dst.view(1, -1) if is_bias else dst)
        if self.segment_rows:
            seg_ids = torch.ones(B, S, dtype=torch.int64, device=dev)
            seg_ids[:, 0] = 0
            ops.embgrad_sorted_reduce(seg_ids.reshape(-1), d_attr, self.weight(_SEGMENT, g))

    def _attr_backward(self, attrs, which, prefix, d_rows, T):
        g = self._arena.ensure_grad()
        for name, kind, _va in which:
            ids = attrs[name]
            if kind == "content_embedding":
                ops.embgrad_sorted_reduce(ids.reshape(T), d_rows, self.weight(f"{prefix}.{name}.weight", g))
            else:
                width = ids.shape[-1] if ids.dim() == 3 else 1
                ops.embgrad_sorted_reduce(ids.reshape(-1), d_rows, self.weight(f"{prefix}.{name}.linear.weight", g), skip_id=0,
                                          row_divisor=width)
                ops.colsum_accumulate(d_rows, self.weight(f"{prefix}.{name}.linear.bias", g))

    # ---- hidden states ----------------------------------------------------------------------------
    def _next_seed(self) -> int:
        self._step_counter += 1
        state = getattr(self, "_step_state", None)
        if state is not None:      # CUDA-graph mode: kernels re-read the seed from device memory on every replay
            return state.indirect_seed()
        return (int(self._seed) << 32) + self._step_counter

    def encode_rows(self, seq: torch.Tensor, padding_mask: Optional[torch.Tensor], attrs: Dict[str, torch.Tensor],
                    rows: torch.Tensor, one_per_sequence: bool = False) -> torch.Tensor:
        """evaluation: hidden states of the selected positions only, (len(rows), H).  Identical to ``encode(...)[rows]``; on the
        tensor-core path the last encoder layer skips the output projection / feed-forward of every other position."""
        if self.engine.use_tc() and not self.postfusion:
            if not self.arena_is_intact():
                self._repack()
            B, S = seq.shape[:2]
            S += self.user_prefix
            saved = Saved(B=B, S=S, seed=0, training=False, key_valid=self._key_valid(padding_mask, seq))
            x, _, y16, st = ops.embed_fwd(self._embed_spec(*self._item_side(seq, attrs), False, 0), B, S, next_ln=self.engine.first_norm())
            return self.engine.blocks_forward(x, saved, select_rows=rows, one_per_sequence=one_per_sequence, first_ln=(y16, st))
        hidden, _ = self.encode(seq, padding_mask, attrs, training=False)
        return ops.gather_rows(hidden, rows)

    def _key_valid(self, padding_mask: Optional[torch.Tensor], seq: torch.Tensor) -> Optional[torch.Tensor]:
        """key-validity bits of the encoder; the user position is never padding (models/ubert4rec/components.py:170-174)"""
        if not self.user_prefix or padding_mask is None:
            return padding_mask
        ones = torch.ones(seq.shape[0], 1, dtype=padding_mask.dtype, device=padding_mask.device)
        return torch.cat([ones, padding_mask], dim=1)

    def encode(self, seq: torch.Tensor, padding_mask: Optional[torch.Tensor], attrs: Dict[str, torch.Tensor],
               training: bool = False) -> Tuple[torch.Tensor, Saved]:
        """embed + encoder blocks (+ post-fusion merge): (T,H) hidden states of every position."""
        if not self.arena_is_intact():
            self._repack()
        B, S = seq.shape[:2]
        S += self.user_prefix          # hidden states cover the prepended user token as well
        saved = Saved(B=B, S=S, seed=self._next_seed() if training else 0, training=training,
                      key_valid=self._key_valid(padding_mask, seq))
        item_seq, attrs = self._item_side(seq, attrs)
        saved.extra["attrs"] = attrs
        spec = self._embed_spec(item_seq, attrs, training, saved.seed)
        saved.embed_spec = spec
        nxt = self.engine.first_norm()
        if nxt is not None:      # tensor-core path: the first block's input LayerNorm is computed by the embedding kernel
            x, saved.embed_stats, y16, st = ops.embed_fwd(spec, B, S, save_stats=training, next_ln=nxt, next_stats=training)
            x = self.engine.blocks_forward(x, saved, first_ln=(y16, st))
        else:
            x, saved.embed_stats = ops.embed_fwd(spec, B, S, save_stats=training)
            x = self.engine.blocks_forward(x, saved)
        if self.postfusion:
            if self.user_prefix:
                raise NotImplementedError("post-fusion attributes together with user attributes")
            singles, bags = self._attr_operands(attrs, self.postfusion, _POST_ATTR, B * S)
            zero_ids = torch.zeros(B * S, dtype=torch.int64, device=seq.device)
            ctx_spec = ops.EmbedSpec(zero_ids, self.weight("_ghost_zero_row").view(1, -1), None, singles, bags)
            ctx, _ = ops.embed_fwd(ctx_spec, B, S)
            saved.extra["post"] = (x, ctx)
            if self.postfusion_merge_function in ("add", "multiply"):
                x = ops.binary(x, ctx, self.postfusion_merge_function)
        return x, saved

    def encode_backward(self, d_hidden: torch.Tensor, saved: Saved):
        if self.postfusion and self.postfusion_merge_function in ("add", "multiply"):
            x, ctx = saved.extra["post"]
            if self.postfusion_merge_function == "add":
                d_ctx = d_hidden
            else:
                d_ctx = ops.binary(d_hidden, x, "multiply")
                d_hidden = ops.binary(d_hidden, ctx, "multiply")
            self._attr_backward(saved.extra["attrs"], self.postfusion, _POST_ATTR, d_ctx, saved.B * saved.S)
        # the sort of the (item id, token) pairs needs no gradient: second stream, long before the embedding backward consumes it
        if not self.user_attrs and d_hidden.is_cuda and getattr(saved.embed_spec, "pool", None) is None:
            spec = saved.embed_spec
            pre = ops.SortedIds(spec.item_ids, self.cfg.hidden, self.weight(self.item_table_path).shape[0])
            ev = self.engine.run_on_side(pre.sort, keep=(pre,))
            saved.extra["sorted_items"] = (pre, ev) if ev is not None else (pre.sort(), None)
        d_x = self.engine.blocks_backward(d_hidden, saved)
        self._embed_backward(saved, d_x)
        self.engine.join_side_stream()          # the weight gradients issued on the second stream are complete from here on

    # ---- projection -------------------------------------------------------------------------------
    def projection_operands(self, grad: bool = False):
        buf = self._arena.ensure_grad() if grad else None
        if self.projection_kind == "tied":
            return self.weight(self.item_table_path, buf), self.weight("_projection_layer.output_bias", buf)
        if self.projection_kind == "linear":
            return self.weight("_projection_layer.linear.weight", buf), self.weight("_projection_layer.linear.bias", buf)
        return self.weight(self.item_table_path, buf), None     # sasrec neg_sampling: h . E[item], no bias

    def projection_operands_bf16(self):
        """(catalog table as (V, Kp) bf16, bias fp32 or None) for the tensor-core scoring kernels.  With hidden % 64 == 0 the
        table is a view into the arena's bf16 shadow; otherwise a zero-padded copy cached until the weights change."""
        w, b = self.projection_operands()
        H = w.shape[1]
        if H % 64 == 0:
            path = {"tied": self.item_table_path, "linear": "_projection_layer.linear.weight"}.get(self.projection_kind, self.item_table_path)
            return self.weight_bf16(path), b
        stamp = (self._arena.version, self._arena.flat._version, w.data_ptr())
        if self._wb_cache is None or self._wb_cache[0] != stamp:
            self._wb_cache = (stamp, ops.cast_bf16(w))
        return self._wb_cache[1], b

    def projection_operands_folded(self):
        """evaluation operands with the output bias folded into the contraction: table (V, H+2 -> multiple of 16) bf16 =
        [w | bias_hi | bias_lo | 0..], cached until the weights change; hidden rows get [h | 1 | 1 | 0..] (ops.cast_bf16_ext).
        The scoring epilogue then has no per-column bias add (it cost ~0.3 ms per 1024 users x 1M items).  Without a bias the
        plain operands are returned."""
        w, b = self.projection_operands()
        if b is None:
            wb, _ = self.projection_operands_bf16()
            return wb, False
        stamp = (self._arena.version, self._arena.flat._version, w.data_ptr())
        cache = getattr(self, "_wb_folded", None)
        if cache is None or cache[0] != stamp:
            self._wb_folded = (stamp, ops.cast_bf16_ext(w, b))
        return self._wb_folded[1], True

    def projection_bias_bounds(self, v0: int = 0, v1: Optional[int] = None):
        """(bias slice as its own 16-byte-aligned tensor, per-32-item {max, min} of it) for the catalog slice [v0, v1) -- what the
        exact top-k sweep needs to keep the bias out of the contraction (ops.tc_score_candidates(bias_bounds=...)); None without a
        bias.  Cached until the weights change."""
        w, b = self.projection_operands()
        if b is None:
            return None
        v1 = b.numel() if v1 is None else v1
        stamp = (self._arena.version, self._arena.flat._version, b.data_ptr(), v0, v1)
        cache = getattr(self, "_bias_bounds", None)
        if cache is None or cache[0] != stamp:
            bs = b[v0:v1].clone() if (v0 != 0 or v1 != b.numel() or b.data_ptr() % 16) else b
            self._bias_bounds = (stamp, (bs, ops.bias_chunk_bounds(bs)))
        return self._bias_bounds[1]

    def projection_norm_bound(self):
        """(2): max row norm of the catalog table and max |bias| (ops.table_norm_bound), cached until the weights change"""
        w, b = self.projection_operands()
        stamp = (self._arena.version, self._arena.flat._version, w.data_ptr())
        cache = getattr(self, "_norm_bound", None)
        if cache is None or cache[0] != stamp:
            self._norm_bound = (stamp, ops.table_norm_bound(w, b))
        return self._norm_bound[1]

    def modify(self, rows: torch.Tensor, save: bool = False, n_live=None):
        if self.modifier_kind == "ffn":
            return self.engine.modifier_forward(rows, save, n_live)
        return rows, None

    def modify_backward(self, d_rows: torch.Tensor, saved_mod, n_live=None):
        if self.modifier_kind == "ffn":
            return self.engine.modifier_backward(d_rows, saved_mod, n_live)
        return d_rows

    # ---- reference-compatible forward (inference; materialises the logits the reference returns) --------
    @torch.no_grad()
    def forward(self, sequence: InputSequence):
        seq = sequence.sequence
        B, S = seq.shape[:2]           # (N,S) or basket inputs (N,S,BS) with embedding_pooling_type
        hidden, _ = self.encode(seq, sequence.padding_mask, sequence.attributes, training=False)
        if self.projection_kind == "sasrec_neg":
            return self._sasrec_forward(sequence, hidden)
        rows, _ = self.modify(hidden)
        w, b = self.projection_operands()
        return ops.gemm(rows, w, bias=b).view(B, S + self.user_prefix, self.item_vocab_size)

    def _sasrec_forward(self, sequence: InputSequence, hidden: torch.Tensor):
        pos = sequence.get_attribute("positive_samples")
        neg = sequence.get_attribute("negative_samples")
        B, S = sequence.sequence.shape
        table = self.weight(self.item_table_path)
        if neg is not None:
            pl, nl = ops.posneg_bce_fwd(hidden, table, pos.reshape(-1), neg.reshape(-1), None, None)
            return pl.view(B, S), nl.view(B, S)
        last = last_position_rows(sequence.sequence, sequence.padding_mask)
        h_last = ops.gather_rows(hidden, last)
        all_scores = ops.gemm(h_last, table)                    # (B,V)
        return all_scores if pos is None else torch.gather(all_scores, 1, pos)

    # ---- fused training: scoring + cross entropy --------------------------------------------------------
    def loss_ce(self, seq, padding_mask, attrs, target, pad_id: int = PAD_TOKEN_ID):
        """Cross entropy over the positions with target != pad (== nn.CrossEntropyLoss(ignore_index=pad) over all
        B*S rows, masked_training_module.py:107-111), without materialising logits.  Returns (loss, ctx).

        The positions are selected ON THE DEVICE (``ops.select_rows``): every buffer of the selected-rows path has the capacity
        B*S, the number of live rows stays in device memory and every kernel of the path reads it there.  Launch geometry and
        buffer sizes therefore depend on (B, S) only -- one CUDA graph serves every batch of a shape -- and the host never
        waits for the count (``torch.nonzero`` would)."""
        hidden, saved = self.encode(seq, padding_mask, attrs, training=self.training)
        rows, row_targets, n_dev = ops.select_rows(target, pad_id)
        plan_rows = self._plan_rows(n_dev)
        ops.live_rows_hint = plan_rows
        h_rows = ops.gather_rows(hidden, rows, n_dev)
        m_rows, saved_mod = self.modify(h_rows, save=self.training, n_live=n_dev)
        use_tc = self.precision == "bf16"
        if use_tc:      # tcgen05 scoring + CE partials on bf16 operands; the backward recomputes the same tiles
            wb, b = self.projection_operands_bf16()
            hb = ops.cast_bf16(m_rows, ld_out=wb.shape[1], n_live=n_dev)
            rmax, rsum, tl = ops.tc_score_ce_partial(hb, wb, b, row_targets, n_live=n_dev, plan_rows=plan_rows)
        else:
            hb = None
            w, b = self.projection_operands()
            rmax, rsum, tl = ops.score_ce_partial(m_rows, w, b, row_targets, n_live=n_dev)
        loss_acc = torch.zeros(2, dtype=torch.float32, device=seq.device)          # [sum, mean]
        lse = ops.ce_loss_from_partials(rmax, rsum, tl, loss_acc[0:1], n_dev, loss_acc[1:2])
        loss = loss_acc[1]                      # NaN when no position has a target, like the mean over an empty set in torch
        ctx = dict(saved=saved, rows=rows, row_targets=row_targets, m_rows=m_rows, hb=hb, saved_mod=saved_mod, lse=lse, n_dev=n_dev,
                   T=hidden.shape[0], plan_rows=plan_rows)
        return loss, ctx

    def _plan_rows(self, n_dev: torch.Tensor) -> int:
        """how many selected rows to PLAN the catalog sweeps for (their split over CTAs; never their result).  Read back once, on
        the first training step of the model -- the one host synchronisation of this path, outside any graph capture -- and kept:
        cloze masking and sequence lengths make the count vary by a few percent from batch to batch, not by factors."""
        hint = getattr(self, "_rows_hint", None)
        if hint is None and n_dev.is_cuda and not torch.cuda.is_current_stream_capturing():
            hint = self._rows_hint = max(int(n_dev.item()), 1)
        return hint or 0

    def loss_ce_backward(self, ctx, dloss: float = 1.0):
        g = self._prepare_grads()
        dw, db = self.projection_operands(grad=True)
        n_dev = ctx["n_dev"]
        if ctx["hb"] is not None:
            wb, b = self.projection_operands_bf16()
            args = (ctx["hb"], wb, b, ctx["row_targets"], ctx["lse"], dloss, self.cfg.hidden)      # the kernels divide by the live count
            # the catalog-gradient sweep (dW, dbias) is a leaf: second stream, next to the whole encoder backward
            plan = ctx["plan_rows"]
            if self.engine.run_on_side(lambda: ops.tc_score_ce_bwd(*args, dw, db, need_dh=False, slot=1, n_live=n_dev, plan_rows=plan),
                                       keep=args, table_grad=True) is not None:
                d_m = ops.tc_score_ce_bwd(*args, None, None, n_live=n_dev, plan_rows=plan)
            else:
                d_m = ops.tc_score_ce_bwd(*args, dw, db, n_live=n_dev, plan_rows=plan)
        else:
            w, b = self.projection_operands()
            d_m = ops.score_ce_bwd(ctx["m_rows"], w, b, ctx["row_targets"], ctx["lse"], dloss, dw, db, n_live=n_dev)
        d_h = self.modify_backward(d_m, ctx["saved_mod"], n_live=n_dev)
        d_hidden = torch.zeros(ctx["T"], self.cfg.hidden, dtype=torch.float32, device=d_h.device)
        ops.scatter_rows(d_h, ctx["rows"], d_hidden, n_dev)
        self.encode_backward(d_hidden, ctx["saved"])
        self.attach_grads()

    # ---- fused training: SASRec positive / negative BCE -----------------------------------------------
    def loss_bce(self, seq, padding_mask, attrs, pos, neg, mask):
        hidden, saved = self.encode(seq, padding_mask, attrs, training=self.training)
        table = self.weight(self.item_table_path)
        sums = torch.zeros(2, dtype=torch.float32, device=seq.device)
        pl, nl = ops.posneg_bce_fwd(hidden, table, pos.reshape(-1), neg.reshape(-1), mask.reshape(-1), sums)
        loss = sums[0] / sums[1]
        ctx = dict(saved=saved, hidden=hidden, pos=pos.reshape(-1), neg=neg.reshape(-1), mask=mask.reshape(-1), pl=pl, nl=nl,
                   sums=sums)
        return loss, ctx

    def loss_bce_backward(self, ctx, dloss: float = 1.0):
        g = self._prepare_grads()
        table = self.weight(self.item_table_path)
        d_hidden, d_pos, d_neg = ops.posneg_bce_bwd(ctx["hidden"], table, ctx["pos"], ctx["neg"], ctx["mask"], ctx["pl"],
                                                    ctx["nl"], ctx["sums"], dloss)
        dtable = self.weight(self.item_table_path, g)
        ops.embgrad_sorted_reduce(ctx["pos"], d_pos, dtable)
        ops.embgrad_sorted_reduce(ctx["neg"], d_neg, dtable)
        self.encode_backward(d_hidden, ctx["saved"])
        self.attach_grads()

    def _prepare_grads(self) -> torch.Tensor:
        """Zero the flat gradient buffer unless the caller is accumulating (param.grad already attached)."""
        g = self._arena.ensure_grad()
        first = next(iter(self._named_arena_params()))[1]
        if first.grad is None:
            ops.fill(g, 0.0)
        return g

    # ---- fused evaluation: scoring + top-k + exact target rank --------------------------------------------
    @torch.no_grad()
    def evaluate_rank(self, seq, padding_mask, attrs, target, k: int = 10, rows: Optional[torch.Tensor] = None,
                      select: str = "mask", mask_id: int = MASK_TOKEN_ID, with_loss: bool = False, pad_id: int = PAD_TOKEN_ID,
                      full_rank: bool = True, rows_one_per_sequence: bool = False):
        """returns dict(topk_val (B,k), topk_idx (B,k) int32, rank (B) int32 1-based, target_score (B)[, loss]).
        ``rows_one_per_sequence``: the caller guarantees rows[b] is a position of sequence b (as the built-in selectors
        produce them) -- the last encoder layer then computes only the attention query tile that holds it."""
        if rows is None:
            rows = mask_position_rows(seq, mask_id) if select == "mask" else last_position_rows(seq, padding_mask)
            rows_one_per_sequence = True
        h_rows = self.encode_rows(seq, padding_mask, attrs, rows, one_per_sequence=rows_one_per_sequence)
        m_rows, _ = self.modify(h_rows)
        out = self._evaluate_rows(m_rows, target, k, with_loss, pad_id, full_rank)
        # scores of chosen items for the same rows (fp32): item-subset samplers gather these instead of dense logits
        out["scorer"] = lambda items, rows_=m_rows: ops.score_items(rows_, *self.projection_operands(), items)
        return out

    @torch.no_grad()
    def recommend(self, seq, padding_mask, attrs, n: int, rows: Optional[torch.Tensor] = None, select: str = "mask",
                  mask_id: int = MASK_TOKEN_ID, rows_one_per_sequence: bool = False):
        """top-``n`` recommendation of the ``predict`` command without dense logits: dict(topk_idx (B,n) int32 best first,
        ties -> lowest id; topk_val (B,n) logits; lse (B) log-sum-exp over the whole catalog), i.e. everything
        ``softmax(logits).sort(descending=True)[:, :n]`` (evaluation/evaluation.py:176-178, :222-224) yields: the softmax score
        of an item is ``exp(logit - lse)``.  Two sweeps over the catalog (top-n, max / sum-exp); n <= 32."""
        if rows is None:
            rows = mask_position_rows(seq, mask_id) if select == "mask" else last_position_rows(seq, padding_mask)
            rows_one_per_sequence = True
        m_rows, _ = self.modify(self.encode_rows(seq, padding_mask, attrs, rows, one_per_sequence=rows_one_per_sequence))
        return self._recommend_rows(m_rows, n)

    def _recommend_rows(self, m_rows, n: int):
        if not 1 <= n <= 32:
            raise ValueError(f"asme_b200: the fused top-n list holds 1..32 entries, got {n}")
        zero = torch.zeros(m_rows.shape[0], dtype=torch.int64, device=m_rows.device)
        if self.precision == "bf16":
            wb, folded = self.projection_operands_folded()
            hb = ops.cast_bf16_ext(m_rows) if folded else ops.cast_bf16(m_rows, ld_out=wb.shape[1])
            if self.exact_topk:
                o = score_rows_tc_exact(m_rows, hb, wb, *self.projection_operands(), self.projection_norm_bound(), None, n, False)
            else:
                o = ops.tc_score_topk(hb, wb, None, n)
            rmax, rsum, _ = ops.tc_score_ce_partial(hb, wb, None, zero)
        else:
            w, b = self.projection_operands()
            val, idx, _, _ = ops.score_topk_rank(m_rows, w, b, n, zero, ops.score_targets(m_rows, w, b, zero))
            o = dict(topk_val=val, topk_idx=idx)
            rmax, rsum, _ = ops.score_ce_partial(m_rows, w, b, zero)
        return dict(topk_idx=o["topk_idx"], topk_val=o["topk_val"], lse=rmax + torch.log(rsum))

    def _evaluate_rows(self, m_rows, target, k, with_loss, pad_id, full_rank):
        if self.precision == "bf16":
            if BIAS_BOUNDS and self.exact_topk and not with_loss and not (full_rank and target is not None):
                # the usual evaluation call (@k metrics): the catalog table as it lies in the arena's bf16 shadow, K = hidden size, the
                # bias bounded per chunk (the folded operands below cost a ninth K step per tile and a second copy of the table)
                wb, _ = self.projection_operands_bf16()
                hb = ops.cast_bf16(m_rows, ld_out=wb.shape[1])
                return score_rows_tc_exact(m_rows, hb, wb, *self.projection_operands(), self.projection_norm_bound(), target, k, False,
                                           bias_bounds=self.projection_bias_bounds())
            wb, folded = self.projection_operands_folded()
            b = None
            hb = ops.cast_bf16_ext(m_rows) if folded else ops.cast_bf16(m_rows, ld_out=wb.shape[1])
            if self.exact_topk:
                out = score_rows_tc_exact(m_rows, hb, wb, *self.projection_operands(), self.projection_norm_bound(), target, k, full_rank)
            else:
                out = score_rows_tc(hb, wb, b, target, k, full_rank)
            if with_loss:
                rmax, rsum, tl = ops.tc_score_ce_partial(hb, wb, b, target)
                out["lse"] = rmax + torch.log(rsum)
                nll = out["lse"] - tl
                keep = target.ne(pad_id)
                out["loss"] = (nll * keep).sum() / keep.sum()
            return out
        w, b = self.projection_operands()
        out = score_rows(m_rows, w, b, target, k)
        if with_loss:      # nn.CrossEntropyLoss(ignore_index=pad) on the selected rows (masked_training_module.py:150)
            rmax, rsum, _tl = ops.score_ce_partial(m_rows, w, b, target)
            out["lse"] = rmax + torch.log(rsum)
            nll = out["lse"] - out["target_score"]
            keep = target.ne(pad_id)
            out["loss"] = (nll * keep).sum() / keep.sum()
        return out


def _evaluate_rank_sharded(self, seq, padding_mask, attrs, target, k: int = 10, rows: Optional[torch.Tensor] = None,
                           select: str = "mask", mask_id: int = MASK_TOKEN_ID, full_rank: bool = False, group=None,
                           with_loss: bool = False, pad_id: int = PAD_TOKEN_ID, rows_one_per_sequence: bool = False):
    """vocab-sharded variant of :meth:`evaluate_rank` (asme_b200.sharded): this rank encodes ITS users, one all-gather hands every
    rank all selected hidden rows, every rank scores all users against its own slice of the catalog on the tensor cores (exact
    lists, csrc/rescore.cu), one all-to-all brings every rank the slices' results for its own users.  Same keys as
    :meth:`evaluate_rank` incl. ``loss`` / ``lse`` (``with_loss``: the softmax statistics of the slices are merged too)."""
    import torch.distributed as dist
    from . import sharded
    one = rows is None or rows_one_per_sequence
    if rows is None:
        rows = mask_position_rows(seq, mask_id) if select == "mask" else last_position_rows(seq, padding_mask)
    m_rows, _ = self.modify(self.encode_rows(seq, padding_mask, attrs, rows, one_per_sequence=one))
    G = dist.get_world_size(group) if dist.is_initialized() else 1
    g = dist.get_rank(group) if dist.is_initialized() else 0
    plain = BIAS_BOUNDS and self.exact_topk and not full_rank and not with_loss and self.precision == "bf16"
    if plain:       # @k metrics only: the plain table slice, the bias bounded per chunk (see _evaluate_rows)
        wb, folded = self.projection_operands_bf16()[0], False
    else:
        wb, folded = self.projection_operands_folded()          # bias (if any) rides in two extra K columns of the table
    v0, v1 = sharded.shard_range(wb.shape[0], G, g)
    exact = None
    if self.exact_topk:
        w32, b32 = self.projection_operands()
        stamp = (self._arena.version, self._arena.flat._version, w32.data_ptr(), v0, v1)
        cache = getattr(self, "_shard_norm_bound", None)
        if cache is None or cache[0] != stamp:
            self._shard_norm_bound = (stamp, ops.table_norm_bound(w32[v0:v1], None if b32 is None else b32[v0:v1]))
        exact = (w32[v0:v1], None if b32 is None else b32[v0:v1], self._shard_norm_bound[1])
    bb = self.projection_bias_bounds(v0, v1) if plain else None
    scorer = sharded.tc_local_scorer(wb[v0:v1], bb[0] if bb is not None else None, v0, folded, exact=exact,
                                     bias_bounds=bb[1] if bb is not None else None)
    return sharded.sharded_topk_rank(m_rows, target, k, scorer, sharded.tc_merge,
                                     full_rank=full_rank, group=group, with_loss=with_loss, pad_id=pad_id, combine_ce=sharded.tc_combine_ce)


TransformerRecommenderModel.evaluate_rank_sharded = torch.no_grad()(_evaluate_rank_sharded)


# ASME_B200_BIAS_BOUNDS=1: the exact top-k sweep runs on the plain (V, H) table with the output bias bounded per 32-item chunk
# (ops.bias_chunk_bounds) instead of the table with the bias folded into 16 extra K columns.  Same lists (tests/test_gpu_exact_topk.py),
# but measured SLOWER on the 1024 x 1M x 128 call even for a tiny bias (0.38 vs 0.37 ms; no bias at all: 0.33 ms) and up to 0.54 ms
# when the bias varies as much as the scores do (every chunk then takes the exact-add path) -> off; the folded operands cost a
# constant 12 %.
BIAS_BOUNDS = os.environ.get("ASME_B200_BIAS_BOUNDS", "0") == "1"


def score_rows(m_rows, w, b, target, k):
    ts = ops.score_targets(m_rows, w, b, target)
    val, idx, ng, nt = ops.score_topk_rank(m_rows, w, b, k, target, ts)
    rank = (ng + nt + 1).to(torch.int32)
    return dict(topk_val=val, topk_idx=idx, rank=rank, target_score=ts, n_greater=ng, n_tie_lower=nt)


def score_rows_tc(hb, wb, b, target, k, full_rank: bool = True):
    """tensor-core scoring: one sweep gives the top-k list and the target's score; the exact full rank (needed only by the
    ``rank`` / full ``MRR`` metrics) costs a second, count-only sweep.  Without it the rank is the target's position in the
    list, or k+1 ("not in the top k"), which is all the @k metrics depend on."""
    o = ops.tc_score_topk(hb, wb, b, k, target=target)
    out = dict(topk_val=o["topk_val"], topk_idx=o["topk_idx"], target_score=o["target_score"])
    if full_rank:
        c = ops.tc_score_topk(hb, wb, b, 0, target=target, target_score_in=o["target_score"], capture_target=False)
        out["n_greater"], out["n_tie_lower"] = c["n_greater"], c["n_tie_lower"]
        out["rank"] = (c["n_greater"] + c["n_tie_lower"] + 1).to(torch.int32)
    else:
        hit = o["topk_idx"].eq(target.to(torch.int32).unsqueeze(1))
        pos = hit.to(torch.int32).argmax(dim=1).to(torch.int32)
        out["rank"] = torch.where(hit.any(dim=1), pos + 1, torch.full_like(pos, k + 1))
    return out


def score_rows_tc_exact(m_rows, hb, wb, w32, b32, norm_bound, target, k, full_rank: bool = False, bias_bounds=None):
    """exact top-k on the tensor-core path: bf16 sweep -> candidates + what the sweep may have dropped (``ops.tc_score_candidates``)
    -> fp32 re-score + order + certificate (``ops.topk_rescore``) -> exact fp32 sweep for the rows the certificate could not
    cover (``ops.score_topk_flagged``; normally none).  ``topk_idx`` / ``topk_val`` / ``target_score`` and the target's position
    among the top k are bit-identical to the fp32 path's for the same hidden rows; with ``full_rank`` a target OUTSIDE the top k
    gets the rank of the bf16 count sweep."""
    want_full = full_rank and target is not None
    if bias_bounds is not None:      # plain operands: (aligned bias, its chunk bounds) ride along instead of folded bias columns
        if want_full:
            raise ValueError("score_rows_tc_exact: the count sweep of full_rank needs the folded operands")
        o = ops.tc_score_candidates(hb, wb, bias_bounds[0], k, 64, bias_bounds=bias_bounds[1])
    else:
        o = ops.tc_score_candidates(hb, wb, None, k, 64, target=target if want_full else None)
    r = ops.topk_rescore(m_rows, w32, b32, o["cand_idx"], o["cand_val"], k, norm_bound, target, cand_bound=o["bound"])
    rank = r["rank"]
    out = dict(topk_val=r["topk_val"], topk_idx=r["topk_idx"], target_score=r["target_score"], n_uncertified=r["n_flagged"])
    if want_full:
        c = ops.tc_score_topk(hb, wb, None, 0, target=target, target_score_in=o["target_score"], capture_target=False)
        out["n_greater"], out["n_tie_lower"] = c["n_greater"], c["n_tie_lower"]
        full = (c["n_greater"] + c["n_tie_lower"] + 1).to(torch.int32)
        rank = torch.where(rank <= k, rank, torch.clamp(full, min=k + 1))
    ops.score_topk_flagged(m_rows, w32, b32, target, r["target_score"], k, r["row_flag"], r["topk_val"], r["topk_idx"], rank)
    out["rank"] = rank
    return out


def mask_position_rows(seq: torch.Tensor, mask_id: int) -> torch.Tensor:
    """flat row index b*S + (position of the MASK token) -- one MASK per row (masked_training_module.py:80-91)."""
    B, S = seq.shape
    pos = (seq == mask_id).to(torch.int32).argmax(dim=1)
    return torch.arange(B, device=seq.device, dtype=torch.int64) * S + pos


def last_position_rows(seq: torch.Tensor, padding_mask: Optional[torch.Tensor]) -> torch.Tensor:
    """flat row index of the last real item (next_item_prediction_training_module.py:226-244); python-style
    wrap-around (-1 -> S-1) for all-padding rows, as advanced indexing does in the reference."""
    B, S = seq.shape
    pm = padding_mask if padding_mask is not None else seq.ne(PAD_TOKEN_ID)
    last = pm.sum(dim=-1).to(torch.int64) - 1
    last = torch.where(last < 0, last + S, last)
    return torch.arange(B, device=seq.device, dtype=torch.int64) * S + last


# ----------------------------------------------------------------------------------------------------
# initialisers (a19)
# ----------------------------------------------------------------------------------------------------
def _is_norm(path: str) -> bool:
    return "norm" in path or path.endswith("transform.2.weight") or path.endswith("transform.2.bias")


def _is_matrix_weight(path: str, p: torch.Tensor) -> bool:
    return p.dim() == 2


@torch.no_grad()
def _init_normal(model: ArenaModule, std: float):
    """normal_initialize_weights (bert4rec_model.py:59-68): Linear/Embedding weights N(0,std), LN 1/0, biases 0;
    ``output_bias`` keeps U(+-1/sqrt(V)) (layers.py:134-136)."""
    for path, p in model._named_arena_params():
        if path.endswith("output_bias"):
            bound = 1.0 / math.sqrt(p.numel())
            p.uniform_(-bound, bound)
        elif _is_norm(path):
            p.fill_(1.0) if path.endswith("weight") else p.zero_()
        elif p.dim() == 2:
            p.copy_(torch.randn(p.shape) * std)
        else:
            p.zero_()


@torch.no_grad()
def _init_xavier(model: ArenaModule):
    """TransformerEncoderModel._init_weights (transformer_encoder_model.py:63-73): xavier-normal Linear+Embedding."""
    for path, p in model._named_arena_params():
        if _is_norm(path):
            p.fill_(1.0) if path.endswith("weight") else p.zero_()
        elif p.dim() == 2:
            fan_out, fan_in = p.shape
            p.copy_(torch.randn(p.shape) * math.sqrt(2.0 / (fan_in + fan_out)))
        else:
            p.zero_()


def _encoder_config(hidden, heads, layers, dropout, bidirectional, intermediate, attention_dropout) -> EncoderConfig:
    if hidden % heads != 0:
        raise AssertionError("transformer_hidden_size must be divisible by num_transformer_heads")
    return EncoderConfig(hidden=hidden, heads=heads, layers=layers,
                         intermediate=4 * hidden if intermediate is None else intermediate, bidirectional=bidirectional,
                         dropout=float(dropout), attention_dropout=float(dropout if attention_dropout is None else attention_dropout))


def _alias(model: nn.Module, alias_path: str, param: nn.Parameter):
    """Register the SAME Parameter under a second path (the reference's tied modules appear twice in state_dict)."""
    mod = model
    parts = alias_path.split(".")
    for part in parts[:-1]:
        if not hasattr(mod, part):
            setattr(mod, part, nn.Module())
        mod = getattr(mod, part)
    mod.register_parameter(parts[-1], param)


def _get_param(model: nn.Module, path: str) -> nn.Parameter:
    mod = model
    parts = path.split(".")
    for part in parts[:-1]:
        mod = getattr(mod, part)
    return mod._parameters[parts[-1]]


# ----------------------------------------------------------------------------------------------------
# the three models
# ----------------------------------------------------------------------------------------------------
class BERT4RecModel(TransformerRecommenderModel):
    """bidirectional encoder, NO positional embedding (quirk Q1), FFN modifier, tied projection by default."""

    item_table_path = f"{_EMB}.item_embedding.embedding.weight"
    ln1_paths = (f"{_EMB}.embedding_norm.weight", f"{_EMB}.embedding_norm.bias")
    modifier_kind = "ffn"
    _dln_first = f"{_EMB}.embedding_norm.weight"

    def __init__(self, transformer_hidden_size: int, num_transformer_heads: int, num_transformer_layers: int,
                 item_vocab_size: int, max_seq_length: int, transformer_dropout: float,
                 project_layer_type: str = "transpose_embedding", embedding_pooling_type: str = None,
                 initializer_range: float = 0.02, transformer_intermediate_size: int = None,
                 transformer_attention_dropout: float = None):
        super().__init__()
        item_vocab_size = resolve_vocab_size("item", item_vocab_size)     # InjectVocabularySize("item")
        if embedding_pooling_type:
            raise NotImplementedError("BERT4RecModel: the reference passes embedding_pooling_type in the place of TransformerEmbedding's "
                                      "positional_embedding flag (quirk Q1) -- it never pools; basket inputs work with KeBERT4RecModel / SASRecModel")
        H, V = transformer_hidden_size, item_vocab_size
        cfg = _encoder_config(H, num_transformer_heads, num_transformer_layers, transformer_dropout, True,
                              transformer_intermediate_size, transformer_attention_dropout)
        specs = [(self.item_table_path, (V, H)),
                 (f"{_EMB}.embedding_norm.weight+", (H,)), (f"{_EMB}.embedding_norm.bias+", (H,)), ("_ghost_ln2", (2 * H,))]
        specs += block_param_specs(cfg) + modifier_param_specs(H)
        if project_layer_type == "transpose_embedding":
            self.projection_kind = "tied"
            specs.append(("_projection_layer.output_bias", (V,)))
        elif project_layer_type == "linear":
            self.projection_kind = "linear"
            specs += [("_projection_layer.linear.weight", (V, H)), ("_projection_layer.linear.bias", (V,))]
        else:
            raise KeyError(f"{project_layer_type} invalid projection layer")
        self._setup(cfg, V, max_seq_length, specs, None, None, None, None, "add")
        if self.projection_kind == "tied":
            _alias(self, "_projection_layer.embedding.weight", _get_param(self, self.item_table_path))
        _init_normal(self, initializer_range)


class KeBERT4RecModel(TransformerRecommenderModel):
    """BERT4Rec + positional embedding + pre-/post-fused attribute embeddings; untied Linear projection."""

    item_table_path = f"{_EMB}.item_embedding_layer.item_embedding.embedding.weight"
    ln2_paths = (f"{_EMB}.norm_embedding.weight", f"{_EMB}.norm_embedding.bias")
    modifier_kind = "ffn"
    projection_kind = "linear"
    _dln_first = "_ghost_ln1"

    def __init__(self, transformer_hidden_size: int, num_transformer_heads: int, num_transformer_layers: int,
                 item_vocab_size: int, max_seq_length: int, transformer_dropout: float,
                 prefusion_attributes: Dict[str, Dict[str, Any]] = None, postfusion_attributes: Dict[str, Dict[str, Any]] = None,
                 additional_attributes_tokenizer: Dict[str, Any] = None, postfusion_merge_function: str = "add",
                 positional_embedding: bool = True, embedding_pooling_type: str = None, initializer_range: float = 0.02,
                 transformer_intermediate_size: Optional[int] = None, transformer_attention_dropout: Optional[float] = None,
                 attribute_vocab_sizes: Dict[str, int] = None):
        super().__init__()
        item_vocab_size = resolve_vocab_size("item", item_vocab_size)     # InjectVocabularySize("item")
        if embedding_pooling_type and embedding_pooling_type not in ("max", "sum", "mean"):
            raise KeyError(embedding_pooling_type)          # sequence_embedding.py:30-34
        self.embedding_pooling_type = embedding_pooling_type or None   # basket inputs (N,S,BS): dense fallback (_embed_spec)
        H, V = transformer_hidden_size, item_vocab_size
        cfg = _encoder_config(H, num_transformer_heads, num_transformer_layers, transformer_dropout, True,
                              transformer_intermediate_size, transformer_attention_dropout)
        # NOTE the reference passes positional_embedding positionally-correct here (keyword), but TransformerEmbedding's
        # default is True and the kwarg is not forwarded (kebert4rec_model.py:47-49): positions are always on.
        self.pos_table_path = f"{_EMB}.item_embedding_layer.position_embedding.weight"
        specs = [(self.item_table_path, (V, H)), (self.pos_table_path, (max_seq_length, H)),
                 ("_ghost_ln1+", (2 * H,)), (f"{_EMB}.norm_embedding.weight+", (H,)), (f"{_EMB}.norm_embedding.bias", (H,))]
        specs += block_param_specs(cfg) + modifier_param_specs(H)
        specs += [("_projection_layer.linear.weight", (V, H)), ("_projection_layer.linear.bias", (V,))]
        self._setup(cfg, V, max_seq_length, specs, prefusion_attributes, postfusion_attributes,
                    resolve_tokenizers(additional_attributes_tokenizer), attribute_vocab_sizes, postfusion_merge_function)
        _init_normal(self, initializer_range)


class SASRecModel(TransformerRecommenderModel):
    """causal encoder; LayerNorm (+dropout) applied TWICE to the embedding (quirk Q2); ``mode`` selects the
    projection: "neg_sampling" (dot with positive / negative item embeddings) or "full" (Linear over the catalog)."""

    item_table_path = f"{_EMB}.item_embedding_layer.item_embedding.embedding.weight"
    pos_table_path = f"{_EMB}.item_embedding_layer.position_embedding.weight"
    ln1_paths = (f"{_EMB}.item_embedding_layer.embedding_norm.weight", f"{_EMB}.item_embedding_layer.embedding_norm.bias")
    ln2_paths = (f"{_EMB}.norm_embedding.weight", f"{_EMB}.norm_embedding.bias")
    modifier_kind = "identity"
    _dln_first = f"{_EMB}.item_embedding_layer.embedding_norm.weight"

    def __init__(self, transformer_hidden_size: int, num_transformer_heads: int, num_transformer_layers: int,
                 item_vocab_size: int, max_seq_length: int, transformer_dropout: float,
                 prefusion_attributes: Dict[str, Dict[str, Any]] = None, postfusion_attributes: Dict[str, Dict[str, Any]] = None,
                 additional_attributes_tokenizer: Dict[str, Any] = None, postfusion_merge_function: str = "add",
                 embedding_pooling_type: str = None, transformer_intermediate_size: int = None,
                 transformer_attention_dropout: float = None, mode: str = "neg_sampling",
                 attribute_vocab_sizes: Dict[str, int] = None):
        super().__init__()
        item_vocab_size = resolve_vocab_size("item", item_vocab_size)     # InjectVocabularySize("item")
        if embedding_pooling_type and embedding_pooling_type not in ("max", "sum", "mean"):
            raise KeyError(embedding_pooling_type)          # sequence_embedding.py:30-34
        self.embedding_pooling_type = embedding_pooling_type or None   # basket inputs (N,S,BS): dense fallback (_embed_spec)
        H, V = transformer_hidden_size, item_vocab_size
        cfg = _encoder_config(H, num_transformer_heads, num_transformer_layers, transformer_dropout, False,
                              transformer_intermediate_size, transformer_attention_dropout)
        self.mode = mode
        specs = [(self.item_table_path, (V, H)), (self.pos_table_path, (max_seq_length, H)),
                 (self.ln1_paths[0] + "+", (H,)), (self.ln1_paths[1] + "+", (H,)),
                 (self.ln2_paths[0] + "+", (H,)), (self.ln2_paths[1], (H,))]
        specs += block_param_specs(cfg)
        if mode == "neg_sampling":
            self.projection_kind = "sasrec_neg"
        elif mode == "full":
            self.projection_kind = "linear"
            specs += [("_projection_layer.linear.weight", (V, H)), ("_projection_layer.linear.bias", (V,))]
        else:
            raise Exception(f"{mode} is an unknown projection mode. Choose either <full> or <neg_sampling>.")
        self._setup(cfg, V, max_seq_length, specs, prefusion_attributes, postfusion_attributes,
                    resolve_tokenizers(additional_attributes_tokenizer), attribute_vocab_sizes, postfusion_merge_function)
        if mode == "neg_sampling":      # SASRecProjectionComponent holds the TransformerEmbedding again (components.py:16-19)
            base = f"{_EMB}.item_embedding_layer"
            for sub in ("item_embedding.embedding.weight", "position_embedding.weight", "embedding_norm.weight",
                        "embedding_norm.bias"):
                _alias(self, f"_projection_layer.embedding.{sub}", _get_param(self, f"{base}.{sub}"))
        _init_xavier(self)


# ----------------------------------------------------------------------------------------------------
# user-attribute models: a user token built from user-attribute embeddings is prepended to the sequence
# ----------------------------------------------------------------------------------------------------
def _tokenizer_sizes(additional_tokenizers, attribute_vocab_sizes):
    """the reference receives ``InjectTokenizers()`` = {"tokenizers.<name>": tokenizer}; plain sizes may be given instead"""
    toks = {}
    for key, tok in (additional_tokenizers or {}).items():
        toks[key] = tok
        if key.startswith("tokenizers."):
            toks[key[len("tokenizers."):]] = tok
    return toks, attribute_vocab_sizes


class UBERT4RecModel(TransformerRecommenderModel):
    """models/ubert4rec/ubert4rec_model.py:19-92: item + position embeddings (no LayerNorm of their own) + item attributes,
    user token at position 0, optional segment embedding, LayerNorm + dropout, CAUSAL encoder (``bidirectional=False`` in the
    reference), FFN modifier, untied Linear projection.  Outputs have S+1 positions when user attributes are configured."""

    item_table_path = f"{_EMB}.item_embedding_layer.item_embedding.embedding.weight"
    ln2_paths = (f"{_EMB}.norm_embedding.weight", f"{_EMB}.norm_embedding.bias")
    modifier_kind = "ffn"
    projection_kind = "linear"
    pre_attr_prefix = _ADD_ATTR
    blocks_path = "_sequence_representation_layer.transformer_encoder.transformer_blocks"
    _dln_first = "_ghost_ln1"

    def __init__(self, transformer_hidden_size: int, num_transformer_heads: int, num_transformer_layers: int,
                 item_vocab_size: int, max_seq_length: int, transformer_dropout: float,
                 additional_attributes: Dict[str, Dict[str, Any]] = None, additional_tokenizers: Dict[str, Any] = None,
                 user_attributes: Dict[str, Dict[str, Any]] = None, positional_embedding: bool = True,
                 segment_embedding: bool = False, embedding_pooling_type: str = None, initializer_range: float = 0.02,
                 transformer_intermediate_size: Optional[int] = None, transformer_attention_dropout: Optional[float] = None,
                 attribute_vocab_sizes: Dict[str, int] = None):
        super().__init__()
        item_vocab_size = resolve_vocab_size("item", item_vocab_size)     # InjectVocabularySize("item")
        if embedding_pooling_type:
            raise NotImplementedError("basket pooling together with user attributes is not built")
        H, V = transformer_hidden_size, item_vocab_size
        if user_attributes:
            max_seq_length += 1                                   # ubert4rec_model.py:40-43
        cfg = _encoder_config(H, num_transformer_heads, num_transformer_layers, transformer_dropout, False,
                              transformer_intermediate_size, transformer_attention_dropout)
        specs = [(self.item_table_path, (V, H))]
        if positional_embedding:
            self.pos_table_path = f"{_EMB}.item_embedding_layer.position_embedding.weight"
            specs.append((self.pos_table_path, (max_seq_length, H)))
        specs += [("_ghost_ln1+", (2 * H,)), (self.ln2_paths[0] + "+", (H,)), (self.ln2_paths[1], (H,))]
        specs += block_param_specs(cfg, self.blocks_path) + modifier_param_specs(H)
        specs += [("_projection_layer.linear.weight", (V, H)), ("_projection_layer.linear.bias", (V,))]
        toks, sizes = _tokenizer_sizes(resolve_tokenizers(additional_tokenizers), attribute_vocab_sizes)
        # rows of the segment table as the reference counts them (components.py:75-90): one for the item attributes as a
        # whole plus one per user attribute; rows 0 (user) and 1 (items) are the ones ever read
        seg_rows = ((1 if additional_attributes else 0) + len(user_attributes or {})) if segment_embedding else 0
        if segment_embedding and (not user_attributes or seg_rows < 2):
            raise NotImplementedError("segment_embedding needs user attributes and at least two segment rows (the reference "
                                      "indexes row 1)")
        self._setup(cfg, V, max_seq_length, specs, additional_attributes, None, toks, sizes, "add",
                    user_attributes=user_attributes, segment_rows=seg_rows)
        _init_normal(self, initializer_range)


class UserSASRecModel(TransformerRecommenderModel):
    """models/user_sasrec/user_sasrec_model.py:21-107: SASRec whose sequence starts with a user token.  The item embedding is
    normed + dropped once by its TransformerEmbedding, the concatenated sequence once more (quirk Q2 applies to the item
    positions only); causal encoder, identity modifier; ``mode="full"`` scores the catalog with a Linear layer."""

    item_table_path = f"{_EMB}.item_embedding_layer.item_embedding.embedding.weight"
    ln1_paths = (f"{_EMB}.item_embedding_layer.embedding_norm.weight", f"{_EMB}.item_embedding_layer.embedding_norm.bias")
    ln2_paths = (f"{_EMB}.norm_embedding.weight", f"{_EMB}.norm_embedding.bias")
    modifier_kind = "identity"
    projection_kind = "linear"
    pre_attr_prefix = _ADD_ATTR
    blocks_path = "_sequence_representation_layer.transformer_encoder.transformer_blocks"
    _dln_first = f"{_EMB}.item_embedding_layer.embedding_norm.weight"

    def __init__(self, transformer_hidden_size: int, num_transformer_heads: int, num_transformer_layers: int,
                 item_vocab_size: int, max_seq_length: int, transformer_dropout: float,
                 additional_attributes: Dict[str, Dict[str, Any]] = None, additional_tokenizers: Dict[str, Any] = None,
                 user_attributes: Dict[str, Dict[str, Any]] = None, segment_embedding: bool = False,
                 embedding_pooling_type: str = None, transformer_intermediate_size: int = None,
                 transformer_attention_dropout: float = None, mode: str = "neg_sampling", positional_embedding: bool = True,
                 replace_first_item: bool = False, attribute_vocab_sizes: Dict[str, int] = None):
        super().__init__()
        item_vocab_size = resolve_vocab_size("item", item_vocab_size)     # InjectVocabularySize("item")
        if embedding_pooling_type:
            raise NotImplementedError("basket pooling together with user attributes is not built")
        if mode not in ("full", "neg_sampling"):
            raise Exception(f"{mode} is an unknown projection mode. Choose either <full> or <neg_sampling>.")
        if replace_first_item and segment_embedding and user_attributes:
            # components.py:123-128 builds S+1 segment ids for the S positions that are left: the reference fails on the shapes
            raise ValueError("replace_first_item=True cannot be combined with segment_embedding (the reference adds (N,S+1) segment "
                             "embeddings to (N,S) positions)")
        self.replace_first_item = bool(replace_first_item)
        if mode == "neg_sampling":
            # user_sasrec/components.py:10-60 == SASRecProjectionComponent: products with the positive / negative item embeddings
            # over the S positions of the item sequence -- only runnable when the encoder output has S positions too
            if user_attributes and not replace_first_item:
                raise ValueError("UserSASRecModel(mode='neg_sampling') multiplies (N,S,H) item embeddings with the (N,S+1,H) encoder "
                                 "output when a user token is PREPENDED (the reference fails on the shapes): use replace_first_item=True")
            self.projection_kind = "sasrec_neg"
        H, V = transformer_hidden_size, item_vocab_size
        if user_attributes:
            max_seq_length += 1
        cfg = _encoder_config(H, num_transformer_heads, num_transformer_layers, transformer_dropout, False,
                              transformer_intermediate_size, transformer_attention_dropout)
        self.mode = mode
        specs = [(self.item_table_path, (V, H))]
        if positional_embedding:
            self.pos_table_path = f"{_EMB}.item_embedding_layer.position_embedding.weight"
            specs.append((self.pos_table_path, (max_seq_length, H)))
        specs += [(self.ln1_paths[0] + "+", (H,)), (self.ln1_paths[1] + "+", (H,)),
                  (self.ln2_paths[0] + "+", (H,)), (self.ln2_paths[1], (H,))]
        specs += block_param_specs(cfg, self.blocks_path)
        if mode == "full":
            specs += [("_projection_layer.linear.weight", (V, H)), ("_projection_layer.linear.bias", (V,))]
        toks, sizes = _tokenizer_sizes(resolve_tokenizers(additional_tokenizers), attribute_vocab_sizes)
        seg_rows = ((1 if additional_attributes else 0) + len(user_attributes or {})) if segment_embedding else 0
        if segment_embedding and (not user_attributes or seg_rows < 2):
            raise NotImplementedError("segment_embedding needs user attributes and at least two segment rows")
        self._setup(cfg, V, max_seq_length, specs, additional_attributes, None, toks, sizes, "add",
                    user_attributes=user_attributes, segment_rows=seg_rows)
        if mode == "neg_sampling":      # UserSASRecProjectionComponent holds the TransformerEmbedding again (components.py:12-14)
            base = f"{_EMB}.item_embedding_layer"
            subs = ["item_embedding.embedding.weight", "embedding_norm.weight", "embedding_norm.bias"]
            if positional_embedding:
                subs.append("position_embedding.weight")
            for sub in subs:
                _alias(self, f"_projection_layer.embedding.{sub}", _get_param(self, f"{base}.{sub}"))
        _init_xavier(self)
