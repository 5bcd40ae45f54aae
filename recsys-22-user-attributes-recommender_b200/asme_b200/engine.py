"""Execution engine of the transformer-encoder recommenders on top of the C-ABI kernels.

``EncoderEngine`` owns no parameters: it reads weights (and writes gradients) through an
:class:`~asme_b200.arena.ArenaModule` using the reference's state-dict names, and strings the fused
kernels together:

  embed (K1-K4) -> L x [LN -> QKV GEMM -> attention (K5,K8) -> out-proj (+residual) -> LN -> FFN (+residual)]
  -> row select (K15) -> modifier (K11) -> fused scoring with CE / top-k+rank / BCE (K12-K20)

Every activation needed by the backward pass is saved explicitly (``Saved``); dropout masks are
recomputed from (seed, site, index).  Nothing here does arithmetic in torch.
"""
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch

import os

from . import ops
from ._lib import ACT_GELU, ACT_NONE

# weight gradients of the tensor-core backward on a second stream (ASME_B200_SIDE_WGRAD=0 keeps everything on one stream)
SIDE_STREAM_WGRAD = os.environ.get("ASME_B200_SIDE_WGRAD", "1") == "1"
# inference: feed-forward block as one kernel (asme_b200_tc_ffn_fused); ASME_B200_FUSED_FFN=0 keeps the two GEMM launches,
# ASME_B200_FUSED_FFN_LN=0 keeps the next block's LayerNorm as its own launch (bit-identical bf16 rows to the unfused path)
FUSED_FFN = os.environ.get("ASME_B200_FUSED_FFN", "1") == "1"
FUSED_FFN_LN = os.environ.get("ASME_B200_FUSED_FFN_LN", "1") == "1"
# inference: the output projection, its residual and the output sublayer's LayerNorm move into the same kernel (asme_b200_tc_block_tail_fused)
FUSED_TAIL = os.environ.get("ASME_B200_FUSED_TAIL", "1") == "1"

BLOCKS = "_sequence_representation_layer.transformer_layer.transformer_blocks"
MODIFIER = "_sequence_representation_modifier_layer"


@dataclass
class EncoderConfig:
    hidden: int
    heads: int
    layers: int
    intermediate: int
    bidirectional: bool
    dropout: float = 0.0
    attention_dropout: float = 0.0


@dataclass
class LayerSaved:
    x: torch.Tensor = None
    y1: torch.Tensor = None
    st1: torch.Tensor = None
    qkv: torch.Tensor = None
    ctx: torch.Tensor = None
    ast: torch.Tensor = None
    x2: torch.Tensor = None
    y2: torch.Tensor = None
    st2: torch.Tensor = None
    z: torch.Tensor = None
    a: torch.Tensor = None
    ctx16: torch.Tensor = None       # bf16 copies kept by the tensor-core path (GEMM operands of the backward pass)
    keep: torch.Tensor = None        # dropout keep bits of the attention probabilities (tensor-core attention)
    attn_tc: bool = False


@dataclass
class Saved:
    B: int = 0
    S: int = 0
    seed: int = 0
    training: bool = False
    key_valid: Optional[torch.Tensor] = None
    embed_spec: Any = None
    embed_stats: Optional[torch.Tensor] = None
    layers: List[LayerSaved] = field(default_factory=list)
    extra: Dict[str, Any] = field(default_factory=dict)


def block_param_specs(cfg: EncoderConfig, blocks: str = BLOCKS) -> List[Tuple[str, Tuple[int, ...]]]:
    """Arena layout of the encoder blocks: Wq|Wk|Wv and bq|bk|bv glued so that one GEMM computes QKV,
    (gamma, beta) pairs glued so that LayerNorm backward writes one (2,H) target."""
    H, FF = cfg.hidden, cfg.intermediate
    specs = []
    for l in range(cfg.layers):
        p = f"{blocks}.{l}"
        specs += [(f"{p}.attention.linear_layers.0.weight+", (H, H)), (f"{p}.attention.linear_layers.1.weight+", (H, H)),
                  (f"{p}.attention.linear_layers.2.weight", (H, H)),
                  (f"{p}.attention.linear_layers.0.bias+", (H,)), (f"{p}.attention.linear_layers.1.bias+", (H,)),
                  (f"{p}.attention.linear_layers.2.bias", (H,)),
                  (f"{p}.attention.output_linear.weight", (H, H)), (f"{p}.attention.output_linear.bias", (H,)),
                  (f"{p}.feed_forward.w_1.weight", (FF, H)), (f"{p}.feed_forward.w_1.bias", (FF,)),
                  (f"{p}.feed_forward.w_2.weight", (H, FF)), (f"{p}.feed_forward.w_2.bias", (H,)),
                  (f"{p}.input_sublayer.norm.weight+", (H,)), (f"{p}.input_sublayer.norm.bias", (H,)),
                  (f"{p}.output_sublayer.norm.weight+", (H,)), (f"{p}.output_sublayer.norm.bias", (H,))]
    return specs


def modifier_param_specs(H: int) -> List[Tuple[str, Tuple[int, ...]]]:
    return [(f"{MODIFIER}.transform.0.weight", (H, H)), (f"{MODIFIER}.transform.0.bias", (H,)),
            (f"{MODIFIER}.transform.2.weight+", (H,)), (f"{MODIFIER}.transform.2.bias", (H,))]


class EncoderEngine:
    def __init__(self, module, cfg: EncoderConfig):
        self.m = module        # ArenaModule: weight(path[, buf]) / weights_span(...)
        self.cfg = cfg
        self.blocks = getattr(module, "blocks_path", BLOCKS)
        self._side = None          # second stream for the weight gradients of the tensor-core backward
        self._side_keep = []       # their operands stay referenced until the streams have joined

    # The weight gradients (dW = dY^T X, dbias) are leaves of the backward pass: nothing downstream reads them before the
    # optimizer.  They run on a second stream -- in a captured step they become a parallel branch of the graph -- and fill the
    # SMs the latency-bound dX chain leaves idle.  Same kernels, same order of additions: results are bit-identical.
    def _wgrad(self, dy16, x16, dw, db):
        if not SIDE_STREAM_WGRAD:
            return ops.tc_wgrad(dy16, x16, dw, db)
        if self._side is None:
            self._side = torch.cuda.Stream(device=dy16.device)
        self._side.wait_stream(torch.cuda.current_stream())          # the operands were produced on the main stream
        with torch.cuda.stream(self._side):
            ops.tc_wgrad(dy16, x16, dw, db, slot=1)
        self._side_keep.append((dy16, x16))                          # no reuse of their memory before the join

    def run_on_side(self, fn, keep, table_grad: bool = False):
        """other leaves of the backward pass (the catalog-gradient sweep of the CE backward).  ``table_grad``: the work adds into the
        item table's gradient, which the embedding backward on the main stream also does -- :meth:`wait_table_grad` orders them."""
        if not SIDE_STREAM_WGRAD:
            return None         # caller runs the work in line; otherwise the event that marks its completion
        if self._side is None:
            self._side = torch.cuda.Stream(device=torch.cuda.current_device())
        self._side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._side):
            fn()
            done = torch.cuda.Event()
            done.record(self._side)
        if table_grad:
            self._table_event = done
        self._side_keep.append(keep)
        return done

    def _defer(self):
        """``defer`` argument of the LayerNorm backward calls: the (gamma, beta) reductions are leaves -> second stream"""
        return self.run_on_side if SIDE_STREAM_WGRAD else None

    def wait_table_grad(self):
        ev = getattr(self, "_table_event", None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
            self._table_event = None

    def join_side_stream(self):
        if self._side_keep:
            torch.cuda.current_stream().wait_stream(self._side)
            self._side_keep.clear()
        self._table_event = None

    # ---------------------------------------------------------------- helpers
    def _w(self, path, grad=False):
        return self.m.weight(path, self.m._arena.ensure_grad() if grad else None)

    def _site(self, layer, k):
        return ops.SITE_LAYER_BASE + 8 * layer + k

    # ---------------------------------------------------------------- precision policy
    def use_tc(self) -> bool:
        """tensor-core (tcgen05, bf16 operands / fp32 accumulation) encoder, when the model asks for it and the shapes fit
        the 64-column TMA boxes; otherwise the strict fp32 SIMT path"""
        cfg = self.cfg
        return getattr(self.m, "precision", "fp32") == "bf16" and cfg.hidden % 64 == 0 and cfg.intermediate % 64 == 0

    # ---------------------------------------------------------------- encoder blocks, tensor-core path
    def first_norm(self):
        """(gamma, beta) of the first block's input LayerNorm when the tensor-core path runs: the embedding kernel applies it
        in the same pass (``blocks_forward(first_ln=...)``); None on the fp32 path"""
        if not self.use_tc() or self.cfg.layers == 0:
            return None
        pre = f"{self.blocks}.0"
        return self._w(f"{pre}.input_sublayer.norm.weight"), self._w(f"{pre}.input_sublayer.norm.bias")

    def _blocks_forward_tc(self, x: torch.Tensor, saved: Saved, select_rows: Optional[torch.Tensor] = None,
                           one_per_sequence: bool = False, first_ln=None) -> torch.Tensor:
        """``select_rows`` (evaluation only): flat indices of the positions whose hidden state is needed.  Every layer but the
        last runs on all tokens; in the last layer only K and V depend on the other positions, so everything after the attention
        (output projection, LayerNorm, feed-forward: ~half of the layer) runs on the selected rows alone.  The selected rows are
        bit-identical to the full computation; with ``one_per_sequence`` the last attention is the fp32 matrix-vector kernel
        (asme_b200_attn_row_fwd) instead of the tensor-core tile, i.e. equal up to the bf16 rounding of the probabilities.
        Returns (len(select_rows), H)."""
        cfg, m = self.cfg, self.m
        H = cfg.hidden
        train = saved.training
        p = cfg.dropout if train else 0.0
        pa = cfg.attention_dropout if train else 0.0
        B, S = saved.B, saved.S
        # LayerNorm of the coming layer's input when it was already produced upstream (by the embedding kernel for layer 0, by the
        # previous layer's last GEMM epilogue when that fusion is on)
        y1_next, st1_next = first_ln if first_ln is not None else (None, None)
        for l in range(cfg.layers):
            pre = f"{self.blocks}.{l}"
            ls = LayerSaved()
            ls.x = x
            if y1_next is not None:
                ls.y1, ls.st1 = y1_next, st1_next
            else:
                ls.y1, _, ls.st1 = ops.layernorm_fwd_bf16(x, self._w(f"{pre}.input_sublayer.norm.weight"),
                                                          self._w(f"{pre}.input_sublayer.norm.bias"), save_stats=train)
            wqkv = m.weights_span_bf16(f"{pre}.attention.linear_layers.0.weight", f"{pre}.attention.linear_layers.2.weight", (3 * H, H))
            bqkv = m.weights_span(f"{pre}.attention.linear_layers.0.bias", f"{pre}.attention.linear_layers.2.bias", (3 * H,))
            ls.attn_tc = S <= 256 and (H // cfg.heads) in (16, 32, 64)
            last_selected = select_rows is not None and l == cfg.layers - 1 and not train
            ctx_rows = None
            if ls.attn_tc:      # tcgen05 attention straight on the bf16 QKV projection
                ls.qkv = ops.tc_gemm(ls.y1, wqkv, bias=bqkv, out_f32=False, out_bf16=True)["bf16"]
                if last_selected and one_per_sequence and select_rows.numel() == B and 256 % (H // 8) == 0:
                    # one position per sequence: two matrix-vector products per head (HBM-bound), straight into (B, H) rows
                    ctx_rows = ops.attn_row_fwd(ls.qkv, saved.key_valid, B, S, cfg.heads, not cfg.bidirectional, select_rows)
                elif last_selected and one_per_sequence and select_rows.numel() == B:
                    # hidden sizes the row kernel's thread layout does not cover: only the query tile holding the row is computed
                    ls.ctx16 = ops.tc_attn_fwd_rows(ls.qkv, saved.key_valid, B, S, cfg.heads, not cfg.bidirectional, select_rows)
                else:
                    ls.ctx16, ls.ast, ls.keep = ops.tc_attn_fwd(ls.qkv, saved.key_valid, B, S, cfg.heads, not cfg.bidirectional, pa,
                                                                saved.seed, self._site(l, 0), save_stats=train)
            else:               # long sequences / odd head sizes: fp32 SIMT attention between tensor-core GEMMs
                ls.qkv = ops.tc_gemm(ls.y1, wqkv, bias=bqkv)["f32"]
                ls.ctx, ls.ast = ops.attn_fwd(ls.qkv, saved.key_valid, B, S, cfg.heads, not cfg.bidirectional, pa, saved.seed,
                                              self._site(l, 0), save_stats=train)
                ls.ctx16 = ops.cast_bf16(ls.ctx, ld_out=H)
            if select_rows is not None and l == cfg.layers - 1 and not train:
                # row selection (index plumbing, no arithmetic)
                ls.ctx16 = ctx_rows if ctx_rows is not None else ls.ctx16.index_select(0, select_rows)
                x = x.index_select(0, select_rows)
            nxt = None
            if l + 1 < cfg.layers:          # the next layer's input LayerNorm rides in the block's last epilogue
                npre = f"{self.blocks}.{l + 1}"
                nxt = (self._w(f"{npre}.input_sublayer.norm.weight"), self._w(f"{npre}.input_sublayer.norm.bias"))
            if not train and FUSED_TAIL and FUSED_FFN and H in (64, 128) and cfg.intermediate % 64 == 0:
                # inference: output projection + residual + LayerNorm + feed-forward + residual (+ next LayerNorm) in ONE kernel
                r = ops.tc_block_tail_fused(ls.ctx16, m.weight_bf16(f"{pre}.attention.output_linear.weight"),
                                            self._w(f"{pre}.attention.output_linear.bias"), x,
                                            (self._w(f"{pre}.output_sublayer.norm.weight"), self._w(f"{pre}.output_sublayer.norm.bias")),
                                            m.weight_bf16(f"{pre}.feed_forward.w_1.weight"), self._w(f"{pre}.feed_forward.w_1.bias"),
                                            m.weight_bf16(f"{pre}.feed_forward.w_2.weight"), self._w(f"{pre}.feed_forward.w_2.bias"),
                                            ln=nxt if FUSED_FFN_LN else None)
                x = r["f32"]
                if nxt is not None and not FUSED_FFN_LN:
                    r["ln16"], _, _ = ops.layernorm_fwd_bf16(x, nxt[0], nxt[1], save_stats=False)
                y1_next, st1_next = (r["ln16"], None) if nxt is not None else (None, None)
                continue
            # output projection + residual, with the output sublayer's LayerNorm fused into the epilogue (H <= 128)
            r = ops.tc_gemm(ls.ctx16, m.weight_bf16(f"{pre}.attention.output_linear.weight"),
                            bias=self._w(f"{pre}.attention.output_linear.bias"), p_drop=p, seed=saved.seed,
                            site=self._site(l, 1), residual=x,
                            ln=(self._w(f"{pre}.output_sublayer.norm.weight"), self._w(f"{pre}.output_sublayer.norm.bias")), ln_stats=train)
            ls.x2, ls.y2, ls.st2 = r["f32"], r["ln16"], r["ln_st"]
            nxt = None
            if l + 1 < cfg.layers:          # the next layer's input LayerNorm rides in this GEMM's epilogue
                npre = f"{self.blocks}.{l + 1}"
                nxt = (self._w(f"{npre}.input_sublayer.norm.weight"), self._w(f"{npre}.input_sublayer.norm.bias"))
            if not train and FUSED_FFN and H in (64, 128) and cfg.intermediate % 64 == 0:
                # inference: the whole feed-forward block (+ the next block's LayerNorm) in one kernel, no (T, 4H) intermediate
                r = ops.tc_ffn_fused(ls.y2, m.weight_bf16(f"{pre}.feed_forward.w_1.weight"), self._w(f"{pre}.feed_forward.w_1.bias"),
                                     m.weight_bf16(f"{pre}.feed_forward.w_2.weight"), self._w(f"{pre}.feed_forward.w_2.bias"), ls.x2,
                                     ln=nxt if FUSED_FFN_LN else None)
                x = r["f32"]
                if nxt is not None and not FUSED_FFN_LN:
                    r["ln16"], _, _ = ops.layernorm_fwd_bf16(x, nxt[0], nxt[1], save_stats=False)
                y1_next, st1_next = (r["ln16"], None) if nxt is not None else (None, None)
                continue
            r = ops.tc_gemm(ls.y2, m.weight_bf16(f"{pre}.feed_forward.w_1.weight"), bias=self._w(f"{pre}.feed_forward.w_1.bias"),
                            act=ACT_GELU, p_drop=p, seed=saved.seed, site=self._site(l, 2), out_f32=False, out_bf16=True,
                            pre_act=train)
            ls.a, ls.z = r["bf16"], r["pre"]
            r = ops.tc_gemm(ls.a, m.weight_bf16(f"{pre}.feed_forward.w_2.weight"), bias=self._w(f"{pre}.feed_forward.w_2.bias"),
                            p_drop=p, seed=saved.seed, site=self._site(l, 3), residual=ls.x2,
                            post_site=self._site(l, 4) if p > 0 else 0, ln=nxt, ln_stats=train)
            x = r["f32"]
            y1_next, st1_next = (r["ln16"], r["ln_st"]) if nxt is not None else (None, None)
            if train:
                saved.layers.append(ls)
        return x

    def _blocks_backward_tc(self, dx: torch.Tensor, saved: Saved) -> torch.Tensor:
        cfg, m = self.cfg, self.m
        H = cfg.hidden
        p, pa = cfg.dropout, cfg.attention_dropout
        B, S = saved.B, saved.S
        g = m._arena.ensure_grad()
        pending = None      # (dx3, dy16) of the layer below, already produced by the LayerNorm backward above it
        for l in reversed(range(cfg.layers)):
            pre = f"{self.blocks}.{l}"
            ls = saved.layers[l]
            # ---- feed forward: out = drop4(x2 + drop3(W2 a + b2)), a = drop2(gelu(z)), z = W1 y2 + b1
            if pending is not None:
                dx3, dy16 = pending
            elif p > 0:
                dx3, dy16 = ops.dropout_cast(dx, p, saved.seed, self._site(l, 4), self._site(l, 3), want_f32=True)
            else:
                dx3, dy16 = dx, ops.cast_bf16(dx, ld_out=H)
            self._wgrad(dy16, ls.a, self._w(f"{pre}.feed_forward.w_2.weight", True), self._w(f"{pre}.feed_forward.w_2.bias", True))
            dz16 = ops.tc_gemm(dy16, m.weight_bf16(f"{pre}.feed_forward.w_2.weight"), b_is_kn=True, gelu_grad_of=ls.z, p_drop=p,
                               seed=saved.seed, site=self._site(l, 2), out_f32=False, out_bf16=True)["bf16"]
            self._wgrad(dz16, ls.y2, self._w(f"{pre}.feed_forward.w_1.weight", True), self._w(f"{pre}.feed_forward.w_1.bias", True))
            dy2 = ops.tc_gemm(dz16, m.weight_bf16(f"{pre}.feed_forward.w_1.weight"), b_is_kn=True)["f32"]
            dgb2 = m.weights_span(f"{pre}.output_sublayer.norm.weight", f"{pre}.output_sublayer.norm.bias", (2, H), g)
            # ---- attention: x2 = x + drop1(ctx Wo^T + bo); the LayerNorm backward also emits do16 = bf16(dx2 * mask1)
            dx2, do16 = ops.layernorm_bwd_drop(dy2, ls.x2, self._w(f"{pre}.output_sublayer.norm.weight"), ls.st2, dgb2, dx3,
                                               p, saved.seed, 0, self._site(l, 1), defer=self._defer())
            self._wgrad(do16, ls.ctx16, self._w(f"{pre}.attention.output_linear.weight", True),
                         self._w(f"{pre}.attention.output_linear.bias", True))
            if ls.attn_tc:
                dctx16 = ops.tc_gemm(do16, m.weight_bf16(f"{pre}.attention.output_linear.weight"), b_is_kn=True, out_f32=False,
                                     out_bf16=True)["bf16"]
                dqkv16 = ops.tc_attn_bwd(ls.qkv, saved.key_valid, B, S, cfg.heads, not cfg.bidirectional, ls.ctx16, dctx16, ls.ast,
                                         ls.keep, pa)
            else:
                dctx = ops.tc_gemm(do16, m.weight_bf16(f"{pre}.attention.output_linear.weight"), b_is_kn=True)["f32"]
                dqkv = ops.attn_bwd(ls.qkv, saved.key_valid, B, S, cfg.heads, not cfg.bidirectional, ls.ctx, dctx, ls.ast, pa,
                                    saved.seed, self._site(l, 0))
                dqkv16 = ops.cast_bf16(dqkv, ld_out=3 * H)
            wqkv = m.weights_span_bf16(f"{pre}.attention.linear_layers.0.weight", f"{pre}.attention.linear_layers.2.weight", (3 * H, H))
            dwqkv = m.weights_span(f"{pre}.attention.linear_layers.0.weight", f"{pre}.attention.linear_layers.2.weight", (3 * H, H), g)
            dbqkv = m.weights_span(f"{pre}.attention.linear_layers.0.bias", f"{pre}.attention.linear_layers.2.bias", (3 * H,), g)
            self._wgrad(dqkv16, ls.y1, dwqkv, dbqkv)
            dy1 = ops.tc_gemm(dqkv16, wqkv, b_is_kn=True)["f32"]
            dgb1 = m.weights_span(f"{pre}.input_sublayer.norm.weight", f"{pre}.input_sublayer.norm.bias", (2, H), g)
            if l > 0:       # the gradient entering the layer below: its two dropped copies come out of this LayerNorm backward
                dx, dy16n = ops.layernorm_bwd_drop(dy1, ls.x, self._w(f"{pre}.input_sublayer.norm.weight"), ls.st1, dgb1, dx2,
                                                   p, saved.seed, self._site(l - 1, 4), self._site(l - 1, 3), defer=self._defer())
                pending = (dx, dy16n)
            else:
                dx = ops.layernorm_bwd(dy1, ls.x, self._w(f"{pre}.input_sublayer.norm.weight"), ls.st1, dgb1, d_residual=dx2,
                                       defer=self._defer())
        return dx

    # ---------------------------------------------------------------- encoder blocks, strict fp32 path
    def blocks_forward(self, x: torch.Tensor, saved: Saved, select_rows: Optional[torch.Tensor] = None,
                       one_per_sequence: bool = False, first_ln=None) -> torch.Tensor:
        if self.use_tc():
            saved.extra["tc"] = True
            return self._blocks_forward_tc(x, saved, select_rows, one_per_sequence, first_ln)
        if select_rows is not None:
            raise RuntimeError("row-selective encoding is implemented by the tensor-core path only")
        cfg, m = self.cfg, self.m
        H = cfg.hidden
        train = saved.training
        p = cfg.dropout if train else 0.0
        pa = cfg.attention_dropout if train else 0.0
        B, S = saved.B, saved.S
        for l in range(cfg.layers):
            pre = f"{self.blocks}.{l}"
            ls = LayerSaved()
            ls.x = x
            ls.y1, ls.st1 = ops.layernorm_fwd(x, self._w(f"{pre}.input_sublayer.norm.weight"),
                                              self._w(f"{pre}.input_sublayer.norm.bias"), save_stats=train)
            wqkv = m.weights_span(f"{pre}.attention.linear_layers.0.weight", f"{pre}.attention.linear_layers.2.weight", (3 * H, H))
            bqkv = m.weights_span(f"{pre}.attention.linear_layers.0.bias", f"{pre}.attention.linear_layers.2.bias", (3 * H,))
            ls.qkv = ops.gemm(ls.y1, wqkv, bias=bqkv)
            ls.ctx, ls.ast = ops.attn_fwd(ls.qkv, saved.key_valid, B, S, cfg.heads, not cfg.bidirectional, pa, saved.seed,
                                          self._site(l, 0), save_stats=train)
            ls.x2 = ops.gemm(ls.ctx, self._w(f"{pre}.attention.output_linear.weight"),
                             bias=self._w(f"{pre}.attention.output_linear.bias"), p_drop=p, seed=saved.seed,
                             site=self._site(l, 1), residual=x)
            ls.y2, ls.st2 = ops.layernorm_fwd(ls.x2, self._w(f"{pre}.output_sublayer.norm.weight"),
                                              self._w(f"{pre}.output_sublayer.norm.bias"), save_stats=train)
            if train:
                ls.a, ls.z = ops.gemm(ls.y2, self._w(f"{pre}.feed_forward.w_1.weight"), bias=self._w(f"{pre}.feed_forward.w_1.bias"),
                                      act=ACT_GELU, pre_act_out=True, p_drop=p, seed=saved.seed, site=self._site(l, 2))
            else:
                ls.a = ops.gemm(ls.y2, self._w(f"{pre}.feed_forward.w_1.weight"), bias=self._w(f"{pre}.feed_forward.w_1.bias"),
                                act=ACT_GELU)
            x3 = ops.gemm(ls.a, self._w(f"{pre}.feed_forward.w_2.weight"), bias=self._w(f"{pre}.feed_forward.w_2.bias"),
                          p_drop=p, seed=saved.seed, site=self._site(l, 3), residual=ls.x2)
            x = ops.dropout(x3, p, saved.seed, self._site(l, 4)) if p > 0 else x3
            if train:
                saved.layers.append(ls)
        return x

    def blocks_backward(self, dx: torch.Tensor, saved: Saved) -> torch.Tensor:
        if saved.extra.get("tc"):
            return self._blocks_backward_tc(dx, saved)
        cfg, m = self.cfg, self.m
        H = cfg.hidden
        p, pa = cfg.dropout, cfg.attention_dropout
        B, S = saved.B, saved.S
        g = m._arena.ensure_grad()
        for l in reversed(range(cfg.layers)):
            pre = f"{self.blocks}.{l}"
            ls = saved.layers[l]
            dx3 = ops.dropout(dx, p, saved.seed, self._site(l, 4)) if p > 0 else dx
            # ---- feed forward: x3 = x2 + drop(W2 a + b2), a = drop(gelu(z)), z = W1 y2 + b1
            dy = ops.dropout(dx3, p, saved.seed, self._site(l, 3)) if p > 0 else dx3
            ops.gemm_wgrad(dy, ls.a, self._w(f"{pre}.feed_forward.w_2.weight", True), self._w(f"{pre}.feed_forward.w_2.bias", True))
            dz = ops.gemm(dy, self._w(f"{pre}.feed_forward.w_2.weight"), trans_b=False, mul_gelu_grad_of=ls.z, p_drop=p,
                          seed=saved.seed, site=self._site(l, 2))
            ops.gemm_wgrad(dz, ls.y2, self._w(f"{pre}.feed_forward.w_1.weight", True), self._w(f"{pre}.feed_forward.w_1.bias", True))
            dy2 = ops.gemm(dz, self._w(f"{pre}.feed_forward.w_1.weight"), trans_b=False)
            dgb2 = m.weights_span(f"{pre}.output_sublayer.norm.weight", f"{pre}.output_sublayer.norm.bias", (2, H), g)
            dx2 = ops.layernorm_bwd(dy2, ls.x2, self._w(f"{pre}.output_sublayer.norm.weight"), ls.st2, dgb2, d_residual=dx3)
            # ---- attention: x2 = x + drop(ctx Wo^T + bo)
            do = ops.dropout(dx2, p, saved.seed, self._site(l, 1)) if p > 0 else dx2
            ops.gemm_wgrad(do, ls.ctx, self._w(f"{pre}.attention.output_linear.weight", True),
                           self._w(f"{pre}.attention.output_linear.bias", True))
            dctx = ops.gemm(do, self._w(f"{pre}.attention.output_linear.weight"), trans_b=False)
            dqkv = ops.attn_bwd(ls.qkv, saved.key_valid, B, S, cfg.heads, not cfg.bidirectional, ls.ctx, dctx, ls.ast, pa,
                                saved.seed, self._site(l, 0))
            wqkv = m.weights_span(f"{pre}.attention.linear_layers.0.weight", f"{pre}.attention.linear_layers.2.weight", (3 * H, H))
            dwqkv = m.weights_span(f"{pre}.attention.linear_layers.0.weight", f"{pre}.attention.linear_layers.2.weight", (3 * H, H), g)
            dbqkv = m.weights_span(f"{pre}.attention.linear_layers.0.bias", f"{pre}.attention.linear_layers.2.bias", (3 * H,), g)
            ops.gemm_wgrad(dqkv, ls.y1, dwqkv, dbqkv)
            dy1 = ops.gemm(dqkv, wqkv, trans_b=False)
            dgb1 = m.weights_span(f"{pre}.input_sublayer.norm.weight", f"{pre}.input_sublayer.norm.bias", (2, H), g)
            dx = ops.layernorm_bwd(dy1, ls.x, self._w(f"{pre}.input_sublayer.norm.weight"), ls.st1, dgb1, d_residual=dx2)
        return dx

    # ---------------------------------------------------------------- FFN modifier  LN(GELU(Wx+b))  (K11)
    def modifier_forward(self, x: torch.Tensor, save: bool, n_live=None):
        """``n_live``: device count of live rows when ``x`` is a capacity-sized row selection (models.loss_ce)"""
        w, b = self._w(f"{MODIFIER}.transform.0.weight"), self._w(f"{MODIFIER}.transform.0.bias")
        if save:
            a, z = ops.gemm(x, w, bias=b, act=ACT_GELU, pre_act_out=True, m_live=n_live)
        else:
            a, z = ops.gemm(x, w, bias=b, act=ACT_GELU, m_live=n_live), None
        y, st = ops.layernorm_fwd(a, self._w(f"{MODIFIER}.transform.2.weight"), self._w(f"{MODIFIER}.transform.2.bias"),
                                  save_stats=save, n_live=n_live)
        return y, (x, z, a, st)

    def modifier_backward(self, dy: torch.Tensor, saved, n_live=None) -> torch.Tensor:
        x, z, a, st = saved
        H = self.cfg.hidden
        g = self.m._arena.ensure_grad()
        dgb = self.m.weights_span(f"{MODIFIER}.transform.2.weight", f"{MODIFIER}.transform.2.bias", (2, H), g)
        on_side = x.is_cuda and SIDE_STREAM_WGRAD
        da = ops.layernorm_bwd(dy, a, self._w(f"{MODIFIER}.transform.2.weight"), st, dgb, n_live=n_live,
                               defer=self.run_on_side if on_side else None)
        dz = ops.gelu_backward(da, z, n_live=n_live)
        dw, db = self._w(f"{MODIFIER}.transform.0.weight", True), self._w(f"{MODIFIER}.transform.0.bias", True)
        # the weight gradient is a leaf: second stream (its own scratch slot), off the dX chain
        if not on_side or self.run_on_side(lambda: ops.gemm_wgrad(dz, x, dw, db, m_live=n_live, slot=1), keep=(dz, x)) is None:
            ops.gemm_wgrad(dz, x, dw, db, m_live=n_live)
        return ops.gemm(dz, self._w(f"{MODIFIER}.transform.0.weight"), trans_b=False, m_live=n_live)
