"""Tensor-level wrappers around the C ABI (include/asme_b200.h).

PyTorch is used for device memory and streams only: every function takes CUDA tensors, hands their
``data_ptr()`` and the current stream to ``libasme_b200.so`` and returns freshly allocated output
tensors.  No arithmetic happens in torch on this path and there is no fallback -- a CPU tensor or a
missing library raises.
"""
import ctypes
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import EmbedDesc, GemmEpilogue

# dropout "sites": every dropout application in the model has its own counter-based stream (common.cuh: dropout_bits4)
SITE_EMBED_A, SITE_EMBED_B = 1, 2
SITE_LAYER_BASE = 16          # + 8 * layer + {0: attn probs, 1: attn out, 2: ffn inner, 3: ffn out, 4: block end}

live_rows_hint = 0       # rows the caller expects to be live in capacity-sized row selections: only used to annotate timed calls


def _note_rows(R: int, n_live) -> int:
    """row count for the timing annotation (bench.py's cost model): the live rows, not the capacity, when a device count is given"""
    return live_rows_hint if (n_live is not None and 0 < live_rows_hint < R) else R


_workspaces: Dict[Tuple[int, int], torch.Tensor] = {}
_retired: List[torch.Tensor] = []      # outgrown scratch buffers: captured CUDA graphs may still hold their addresses


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _f32(t: torch.Tensor, name: str = "tensor") -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"asme_b200: {name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"asme_b200: {name} must be float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _i64(t: torch.Tensor, name: str = "ids") -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"asme_b200: {name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != torch.int64:
        t = t.to(torch.int64)
    return t if t.is_contiguous() else t.contiguous()


def _u8(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype == torch.bool:
        t = t.contiguous().view(torch.uint8)
    elif t.dtype != torch.uint8:
        t = t.to(torch.uint8)
    return t if t.is_contiguous() else t.contiguous()


def workspace(nbytes: int, device: torch.device, slot: int = 0) -> torch.Tensor:
    """A cached, grow-only scratch buffer per (device, slot); kernels on one stream serialise on it."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), slot)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _retired.append(buf)       # never freed: a graph captured earlier replays kernels that point into it.  Sizes at least
        buf = torch.empty(max(int(nbytes), 1 << 20, 2 * (buf.numel() if buf is not None else 0)), dtype=torch.uint8, device=device)
        _workspaces[key] = buf         # double, so the retired buffers together stay smaller than the live one
    return buf


# ------------------------------------------------------------------------------------------------
# embedding
# ------------------------------------------------------------------------------------------------
class EmbedSpec:
    """Python-side description of one fused embedding call (mirrors ``asme_embed_desc``)."""

    def __init__(self, item_ids, item_table, pos_table=None, attrs: Sequence[Tuple[torch.Tensor, torch.Tensor]] = (),
                 bags: Sequence[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = (), ln1=None, ln2=None, p_drop=0.0,
                 seed=0, site_a=SITE_EMBED_A, site_b=SITE_EMBED_B, users: Sequence[Tuple[torch.Tensor, torch.Tensor]] = (),
                 seg_table=None):
        self.item_ids = _i64(item_ids)
        self.item_table = _f32(item_table, "item_table")
        self.pos_table = None if pos_table is None else _f32(pos_table, "pos_table")
        self.attrs = [(_i64(i), _f32(t, "attr_table")) for i, t in attrs]
        # bags: (ids (T,width), table_t (Va,H) = Linear.weight^T, bias (H))
        self.bags = [(_i64(i), _f32(t, "bag_table_t"), _f32(b, "bag_bias")) for i, t, b in bags]
        self.ln1 = None if ln1 is None else (_f32(ln1[0]), _f32(ln1[1]))
        self.ln2 = None if ln2 is None else (_f32(ln2[0]), _f32(ln2[1]))
        self.p_drop, self.seed, self.site_a, self.site_b = float(p_drop), int(seed), site_a, site_b
        # user prefix: (ids (B), table (Vu,H)) per user attribute; the output then has one more position per sequence
        self.users = [(_i64(i).reshape(-1), _f32(t, "user_table")) for i, t in users]
        self.seg_table = None if seg_table is None else _f32(seg_table, "seg_table")
        if self.seg_table is not None and not self.users:
            raise RuntimeError("asme_b200: a segment embedding needs user attributes (the reference indexes row 1 of it)")
        if len(self.users) > _lib.ASME_MAX_ATTR:
            raise RuntimeError("asme_b200: too many user-attribute tables")
        if len(self.attrs) > _lib.ASME_MAX_ATTR or len(self.bags) > _lib.ASME_MAX_ATTR:
            raise RuntimeError("asme_b200: too many attribute tables")

    def desc(self) -> EmbedDesc:
        d = EmbedDesc()
        d.item_ids, d.item_table = self.item_ids.data_ptr(), self.item_table.data_ptr()
        d.pos_table = None if self.pos_table is None else self.pos_table.data_ptr()
        d.n_attr = len(self.attrs)
        for k, (ids, tab) in enumerate(self.attrs):
            d.attr_ids[k], d.attr_table[k] = ids.data_ptr(), tab.data_ptr()
        d.n_bag = len(self.bags)
        for k, (ids, tab, bias) in enumerate(self.bags):
            d.bag_ids[k], d.bag_width[k] = ids.data_ptr(), ids.shape[-1]
            d.bag_table_t[k], d.bag_bias[k] = tab.data_ptr(), bias.data_ptr()
        if self.ln1 is not None:
            d.ln1_gamma, d.ln1_beta = self.ln1[0].data_ptr(), self.ln1[1].data_ptr()
        if self.ln2 is not None:
            d.ln2_gamma, d.ln2_beta = self.ln2[0].data_ptr(), self.ln2[1].data_ptr()
        d.p_drop, d.seed, d.site_a, d.site_b = self.p_drop, self.seed, self.site_a, self.site_b
        d.n_user = len(self.users)
        for k, (ids, tab) in enumerate(self.users):
            d.user_ids[k], d.user_table[k] = ids.data_ptr(), tab.data_ptr()
        d.seg_table = None if self.seg_table is None else self.seg_table.data_ptr()
        return d


POOL_MODES = {"sum": 0, "mean": 1, "max": 2}


def embed_pool_fwd(ids: torch.Tensor, table: torch.Tensor, mode: str):
    """basket inputs: ids (T,BS) int64 -> (pooled (T,H) fp32 = sum / mean / max over the basket of table[ids], arg (T,H) uint8 slot of
    the maximum or None)"""
    ids, table = _i64(ids), _f32(table, "table")
    T, BS = ids.shape
    H = table.shape[1]
    out = torch.empty(T, H, dtype=torch.float32, device=table.device)
    arg = torch.empty(T, H, dtype=torch.uint8, device=table.device) if mode == "max" else None
    _lib.call("asme_b200_embed_pool_fwd", _p(ids), _p(table), T, BS, H, POOL_MODES[mode], _p(out), _p(arg), _stream())
    return out, arg


def embed_pool_bwd(d_out: torch.Tensor, arg: Optional[torch.Tensor], BS: int, mode: str) -> torch.Tensor:
    """(T*BS, H) gradient rows of every (token, basket slot)"""
    d_out = _f32(d_out)
    T, H = d_out.shape
    d_rows = torch.empty(T * BS, H, dtype=torch.float32, device=d_out.device)
    _lib.call("asme_b200_embed_pool_bwd", _p(d_out), _p(arg), T, BS, H, POOL_MODES[mode], _p(d_rows), _stream())
    return d_rows


def embed_fwd(spec: EmbedSpec, B: int, S: int, save_stats: bool = False, next_ln=None, next_stats: bool = False):
    """S counts the user position when ``spec.users`` is not empty (item ids are then (B, S-1)).
    ``next_ln=(gamma, beta)``: also returns (y16, st) = the following LayerNorm of the output rows as bf16 and, with
    ``next_stats``, its (2,T) row statistics -- the first encoder block's input norm, computed in the same pass."""
    H = spec.item_table.shape[1]
    T = B * S
    out = torch.empty(T, H, dtype=torch.float32, device=spec.item_table.device)
    stats = torch.empty(4, T, dtype=torch.float32, device=out.device) if save_stats else None
    d = spec.desc()
    y16 = st = None
    if next_ln is not None:
        g, b = _f32(next_ln[0]), _f32(next_ln[1])
        y16 = torch.empty(T, H, dtype=torch.bfloat16, device=out.device)
        st = torch.empty(2, T, dtype=torch.float32, device=out.device) if next_stats else None
        d.next_gamma, d.next_beta, d.next_out = g.data_ptr(), b.data_ptr(), y16.data_ptr()
        d.next_stats = None if st is None else st.data_ptr()
    if _lib.timing is not None:
        _lib.note = f"T={T},H={H},tables={1 + (spec.pos_table is not None) + len(spec.attrs) + sum(b[0].shape[-1] for b in spec.bags)},ids={1 + len(spec.attrs) + sum(b[0].shape[-1] for b in spec.bags)},next={int(next_ln is not None)}"
    _lib.call("asme_b200_embed_fwd", ctypes.byref(d), T, S, H, _p(out), _p(stats), _stream())
    if next_ln is not None:
        return out, stats, y16, st
    return out, stats


def embed_bwd(spec: EmbedSpec, B: int, S: int, d_out: torch.Tensor, stats: Optional[torch.Tensor], dln: Optional[torch.Tensor]):
    """returns (d_item_rows, d_attr_rows); accumulates LayerNorm parameter grads into dln (4,H)."""
    H = spec.item_table.shape[1]
    T = B * S
    d_out = _f32(d_out)
    d_item = torch.empty(T, H, dtype=torch.float32, device=d_out.device)
    two = spec.ln1 is not None and (spec.attrs or spec.bags or spec.users)
    d_attr = torch.empty_like(d_item) if two else d_item
    ws_bytes = _lib.query("asme_b200_embed_bwd_workspace_bytes", T, H)
    ws = workspace(ws_bytes, d_out.device)
    d = spec.desc()
    if _lib.timing is not None:
        _lib.note = f"T={T},H={H},tables={1 + (spec.pos_table is not None) + len(spec.attrs) + sum(b[0].shape[-1] for b in spec.bags)},ids={1 + len(spec.attrs) + sum(b[0].shape[-1] for b in spec.bags)}"
    _lib.call("asme_b200_embed_bwd", ctypes.byref(d), T, S, H, _p(d_out), _p(stats), _p(d_item), _p(d_attr), _p(dln),
              _p(ws), ws.numel(), _stream())
    return d_item, d_attr


def embgrad_sorted_reduce(ids: torch.Tensor, d_rows: torch.Tensor, d_table: torch.Tensor, skip_id: int = -1,
                          row_divisor: int = 1):
    """d_table[ids[t]] += d_rows[t // row_divisor] (deterministic: sort -> segmented reduce)."""
    ids = _i64(ids).reshape(-1)
    d_rows = _f32(d_rows)
    T, H = ids.numel(), d_table.shape[1]
    ws_bytes = _lib.query("asme_b200_embgrad_workspace_bytes", T, H)
    ws = workspace(ws_bytes, d_rows.device)
    if _lib.timing is not None:
        _lib.note = f"T={T},H={H}"
    _lib.call("asme_b200_embgrad_sorted_reduce", _p(ids), T, _p(d_rows), int(row_divisor), H, _p(d_table), d_table.shape[0], skip_id,
              _p(ws), ws.numel(), _stream())


class SortedIds:
    """phase 1 of the embedding-table gradient: the (id, token) pairs of one id tensor sorted by id, in a scratch buffer of its own"""

    def __init__(self, ids: torch.Tensor, H: int, V: int, skip_id: int = -1, row_divisor: int = 1):
        self.ids = _i64(ids).reshape(-1)
        self.T, self.H, self.V = self.ids.numel(), H, V
        self.skip_id, self.row_divisor = skip_id, int(row_divisor)
        self.ws = torch.empty(max(_lib.query("asme_b200_embgrad_workspace_bytes", self.T, H), 256), dtype=torch.uint8, device=ids.device)

    def sort(self):
        """launches phase 1 on the current stream (may differ from the stream the object was created on)"""
        if _lib.timing is not None:
            _lib.note = f"T={self.T},H={self.H}"
        _lib.call("asme_b200_embgrad_sort", _p(self.ids), self.T, self.row_divisor, self.H, self.V, self.skip_id, _p(self.ws),
                  self.ws.numel(), _stream())
        return self

    def reduce(self, d_rows: torch.Tensor, d_table: torch.Tensor):
        """phase 2: d_table[id] += the rows of d_rows that carry that id (same result as embgrad_sorted_reduce)"""
        d_rows = _f32(d_rows)
        assert d_table.shape == (self.V, self.H)
        if _lib.timing is not None:
            _lib.note = f"T={self.T},H={self.H}"
        _lib.call("asme_b200_embgrad_reduce_sorted", self.T, _p(d_rows), self.H, _p(d_table), self.V, _p(self.ws), self.ws.numel(), _stream())


def posgrad_reduce(d_rows: torch.Tensor, B: int, S: int, d_pos: torch.Tensor, prefix: int = 0):
    """d_pos[s] += sum_b d_rows[b, prefix + s]; d_rows is (B*(S+prefix), H)"""
    d_rows = _f32(d_rows)
    H = d_rows.shape[-1]
    if _lib.timing is not None:
        _lib.note = f"B={B},S={S},H={H}"
    if prefix == 0:
        _lib.call("asme_b200_posgrad_reduce", _p(d_rows), B, S, H, _p(d_pos), _stream())
    else:
        _lib.call("asme_b200_posgrad_reduce_strided", d_rows.data_ptr() + prefix * H * 4, B, S, S + prefix, H, _p(d_pos), _stream())


def colsum_accumulate(x: torch.Tensor, out: torch.Tensor):
    x = _f32(x)
    M, N = x.shape
    ws_bytes = _lib.query("asme_b200_colsum_workspace_bytes", M, N)
    ws = workspace(ws_bytes, x.device)
    if _lib.timing is not None:
        _lib.note = f"M={M},N={N}"
    _lib.call("asme_b200_colsum_accumulate", _p(x), M, N, _p(out), _p(ws), ws.numel(), _stream())


# ------------------------------------------------------------------------------------------------
# LayerNorm
# ------------------------------------------------------------------------------------------------
def layernorm_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, save_stats: bool = False, n_live=None):
    """``n_live`` (here and below): int32 device scalar = how many leading rows are live when ``x`` is a capacity-sized row
    selection (:func:`select_rows`); the other rows are left untouched"""
    x = _f32(x)
    M, H = x.shape
    y = torch.empty_like(x)
    stats = torch.empty(2, M, dtype=torch.float32, device=x.device) if save_stats else None
    if _lib.timing is not None:
        _lib.note = f"M={_note_rows(M, n_live)},H={H}"
    _lib.call("asme_b200_layernorm_fwd", _p(x), _p(gamma), _p(beta), M, H, _p(y), _p(stats), _p(n_live), _stream())
    return y, stats


def layernorm_fwd_bf16(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, save_stats: bool = False, want_f32: bool = False):
    """LayerNorm whose result feeds a tensor-core GEMM: returns (y_bf16, y_f32 or None, stats or None)"""
    x = _f32(x)
    M, H = x.shape
    y16 = torch.empty(M, H, dtype=torch.bfloat16, device=x.device)
    y32 = torch.empty_like(x) if want_f32 else None
    stats = torch.empty(2, M, dtype=torch.float32, device=x.device) if save_stats else None
    if _lib.timing is not None:
        _lib.note = f"M={M},H={H},f32={int(want_f32)}"
    _lib.call("asme_b200_layernorm_fwd_bf16", _p(x), _p(gamma), _p(beta), M, H, _p(y32), _p(y16), _p(stats), _stream())
    return y16, y32, stats


def dropout_cast(x: torch.Tensor, p: float, seed: int, site_a: int, site_b: int, want_f32: bool):
    """(x * mask_a as fp32 or None, bf16(x * mask_a * mask_b)); site 0 = no mask"""
    x = _f32(x)
    y16 = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    y32 = torch.empty_like(x) if want_f32 else None
    if _lib.timing is not None:
        _lib.note = f"n={x.numel()},f32={int(want_f32)}"
    _lib.call("asme_b200_dropout_cast", _p(x), x.numel(), float(p), int(seed), int(site_a), int(site_b), _p(y32), _p(y16), _stream())
    return y32, y16


def _ln_partials(M: int, H: int, device, defer):
    """scratch of the (gamma, beta) partials: the shared workspace, or -- when the reduction is deferred to another stream -- a
    buffer of its own that lives until that reduction has run"""
    ws_bytes = _lib.query("asme_b200_layernorm_bwd_workspace_bytes", M, H)
    if defer is None:
        return workspace(ws_bytes, device)
    return torch.empty(ws_bytes, dtype=torch.uint8, device=device)


def _ln_finish(ws, M: int, H: int, dgb: torch.Tensor, defer):
    """``defer(fn, keep)`` (EncoderEngine.run_on_side) runs the reduction of the partials as a leaf on the second stream; when it
    declines (returns None) the reduction runs here"""
    if defer is None:
        return
    chunks = int(_lib.load().asme_b200_layernorm_bwd_chunks(M, H))

    def reduce():
        _lib.call("asme_b200_rows_reduce", _p(ws), chunks, 2 * H, _p(dgb), 1, _stream())
    if defer(reduce, (ws, dgb)) is None:
        reduce()


def layernorm_bwd(dy, x, gamma, stats, dgb: torch.Tensor, d_residual: Optional[torch.Tensor] = None, n_live=None, defer=None):
    """dx = d_residual + LN'(dy); accumulates (dgamma, dbeta) into dgb (2,H).  ``defer``: see :func:`_ln_finish`."""
    dy, x = _f32(dy), _f32(x)
    M, H = x.shape
    dx = torch.empty_like(x)
    ws = _ln_partials(M, H, x.device, defer)
    if _lib.timing is not None:
        _lib.note = f"M={_note_rows(M, n_live)},H={H},res={int(d_residual is not None)}"
    _lib.call("asme_b200_layernorm_bwd", _p(dy), _p(x), _p(gamma), _p(stats), M, H, _p(d_residual), _p(dx),
              _p(dgb) if defer is None else None, _p(ws), ws.numel(), _p(n_live), _stream())
    _ln_finish(ws, M, H, dgb, defer)
    return dx


def layernorm_bwd_drop(dy, x, gamma, stats, dgb: torch.Tensor, d_residual, p: float, seed: int, site_a: int, site_b: int, defer=None):
    """LayerNorm backward fused with the dropout_cast of the next backward stage:
    returns (dx = (d_residual + LN'(dy)) * mask_a as fp32, bf16(dx * mask_b)); site 0 = no mask"""
    dy, x = _f32(dy), _f32(x)
    M, H = x.shape
    dx = torch.empty_like(x)
    dx16 = torch.empty(M, H, dtype=torch.bfloat16, device=x.device)
    ws = _ln_partials(M, H, x.device, defer)
    if _lib.timing is not None:
        _lib.note = f"M={M},H={H},res={int(d_residual is not None)},drop=1"
    _lib.call("asme_b200_layernorm_bwd_drop", _p(dy), _p(x), _p(gamma), _p(stats), M, H, _p(d_residual), _p(dx),
              _p(dgb) if defer is None else None, _p(ws), ws.numel(), float(p), int(seed), int(site_a), int(site_b), _p(dx16), _stream())
    _ln_finish(ws, M, H, dgb, defer)
    return dx, dx16


# ------------------------------------------------------------------------------------------------
# dense layers
# ------------------------------------------------------------------------------------------------
def gemm(a: torch.Tensor, b: torch.Tensor, trans_b: bool = True, bias=None, act: int = 0, pre_act_out: bool = False,
         mul_gelu_grad_of=None, p_drop: float = 0.0, seed: int = 0, site: int = 0, residual=None, out=None, m_live=None):
    """C = epilogue(A @ B^T) (trans_b) or epilogue(A @ B).  Returns C or (C, pre_act)."""
    a, b = _f32(a, "A"), _f32(b, "B")
    M, K = a.shape
    N = b.shape[0] if trans_b else b.shape[1]
    assert (b.shape[1] if trans_b else b.shape[0]) == K, "gemm: inner dimensions differ"
    c = out if out is not None else torch.empty(M, N, dtype=torch.float32, device=a.device)
    pre = torch.empty_like(c) if pre_act_out else None
    epi = GemmEpilogue()
    epi.bias = None if bias is None else bias.data_ptr()
    epi.pre_act = None if pre is None else pre.data_ptr()
    epi.act = act
    epi.mul_gelu_grad_of = None if mul_gelu_grad_of is None else mul_gelu_grad_of.data_ptr()
    epi.p_drop, epi.seed, epi.site = float(p_drop), int(seed), int(site)
    epi.residual = None if residual is None else residual.data_ptr()
    epi.m_live = None if m_live is None else m_live.data_ptr()
    if _lib.timing is not None:
        _lib.note = f"M={_note_rows(M, m_live)},N={N},K={K},tb={int(trans_b)},res={int(residual is not None)},pre={int(pre is not None)},aux={int(mul_gelu_grad_of is not None)}"
    _lib.call("asme_b200_gemm", _p(a), _p(b), _p(c), M, N, K, 1 if trans_b else 0, ctypes.byref(epi), _stream())
    return (c, pre) if pre_act_out else c


def gemm_wgrad(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, dbias: Optional[torch.Tensor], accumulate: bool = True,
               m_live=None, slot: int = 0):
    """dW (+)= dY^T X, dbias (+)= colsum(dY).  ``slot``: scratch buffer (calls issued on different streams must not share one)"""
    dy, x = _f32(dy), _f32(x)
    M, N = dy.shape
    K = x.shape[1]
    ws_bytes = _lib.query("asme_b200_gemm_wgrad_workspace_bytes", M, N, K)
    ws = workspace(ws_bytes, x.device, slot)
    if _lib.timing is not None:
        _lib.note = f"M={_note_rows(M, m_live)},N={N},K={K}"
    _lib.call("asme_b200_gemm_wgrad", _p(dy), _p(x), M, N, K, _p(dw), _p(dbias), 1 if accumulate else 0, _p(ws),
              ws.numel(), _p(m_live), _stream())


def dropout(x: torch.Tensor, p: float, seed: int, site: int) -> torch.Tensor:
    x = _f32(x)
    y = torch.empty_like(x)
    if _lib.timing is not None:
        _lib.note = f"n={x.numel()}"
    _lib.call("asme_b200_dropout", _p(x), _p(y), x.numel(), float(p), int(seed), int(site), _stream())
    return y


def gelu_backward(dy: torch.Tensor, z: torch.Tensor, n_live=None) -> torch.Tensor:
    dy, z = _f32(dy), _f32(z)
    dz = torch.empty_like(dy)
    if _lib.timing is not None:
        _lib.note = f"n={_note_rows(dy.shape[0], n_live) * dy.shape[-1]}"
    _lib.call("asme_b200_gelu_bwd", _p(dy), _p(z), _p(dz), dy.numel(), dy.shape[-1], _p(n_live), _stream())
    return dz


def binary(a: torch.Tensor, b: torch.Tensor, op: str) -> torch.Tensor:
    a, b = _f32(a), _f32(b)
    y = torch.empty_like(a)
    _lib.call("asme_b200_binary", _p(a), _p(b), _p(y), a.numel(), {"add": 0, "multiply": 1}[op], _stream())
    return y


# ------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------
def attn_fwd(qkv: torch.Tensor, key_valid: Optional[torch.Tensor], B: int, S: int, heads: int, causal: bool,
             p_drop: float = 0.0, seed: int = 0, site: int = 0, save_stats: bool = False):
    qkv = _f32(qkv)
    H = qkv.shape[1] // 3
    d = H // heads
    kv = _u8(key_valid)
    ctx = torch.empty(B * S, H, dtype=torch.float32, device=qkv.device)
    stats = torch.empty(2, B * heads * S, dtype=torch.float32, device=qkv.device) if save_stats else None
    if _lib.timing is not None:
        _lib.note = f"T={B * S},H={H},S={S},heads={heads}"
    _lib.call("asme_b200_attn_fwd", _p(qkv), _p(kv), B, S, heads, d, 1 if causal else 0, float(p_drop), int(seed),
              int(site), _p(ctx), _p(stats), _stream())
    return ctx, stats


def attn_bwd(qkv, key_valid, B, S, heads, causal, ctx, d_ctx, stats, p_drop=0.0, seed=0, site=0):
    qkv, ctx, d_ctx = _f32(qkv), _f32(ctx), _f32(d_ctx)
    H = qkv.shape[1] // 3
    d = H // heads
    kv = _u8(key_valid)
    d_qkv = torch.empty_like(qkv)
    ws_bytes = _lib.query("asme_b200_attn_bwd_workspace_bytes", B, S, heads)
    ws = workspace(ws_bytes, qkv.device)
    if _lib.timing is not None:
        _lib.note = f"T={B * S},H={H},S={S},heads={heads}"
    _lib.call("asme_b200_attn_bwd", _p(qkv), _p(kv), B, S, heads, d, 1 if causal else 0, float(p_drop), int(seed),
              int(site), _p(ctx), _p(d_ctx), _p(stats), _p(d_qkv), _p(ws), ws.numel(), _stream())
    return d_qkv


# ------------------------------------------------------------------------------------------------
# catalog scoring
# ------------------------------------------------------------------------------------------------
def score_targets(h: torch.Tensor, w: torch.Tensor, bias, target: torch.Tensor, v0: int = 0, out=None) -> torch.Tensor:
    h, w = _f32(h), _f32(w)
    R, H = h.shape
    if out is None:
        out = torch.zeros(R, dtype=torch.float32, device=h.device)
    _lib.call("asme_b200_score_targets", _p(h), R, H, _p(w), _p(bias), v0, w.shape[0], _p(_i64(target)), _p(out), _stream())
    return out


def score_items(h: torch.Tensor, w: torch.Tensor, bias, items: torch.Tensor) -> torch.Tensor:
    """scores of chosen catalog items: out[n, i] = h[n] . w[items[n, i]] (+ bias[items[n, i]]), fp32 -- what a sampled metric needs
    instead of dense (N,V) logits (metrics/container/metrics_sampler.py:74-204 gather them from the dense tensor)"""
    items = _i64(items)
    N, I = items.shape
    rows = torch.arange(N, device=h.device, dtype=torch.int64).repeat_interleave(I)
    return score_targets(gather_rows(_f32(h), rows), w, bias, items.reshape(-1)).view(N, I)


def weighted_negatives(cdf: torch.Tensor, input_seq: torch.Tensor, targets: torch.Tensor, n_samples: int, seed: int) -> torch.Tensor:
    """(N, n_samples) distinct item ids per user drawn from the item weights (``cdf`` = float64 cumulative sum), never the user's
    target nor an item of the input sequence"""
    input_seq, targets = _i64(input_seq), _i64(targets)
    N, S = input_seq.shape
    out = torch.empty(N, n_samples, dtype=torch.int64, device=input_seq.device)
    failed = torch.zeros(1, dtype=torch.int32, device=input_seq.device)
    _lib.call("asme_b200_weighted_negatives", _p(cdf), cdf.numel(), _p(input_seq), S, _p(targets), N, int(n_samples), int(seed), _p(out),
              _p(failed), _stream())
    return out, failed


def score_topk_rank(h, w, bias, k: int, target=None, target_score=None, v0: int = 0):
    """returns (topk_val (R,k), topk_idx (R,k) int32, n_greater (R) int32, n_tie_lower (R) int32)"""
    h, w = _f32(h), _f32(w)
    R, H = h.shape
    Vloc = w.shape[0]
    dev = h.device
    val = torch.empty(R, k, dtype=torch.float32, device=dev)
    idx = torch.empty(R, k, dtype=torch.int32, device=dev)
    ng = torch.zeros(R, dtype=torch.int32, device=dev) if target is not None else None
    nt = torch.zeros(R, dtype=torch.int32, device=dev) if target is not None else None
    ws_bytes = _lib.query("asme_b200_score_topk_workspace_bytes", R, Vloc, k)
    ws = workspace(ws_bytes, dev)
    tgt = None if target is None else _i64(target)
    if _lib.timing is not None:
        _lib.note = f"R={R},V={Vloc},H={H},k={k}"
    _lib.call("asme_b200_score_topk_rank", _p(h), R, H, _p(w), _p(bias), v0, Vloc, _p(tgt), _p(target_score), k, _p(val),
              _p(idx), _p(ng), _p(nt), _p(ws), ws.numel(), _stream())
    return val, idx, ng, nt


# ------------------------------------------------------------------------------------------------
# catalog scoring on the tensor cores (tcgen05 / TMEM / TMA): bf16 operands, fp32 accumulation
# ------------------------------------------------------------------------------------------------
def padded_k(H: int) -> int:
    """hidden size zero-padded to the 64-element (128-byte) swizzle chunk the TMA / UMMA path works in"""
    return (H + 63) // 64 * 64


def cast_bf16(x: torch.Tensor, ld_out: Optional[int] = None, n_live=None) -> torch.Tensor:
    """(rows, cols) fp32 -> (rows, ld_out) bf16, round-to-nearest-even, zero padded; rows past ``n_live`` become zeros"""
    x = _f32(x)
    rows, cols = x.shape
    ld_out = padded_k(cols) if ld_out is None else ld_out
    y = torch.empty(rows, ld_out, dtype=torch.bfloat16, device=x.device)
    if _lib.timing is not None:
        _lib.note = f"rows={_note_rows(rows, n_live)},cols={cols},ld={ld_out}"
    _lib.call("asme_b200_cast_bf16", _p(x), _p(y), rows, cols, cols, ld_out, _p(n_live), _stream())
    return y


def cast_bf16_ext(x: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(rows, cols) fp32 -> (rows, Kp) bf16 with the two bias-folding columns appended (ones for activations when ``bias`` is
    None, bf16 hi/lo parts of ``bias`` for weights); Kp = cols + 2 rounded up to a multiple of 16"""
    x = _f32(x)
    rows, cols = x.shape
    ld_out = (cols + 2 + 15) // 16 * 16
    y = torch.empty(rows, ld_out, dtype=torch.bfloat16, device=x.device)
    if _lib.timing is not None:
        _lib.note = f"rows={rows},cols={cols},ld={ld_out}"
    _lib.call("asme_b200_cast_bf16_ext", _p(x), _p(bias), _p(y), rows, cols, cols, ld_out, 1 if bias is None else 2, _stream())
    return y


def _bf16(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda or t.dtype != torch.bfloat16 or not t.is_contiguous():
        raise RuntimeError(f"asme_b200: {name} must be a contiguous CUDA bfloat16 tensor")
    return t


FUSE_LAYERNORM = os.environ.get("ASME_B200_FUSE_LN", "0") == "1"     # tc_gemm(ln=...): fuse the LayerNorm into the GEMM epilogue.  Measured: no gain (1.135 vs 1.123 ms per C2 step, 1.99 vs 1.97 ms per C5 step: the epilogue is latency-bound and the separate LayerNorm runs near the HBM roofline) -> off, kept as an option


def tc_score_topk(hb: torch.Tensor, wb: torch.Tensor, bias, k: int, target=None, target_score_in=None, v0: int = 0,
                  capture_target: bool = True):
    """hb (R,Kp) / wb (Vloc,Kp) bf16.  returns dict(topk_val, topk_idx, target_score, n_greater, n_tie_lower)
    (entries are None when not requested)."""
    hb, wb = _bf16(hb, "hb"), _bf16(wb, "wb")
    R, Kp = hb.shape
    Vloc = wb.shape[0]
    dev = hb.device
    val = torch.empty(R, k, dtype=torch.float32, device=dev) if k > 0 else None
    idx = torch.empty(R, k, dtype=torch.int32, device=dev) if k > 0 else None
    tgt = None if target is None else _i64(target)
    ts_out = torch.zeros(R, dtype=torch.float32, device=dev) if (tgt is not None and capture_target) else None
    count = target_score_in is not None
    ng = torch.zeros(R, dtype=torch.int32, device=dev) if count else None
    nt = torch.zeros(R, dtype=torch.int32, device=dev) if count else None
    ws_bytes = _lib.query("asme_b200_tc_score_topk_workspace_bytes", R, Kp, Vloc, k)
    ws = workspace(ws_bytes, dev)
    if _lib.timing is not None:
        _lib.note = f"R={R},V={Vloc},H={Kp},k={k},count={int(count)}"
    _lib.call("asme_b200_tc_score_topk", _p(hb), R, Kp, _p(wb), _p(bias), v0, Vloc, _p(tgt), _p(target_score_in), k, _p(val),
              _p(idx), _p(ts_out), _p(ng), _p(nt), _p(ws), ws.numel(), _stream())
    return dict(topk_val=val, topk_idx=idx, target_score=ts_out, n_greater=ng, n_tie_lower=nt)


def tc_score_ce_partial(hb: torch.Tensor, wb: torch.Tensor, bias, target, v0: int = 0, n_live=None, plan_rows: int = 0):
    """per row (max, sumexp, target logit) over the catalog slice [v0, v0+Vloc), tensor-core path.  ``plan_rows``: the host's guess
    of ``n_live`` (0 = all rows) -- only the split of the catalog over CTAs is chosen from it"""
    hb, wb = _bf16(hb, "hb"), _bf16(wb, "wb")
    R, Kp = hb.shape
    Vloc = wb.shape[0]
    dev = hb.device
    rmax = torch.empty(R, dtype=torch.float32, device=dev)
    rsum = torch.empty_like(rmax)
    tl = torch.zeros_like(rmax)
    plan_rows = int(plan_rows) if n_live is not None else 0
    ws_bytes = _lib.query("asme_b200_tc_score_ce_workspace_bytes", R, Kp, Vloc, plan_rows)
    ws = workspace(ws_bytes, dev)
    if _lib.timing is not None:
        _lib.note = f"R={_note_rows(R, n_live)},V={Vloc},H={Kp}"
    _lib.call("asme_b200_tc_score_ce_partial", _p(hb), R, Kp, _p(wb), _p(bias), v0, Vloc, _p(_i64(target)), _p(rmax), _p(rsum),
              _p(tl), _p(ws), ws.numel(), _p(n_live), plan_rows, _stream())
    return rmax, rsum, tl


def tc_score_ce_bwd(hb, wb, bias, target, lse, scale: float, H: int, dW: Optional[torch.Tensor], dbias: Optional[torch.Tensor],
                    v0: int = 0, need_dh: bool = True, slot: int = 0, n_live=None, plan_rows: int = 0) -> Optional[torch.Tensor]:
    """tensor-core backward of scoring + CE: returns dH (R,H) fp32; accumulates into dW (Vloc,H) / dbias (Vloc).  With ``n_live``
    the kernels use ``scale / n_live`` (the mean over the live rows) and leave the other rows of dH untouched"""
    hb, wb = _bf16(hb, "hb"), _bf16(wb, "wb")
    R, Kp = hb.shape
    Vloc = wb.shape[0]
    dh = torch.empty(R, H, dtype=torch.float32, device=hb.device) if need_dh else None
    plan_rows = int(plan_rows) if n_live is not None else 0
    ws_bytes = _lib.query("asme_b200_tc_score_ce_bwd_workspace_bytes", R, H, Kp, Vloc, plan_rows)
    ws = workspace(ws_bytes, hb.device, slot)
    if _lib.timing is not None:
        _lib.note = f"R={_note_rows(R, n_live)},V={Vloc},H={Kp}"
    _lib.call("asme_b200_tc_score_ce_bwd", _p(hb), R, H, Kp, _p(wb), _p(bias), v0, Vloc, _p(_i64(target)), _p(lse), float(scale),
              _p(dh), _p(dW), _p(dbias), _p(ws), ws.numel(), _p(n_live), plan_rows, _stream())
    return dh


def tc_gemm(a: torch.Tensor, b: torch.Tensor, b_is_kn: bool = False, bias=None, act: int = 0, gelu_grad_of=None,
            p_drop: float = 0.0, seed: int = 0, site: int = 0, residual=None, out_f32: bool = True, out_bf16: bool = False,
            pre_act: bool = False, post_site: int = 0, bf16_into: Optional[torch.Tensor] = None, ln=None, ln_stats: bool = False):
    """tensor-core dense layer; a (M,K) bf16, b (N,K) [or (K,N) when b_is_kn] bf16.
    returns dict(f32=..., bf16=..., pre=...) with the requested outputs.  ``ln=(gamma, beta)``: additionally
    ``ln16`` = LayerNorm(f32 output) as bf16 and, with ``ln_stats``, ``ln_st`` (2,M) -- fused into the epilogue when one column
    tile owns whole rows (N <= 128), a separate LayerNorm launch otherwise."""
    a, b = _bf16(a, "a"), _bf16(b, "b")
    M, K = a.shape
    N = b.shape[1] if b_is_kn else b.shape[0]
    dev = a.device
    c32 = torch.empty(M, N, dtype=torch.float32, device=dev) if out_f32 else None
    c16 = bf16_into if bf16_into is not None else (torch.empty(M, N, dtype=torch.bfloat16, device=dev) if out_bf16 else None)
    ld16 = c16.stride(0) if c16 is not None else 0
    pre = torch.empty(M, N, dtype=torch.bfloat16, device=dev) if pre_act else None
    fuse_ln = ln is not None and out_f32 and N <= 128 and FUSE_LAYERNORM
    if _lib.timing is not None:
        _lib.note = f"M={M},N={N},K={K},kn={int(b_is_kn)},res={int(residual is not None)},f32={int(out_f32)},bf16={int(c16 is not None)},pre={int(pre_act)},aux={int(gelu_grad_of is not None)}" + (",ln=1" if fuse_ln else "")
    common = (_p(a), _p(b), M, N, K, 1 if b_is_kn else 0, _p(bias), int(act), _p(gelu_grad_of), float(p_drop),
              int(seed), int(site), int(post_site), _p(residual), _p(c32), _p(c16), int(ld16), _p(pre))
    out = dict(f32=c32, bf16=c16, pre=pre)
    if fuse_ln:
        out["ln16"] = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        out["ln_st"] = torch.empty(2, M, dtype=torch.float32, device=dev) if ln_stats else None
        _lib.call("asme_b200_tc_gemm_ln", *common, _p(_f32(ln[0])), _p(_f32(ln[1])), _p(out["ln16"]), _p(out["ln_st"]), _stream())
        return out
    _lib.call("asme_b200_tc_gemm", *common, _stream())
    if ln is not None:
        out["ln16"], _, out["ln_st"] = layernorm_fwd_bf16(c32, ln[0], ln[1], save_stats=ln_stats)
    return out


def tc_ffn_fused(y16: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor, residual: torch.Tensor,
                 ln=None, out_f32: bool = True):
    """feed-forward block in one kernel (inference): out = residual + W2 gelu(W1 y + b1) + b2; y16 (M,H) bf16, w1 (FF,H) / w2 (H,FF)
    bf16, residual (M,H) fp32.  returns dict(f32=out or None, ln16=LayerNorm(out; *ln) as bf16 or None).  The (M,FF) intermediate
    never reaches HBM; ``f32`` is bit-identical to the two-GEMM path."""
    y16, w1, w2 = _bf16(y16, "y"), _bf16(w1, "w1"), _bf16(w2, "w2")
    M, H = y16.shape
    FF = w1.shape[0]
    if tuple(w1.shape) != (FF, H) or tuple(w2.shape) != (H, FF) or tuple(residual.shape) != (M, H):
        raise ValueError(f"tc_ffn_fused: shapes y{tuple(y16.shape)} w1{tuple(w1.shape)} w2{tuple(w2.shape)} residual{tuple(residual.shape)}")
    dev = y16.device
    o32 = torch.empty(M, H, dtype=torch.float32, device=dev) if out_f32 else None
    ln16 = torch.empty(M, H, dtype=torch.bfloat16, device=dev) if ln is not None else None
    if _lib.timing is not None:
        _lib.note = f"M={M},H={H},FF={FF},f32={int(out_f32)},ln={int(ln is not None)}"
    _lib.call("asme_b200_tc_ffn_fused", _p(y16), _p(w1), _p(_f32(b1)), _p(w2), _p(_f32(b2)), _p(_f32(residual)), M, H, FF, _p(o32),
              _p(_f32(ln[0])) if ln is not None else None, _p(_f32(ln[1])) if ln is not None else None, _p(ln16), _stream())
    return dict(f32=o32, ln16=ln16)


def tc_block_tail_fused(ctx16: torch.Tensor, wo: torch.Tensor, bo: torch.Tensor, residual: torch.Tensor, pro_ln, w1: torch.Tensor,
                        b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor, ln=None, out_f32: bool = True):
    """tail of an encoder block in one kernel (inference): x2 = residual + ctx Wo^T + bo; y = LayerNorm(x2; *pro_ln);
    out = x2 + W2 gelu(W1 y + b1) + b2.  returns dict(f32=out or None, ln16=LayerNorm(out; *ln) as bf16 or None)"""
    ctx16, wo, w1, w2 = _bf16(ctx16, "ctx"), _bf16(wo, "wo"), _bf16(w1, "w1"), _bf16(w2, "w2")
    M, H = ctx16.shape
    FF = w1.shape[0]
    if tuple(wo.shape) != (H, H) or tuple(w1.shape) != (FF, H) or tuple(w2.shape) != (H, FF) or tuple(residual.shape) != (M, H):
        raise ValueError("tc_block_tail_fused: operand shapes do not match")
    dev = ctx16.device
    o32 = torch.empty(M, H, dtype=torch.float32, device=dev) if out_f32 else None
    ln16 = torch.empty(M, H, dtype=torch.bfloat16, device=dev) if ln is not None else None
    if _lib.timing is not None:
        _lib.note = f"M={M},H={H},FF={FF},f32={int(out_f32)},ln={int(ln is not None)},pro=1"
    _lib.call("asme_b200_tc_block_tail_fused", _p(ctx16), _p(wo), _p(_f32(bo)), _p(_f32(residual)), _p(_f32(pro_ln[0])), _p(_f32(pro_ln[1])),
              _p(w1), _p(_f32(b1)), _p(w2), _p(_f32(b2)), M, H, FF, _p(o32),
              _p(_f32(ln[0])) if ln is not None else None, _p(_f32(ln[1])) if ln is not None else None, _p(ln16), _stream())
    return dict(f32=o32, ln16=ln16)


def tc_wgrad(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, dbias: Optional[torch.Tensor], accumulate: bool = True,
             slot: int = 0):
    """dw (N,K) fp32 (+)= dy(M,N)^T x(M,K); dbias (N) (+)= colsum(dy); dy, x bf16.  ``slot``: scratch buffer to use (calls issued
    on different streams must not share one)"""
    dy, x = _bf16(dy, "dy"), _bf16(x, "x")
    M, N = dy.shape
    K = x.shape[1]
    ws_bytes = _lib.query("asme_b200_tc_wgrad_workspace_bytes", M, N, K)
    ws = workspace(ws_bytes, x.device, slot)
    if _lib.timing is not None:
        _lib.note = f"M={M},N={N},K={K}"
    _lib.call("asme_b200_tc_wgrad", _p(dy), _p(x), M, N, K, _p(dw), _p(dbias), 1 if accumulate else 0, _p(ws), ws.numel(), _stream())


def tc_attn_fwd(qkv16: torch.Tensor, key_valid: Optional[torch.Tensor], B: int, S: int, heads: int, causal: bool,
                p_drop: float = 0.0, seed: int = 0, site: int = 0, save_stats: bool = False):
    """tensor-core attention: qkv16 (B*S, 3H) bf16 -> (ctx (B*S, H) bf16, stats (2, B*heads*S) or None,
    keep_bits (B*heads*S, 8) int32 or None [dropout keep bits, saved for the backward pass])"""
    qkv16 = _bf16(qkv16, "qkv")
    H = qkv16.shape[1] // 3
    d = H // heads
    kv = _u8(key_valid)
    ctx = torch.empty(B * S, H, dtype=torch.bfloat16, device=qkv16.device)
    stats = torch.empty(2, B * heads * S, dtype=torch.float32, device=qkv16.device) if save_stats else None
    keep = torch.zeros(B * heads * S, 8, dtype=torch.int32, device=qkv16.device) if (save_stats and p_drop > 0) else None
    if _lib.timing is not None:
        _lib.note = f"T={B * S},H={H},S={S},heads={heads}"
    _lib.call("asme_b200_tc_attn_fwd", _p(qkv16), _p(kv), B, S, heads, d, 1 if causal else 0, float(p_drop), int(seed), int(site),
              _p(ctx), _p(stats), _p(keep), _stream())
    return ctx, stats, keep


def tc_attn_fwd_rows(qkv16: torch.Tensor, key_valid: Optional[torch.Tensor], B: int, S: int, heads: int, causal: bool,
                     only_row: torch.Tensor) -> torch.Tensor:
    """evaluation of selected positions: ctx (B*S, H) bf16 of which only the 128-query tile holding flat row ``only_row[b]`` of
    every sequence is computed (the other rows are uninitialised)"""
    qkv16 = _bf16(qkv16, "qkv")
    H = qkv16.shape[1] // 3
    ctx = torch.empty(B * S, H, dtype=torch.bfloat16, device=qkv16.device)
    if _lib.timing is not None:
        _lib.note = f"T={B * S},H={H},S={S},heads={heads},rows=1"
    _lib.call("asme_b200_tc_attn_fwd_rows", _p(qkv16), _p(_u8(key_valid)), B, S, heads, H // heads, 1 if causal else 0,
              _p(_i64(only_row)), _p(ctx), _stream())
    return ctx


def attn_row_fwd(qkv16: torch.Tensor, key_valid: Optional[torch.Tensor], B: int, S: int, heads: int, causal: bool,
                 only_row: torch.Tensor) -> torch.Tensor:
    """evaluation of one position per sequence: (B, H) bf16 attention output of flat rows ``only_row`` (two matrix-vector
    products per head over the sequence's K / V rows; fp32 arithmetic)"""
    qkv16 = _bf16(qkv16, "qkv")
    H = qkv16.shape[1] // 3
    out = torch.empty(B, H, dtype=torch.bfloat16, device=qkv16.device)
    if _lib.timing is not None:
        _lib.note = f"T={B * S},H={H},S={S},heads={heads}"
    _lib.call("asme_b200_attn_row_fwd", _p(qkv16), _p(_u8(key_valid)), B, S, heads, H // heads, 1 if causal else 0,
              _p(_i64(only_row)), _p(out), _stream())
    return out


def tc_attn_bwd(qkv16, key_valid, B, S, heads, causal, ctx16, d_ctx16, stats, keep_bits, p_drop=0.0):
    """d_qkv (B*S, 3H) bf16"""
    qkv16, ctx16, d_ctx16 = _bf16(qkv16, "qkv"), _bf16(ctx16, "ctx"), _bf16(d_ctx16, "d_ctx")
    H = qkv16.shape[1] // 3
    d = H // heads
    kv = _u8(key_valid)
    d_qkv = torch.empty_like(qkv16)
    if _lib.timing is not None:
        _lib.note = f"T={B * S},H={H},S={S},heads={heads}"
    _lib.call("asme_b200_tc_attn_bwd", _p(qkv16), _p(kv), B, S, heads, d, 1 if causal else 0, float(p_drop), _p(ctx16),
              _p(d_ctx16), _p(stats), _p(keep_bits), _p(d_qkv), _stream())
    return d_qkv


def table_norm_bound(w: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """(3) fp32: max row L2 norm of the catalog table, max |bias|, max row norm of (table - bf16(table)) -- the constants of the
    exact top-k certificate"""
    w = _f32(w, "table")
    out = torch.empty(3, dtype=torch.float32, device=w.device)
    _lib.call("asme_b200_table_norm_bound", _p(w), w.shape[0], w.shape[1], _p(bias), _p(out), _stream())
    return out


def bias_chunk_bounds(bias: torch.Tensor) -> torch.Tensor:
    """(ceil(V/32), 2) fp32: {max, min} of ``bias`` over every 32-item chunk -- lets the top-k sweeps skip the bias add for chunks
    that cannot reach a row's threshold (``tc_score_candidates(bias_bounds=...)``)"""
    bias = _f32(bias, "bias")
    V = bias.numel()
    out = torch.empty((V + 31) // 32, 2, dtype=torch.float32, device=bias.device)
    _lib.call("asme_b200_bias_chunk_bounds", _p(bias), V, _p(out), _stream())
    return out


def tc_score_candidates(hb: torch.Tensor, wb: torch.Tensor, bias, k: int, k_out: int = 64, target=None, v0: int = 0, bias_bounds=None):
    """candidates for an exact top-``k``: dict(cand_val, cand_idx (R,k_out) best first by bf16 score, bound (R) = upper bound of the bf16
    score of every item in none of the sweep's lists, target_score (R) bf16 score of the target or None).  ``bias`` / ``bias_bounds``
    (:func:`bias_chunk_bounds` of the same slice): the bias stays out of the contraction and is added per chunk only where needed."""
    hb, wb = _bf16(hb, "hb"), _bf16(wb, "wb")
    R, Kp = hb.shape
    Vloc = wb.shape[0]
    dev = hb.device
    k_out = max(k, min(int(k_out), 64))
    val = torch.empty(R, k_out, dtype=torch.float32, device=dev)
    idx = torch.empty(R, k_out, dtype=torch.int32, device=dev)
    bound = torch.empty(R, dtype=torch.float32, device=dev)
    tgt = None if target is None else _i64(target)
    ts = torch.zeros(R, dtype=torch.float32, device=dev) if tgt is not None else None
    ws = workspace(_lib.query("asme_b200_tc_score_candidates_workspace_bytes", R, Kp, Vloc, k, k_out), dev)
    if _lib.timing is not None:
        _lib.note = f"R={R},V={Vloc},H={Kp},k={k},kout={k_out}"
    if bias_bounds is not None and (bias is None or tuple(bias_bounds.shape) != ((Vloc + 31) // 32, 2)):
        raise ValueError("tc_score_candidates: bias_bounds must be bias_chunk_bounds(bias) of the same catalog slice")
    _lib.call("asme_b200_tc_score_candidates", _p(hb), R, Kp, _p(wb), _p(bias), _p(bias_bounds), v0, Vloc, _p(tgt), k, k_out, _p(val), _p(idx), _p(bound),
              _p(ts), _p(ws), ws.numel(), _stream())
    return dict(cand_val=val, cand_idx=idx, bound=bound, target_score=ts)


def topk_rescore(h: torch.Tensor, w: torch.Tensor, bias, cand_idx: torch.Tensor, cand_val: torch.Tensor, k: int,
                 norm_bound: torch.Tensor, target=None, v0: int = 0, want_rank: bool = True, cand_bound=None):
    """exact top-k from the bf16 sweep's candidates: re-score in fp32 (the fp32 path's arithmetic), order (score desc, id asc),
    certify.  returns dict(topk_val (R,k), topk_idx (R,k), target_score (R) or None, rank (R) or None [position among the exact top
    k, k+1 otherwise], row_flag (R) int32 [1 = not certified: run :func:`score_topk_flagged`], n_flagged (1) int32)"""
    h, w = _f32(h, "hidden rows"), _f32(w, "table")
    R, H = h.shape
    KC = cand_idx.shape[1]
    dev = h.device
    val = torch.empty(R, k, dtype=torch.float32, device=dev)
    idx = torch.empty(R, k, dtype=torch.int32, device=dev)
    tgt = None if target is None else _i64(target)
    ts = torch.zeros(R, dtype=torch.float32, device=dev) if tgt is not None else None
    rank = torch.empty(R, dtype=torch.int32, device=dev) if (tgt is not None and want_rank) else None
    flag = torch.empty(R, dtype=torch.int32, device=dev)
    n_flagged = torch.empty(1, dtype=torch.int32, device=dev)
    if _lib.timing is not None:
        _lib.note = f"R={R},H={H},KC={KC},k={k}"
    _lib.call("asme_b200_topk_rescore", _p(h), R, H, _p(w), _p(bias), v0, w.shape[0], _p(cand_idx.contiguous()), _p(_f32(cand_val)),
              _p(cand_bound), KC, k,
              _p(norm_bound), _p(tgt), _p(val), _p(idx), _p(ts), _p(rank), _p(flag), _p(n_flagged), _stream())
    return dict(topk_val=val, topk_idx=idx, target_score=ts, rank=rank, row_flag=flag, n_flagged=n_flagged)


def score_topk_flagged(h, w, bias, target, target_score, k: int, row_flag, topk_val, topk_idx, rank=None, v0: int = 0):
    """the exact fp32 sweep for the rows with ``row_flag != 0`` only; overwrites their rows of topk_val / topk_idx (/ rank = exact
    full rank).  With no flagged row this is a few microseconds of empty CTAs -- decided on the device, no host round trip."""
    h, w = _f32(h), _f32(w)
    R, H = h.shape
    ws = workspace(_lib.query("asme_b200_score_topk_flagged_workspace_bytes", R, w.shape[0]), h.device)
    tgt = None if target is None else _i64(target)
    if _lib.timing is not None:
        _lib.note = f"R={R},V={w.shape[0]},H={H},k={k},flagged=1"
    _lib.call("asme_b200_score_topk_flagged", _p(h), R, H, _p(w), _p(bias), v0, w.shape[0], _p(tgt), _p(target_score), k, _p(row_flag),
              _p(topk_val), _p(topk_idx), _p(rank), _p(ws), ws.numel(), _stream())


def topk_merge(vals: torch.Tensor, idx: torch.Tensor, k: int):
    """vals/idx: (G,R,k) partial lists -> merged (R,k)"""
    vals, idx = _f32(vals), idx.contiguous()
    G, R, kk = vals.shape
    assert kk == k
    out_v = torch.empty(R, k, dtype=torch.float32, device=vals.device)
    out_i = torch.empty(R, k, dtype=torch.int32, device=vals.device)
    _lib.call("asme_b200_topk_merge", _p(vals), _p(idx), G, R, k, _p(out_v), _p(out_i), _stream())
    return out_v, out_i


def ranking_metrics(rank: torch.Tensor, ks: torch.Tensor, out: torch.Tensor):
    """out (4, n_k) += sums of recall, NDCG, MRR, precision @ ks over the batch."""
    _lib.call("asme_b200_ranking_metrics", _p(rank), rank.numel(), _p(ks), ks.numel(), _p(out), _stream())


def dense_ranking(pred: torch.Tensor, pos_mask: torch.Tensor, metric_mask: Optional[torch.Tensor], k: int) -> torch.Tensor:
    """(7,N): recall, precision, DCG, NDCG, MRR, F1 @k and full rank, from dense (N,I) predictions."""
    pred = _f32(pred, "predictions")
    N, I = pred.shape
    pm = _i64(pos_mask, "positive_item_mask")
    mm = None if metric_mask is None else _i64(metric_mask, "metric_mask")
    out = torch.empty(7, N, dtype=torch.float32, device=pred.device)
    _lib.call("asme_b200_dense_ranking", _p(pred), _p(pm), _p(mm), N, I, int(k), _p(out), _stream())
    return out


def score_ce_partial(h, w, bias, target, v0: int = 0, n_live=None):
    """returns (row_max, row_sumexp, target_logit) over the slice [v0, v0+Vloc)"""
    h, w = _f32(h), _f32(w)
    R, H = h.shape
    Vloc = w.shape[0]
    dev = h.device
    rmax = torch.empty(R, dtype=torch.float32, device=dev)
    rsum = torch.empty_like(rmax)
    tl = torch.empty_like(rmax)
    ws_bytes = _lib.query("asme_b200_score_ce_workspace_bytes", R, Vloc)
    ws = workspace(ws_bytes, dev)
    if _lib.timing is not None:
        _lib.note = f"R={_note_rows(R, n_live)},V={Vloc},H={H}"
    _lib.call("asme_b200_score_ce_partial", _p(h), R, H, _p(w), _p(bias), v0, Vloc, _p(_i64(target)), _p(rmax), _p(rsum),
              _p(tl), _p(ws), ws.numel(), _p(n_live), _stream())
    return rmax, rsum, tl


def ce_loss_from_partials(rmax, rsum, tl, loss_sum: torch.Tensor, n_live=None, loss_mean: Optional[torch.Tensor] = None):
    """lse (R); loss_sum += sum of the rows' negative log-likelihoods; ``loss_mean`` (1) = loss_sum / number of (live) rows"""
    lse = torch.empty_like(rmax)
    if _lib.timing is not None:
        _lib.note = f"R={_note_rows(rmax.numel(), n_live)}"
    _lib.call("asme_b200_ce_loss_from_partials", _p(rmax), _p(rsum), _p(tl), rmax.numel(), _p(lse), _p(loss_sum), _p(n_live),
              _p(loss_mean), _stream())
    return lse


def ce_combine(rmax: torch.Tensor, rsum: torch.Tensor):
    """(G,R) per-shard row maxima / sum-exps -> (R) maxima / sum-exps over the whole catalog"""
    rmax, rsum = _f32(rmax), _f32(rsum)
    G, R = rmax.shape
    m = torch.empty(R, dtype=torch.float32, device=rmax.device)
    s = torch.empty_like(m)
    _lib.call("asme_b200_ce_combine", _p(rmax), _p(rsum), G, R, _p(m), _p(s), _stream())
    return m, s


def ce_rescale(rsum: torch.Tensor, rmax: torch.Tensor, gmax: torch.Tensor) -> torch.Tensor:
    """rsum * exp(rmax - gmax): a shard's sum-exp against the all-reduced row maximum"""
    out = torch.empty_like(rsum)
    _lib.call("asme_b200_ce_rescale", _p(_f32(rsum)), _p(_f32(rmax)), _p(_f32(gmax)), rsum.numel(), _p(out), _stream())
    return out


def score_ce_bwd(h, w, bias, target, lse, scale: float, dW: Optional[torch.Tensor], dbias: Optional[torch.Tensor],
                 need_dh: bool = True, v0: int = 0, n_live=None):
    h, w = _f32(h), _f32(w)
    R, H = h.shape
    Vloc = w.shape[0]
    dh = torch.empty_like(h) if need_dh else None
    ws_bytes = _lib.query("asme_b200_score_ce_bwd_workspace_bytes", R, H, Vloc)
    ws = workspace(ws_bytes, h.device)
    if _lib.timing is not None:
        _lib.note = f"R={_note_rows(R, n_live)},V={Vloc},H={H}"
    _lib.call("asme_b200_score_ce_bwd", _p(h), R, H, _p(w), _p(bias), v0, Vloc, _p(_i64(target)), _p(lse), float(scale),
              _p(dh), _p(dW), _p(dbias), _p(ws), ws.numel(), _p(n_live), _stream())
    return dh


def posneg_bce_fwd(h, table, pos, neg, mask, sums: Optional[torch.Tensor]):
    h, table = _f32(h), _f32(table)
    T, H = h.shape
    pl = torch.empty(T, dtype=torch.float32, device=h.device)
    nl = torch.empty_like(pl)
    _lib.call("asme_b200_posneg_bce_fwd", _p(h), _p(table), _p(_i64(pos)), _p(_i64(neg)), _p(_u8(mask)), T, H, _p(pl),
              _p(nl), _p(sums), _stream())
    return pl, nl


def posneg_bce_bwd(h, table, pos, neg, mask, pl, nl, sums, dloss: float = 1.0):
    h, table = _f32(h), _f32(table)
    T, H = h.shape
    dh = torch.empty_like(h)
    dpos = torch.empty_like(h)
    dneg = torch.empty_like(h)
    _lib.call("asme_b200_posneg_bce_bwd", _p(h), _p(table), _p(_i64(pos)), _p(_i64(neg)), _p(_u8(mask)), T, H, _p(pl),
              _p(nl), _p(sums), float(dloss), _p(dh), _p(dpos), _p(dneg), _stream())
    return dh, dpos, dneg


def gather_rows(x: torch.Tensor, row_index: torch.Tensor, n_live=None) -> torch.Tensor:
    x = _f32(x)
    R, H = row_index.numel(), x.shape[1]
    out = torch.empty(R, H, dtype=torch.float32, device=x.device)
    if _lib.timing is not None:
        _lib.note = f"R={_note_rows(R, n_live)},H={H}"
    _lib.call("asme_b200_gather_rows", _p(x), _p(_i64(row_index)), R, H, _p(out), _p(n_live), _stream())
    return out


def scatter_rows(rows: torch.Tensor, row_index: torch.Tensor, out: torch.Tensor, n_live=None):
    rows = _f32(rows)
    if _lib.timing is not None:
        _lib.note = f"R={_note_rows(rows.shape[0], n_live)},H={rows.shape[1]}"
    _lib.call("asme_b200_scatter_rows", _p(rows), _p(_i64(row_index)), rows.shape[0], rows.shape[1], _p(out), _p(n_live), _stream())


def select_rows(target: torch.Tensor, ignore_id: int):
    """(rows, row_targets, n_rows): the flat positions whose target is not ``ignore_id`` (ascending), their targets, and their
    number as an int32 DEVICE scalar.  rows / row_targets have the capacity ``target.numel()``; slots past ``n_rows`` carry
    -1 / ``ignore_id``.  No host synchronisation (``torch.nonzero`` needs one to size its result)."""
    target = _i64(target).reshape(-1)
    T = target.numel()
    rows = torch.empty(T, dtype=torch.int64, device=target.device)
    row_targets = torch.empty(T, dtype=torch.int64, device=target.device)
    n_rows = torch.empty(1, dtype=torch.int32, device=target.device)
    ws = workspace(_lib.query("asme_b200_select_rows_workspace_bytes", T), target.device)
    if _lib.timing is not None:
        _lib.note = f"T={T}"
    _lib.call("asme_b200_select_rows", _p(target), T, int(ignore_id), _p(rows), _p(row_targets), _p(n_rows), _p(ws), ws.numel(), _stream())
    return rows, row_targets, n_rows


def clip_grad_norm(grad: torch.Tensor, max_norm: float, norm_out: Optional[torch.Tensor] = None):
    """grad *= min(1, max_norm / (||grad||_2 + 1e-6)) over the flat gradient arena (torch.nn.utils.clip_grad_norm_)"""
    ws = workspace(max(_lib.query("asme_b200_clip_grad_norm_workspace_bytes"), 256), grad.device, slot=2)
    _lib.call("asme_b200_clip_grad_norm", _p(grad), grad.numel(), float(max_norm), _p(norm_out), _p(ws), ws.numel(), _stream())


def adam_step(param, grad, m, v, lr, beta1, beta2, eps, weight_decay, step):
    if _lib.timing is not None:
        _lib.note = f"n={param.numel()}"
    _lib.call("asme_b200_adam_step", _p(param), _p(grad), _p(m), _p(v), param.numel(), float(lr), float(beta1),
              float(beta2), float(eps), float(weight_decay), int(step), _stream())


def step_state_advance(state: torch.Tensor):
    _lib.call("asme_b200_step_state_advance", _p(state), _stream())


def adam_step_dev(param, grad, m, v, state: torch.Tensor, beta1, beta2, eps, weight_decay):
    if _lib.timing is not None:
        _lib.note = f"n={param.numel()}"
    _lib.call("asme_b200_adam_step_dev", _p(param), _p(grad), _p(m), _p(v), param.numel(), _p(state), float(beta1), float(beta2),
              float(eps), float(weight_decay), _stream())


def fill(x: torch.Tensor, value: float):
    if _lib.timing is not None:
        _lib.note = f"n={x.numel()}"
    _lib.call("asme_b200_fill", _p(x), x.numel(), float(value), _stream())
