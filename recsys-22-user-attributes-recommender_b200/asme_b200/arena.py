"""Flat parameter arena: all parameters of a model live in ONE contiguous fp32 device buffer (and
their gradients / Adam moments in same-layout buffers), so that

  * the optimizer step is a single fused kernel launch over the arena (K22),
  * the data-parallel gradient exchange is a single NCCL all-reduce of the flat gradient buffer (N1),
  * matrices the kernels want adjacent (Wq|Wk|Wv -> one (3H,H) GEMM operand, (gamma,beta) pairs ->
    one (2,H) gradient target) simply ARE adjacent.

``nn.Parameter`` objects with the reference's state-dict names are views into the arena, so
``state_dict()`` / ``load_state_dict()`` / torch optimizers keep working unchanged.
"""
from typing import Dict, List, Sequence, Tuple

import torch
from torch import nn


class ParamArena:
    ALIGN = 64  # elements (256 B): every entry starts 256-byte aligned -> 128-bit vector loads are always legal

    def __init__(self, specs: Sequence[Tuple[str, Tuple[int, ...]]]):
        self.specs: List[Tuple[str, Tuple[int, ...]]] = [(sp[0], tuple(sp[1])) for sp in specs]
        self.offsets: Dict[str, int] = {}
        self.shapes: Dict[str, Tuple[int, ...]] = dict(self.specs)
        off = 0
        for name, shape in self.specs:
            n = 1
            for s in shape:
                n *= s
            self.offsets[name] = off
            off += n
            if not name.endswith("+"):          # names ending in "+" glue the NEXT entry right behind this one
                off = (off + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.numel = max(off, self.ALIGN)
        self.flat = torch.zeros(self.numel, dtype=torch.float32)
        self.grad = None
        self.exp_avg = None
        self.exp_avg_sq = None
        self.version = 0            # bumped by everything that rewrites the weights behind torch's back (fused Adam, repack)
        self.generation = 0         # bumped when the flat buffer itself is re-created (captured CUDA graphs hold its address)
        self.flat_bf16 = None       # bf16 shadow of ``flat`` for the tensor-core kernels (same layout, same views)
        self._bf16_stamp = None

    def bump(self):
        self.version += 1

    def bf16(self, refresh: bool = True) -> torch.Tensor:
        """bf16 copy of the whole arena, re-cast (one kernel) whenever the fp32 weights changed"""
        from . import ops
        stamp = (self.version, self.flat._version, self.flat.data_ptr())
        if self.flat_bf16 is None or self.flat_bf16.device != self.flat.device or (refresh and stamp != self._bf16_stamp):
            self.flat_bf16 = ops.cast_bf16(self.flat.view(1, -1), ld_out=self.numel).view(-1)
            self._bf16_stamp = stamp
        return self.flat_bf16

    def _numel(self, name):
        shape = self.shapes[name]
        n = 1
        for s in shape:
            n *= s
        return n

    def view(self, name: str, buf: torch.Tensor = None) -> torch.Tensor:
        buf = self.flat if buf is None else buf
        shape = self.shapes[name]
        off = self.offsets[name]
        return buf[off:off + self._numel(name)].view(shape)

    def span(self, first: str, last: str, shape, buf: torch.Tensor = None) -> torch.Tensor:
        """A view covering consecutive glued entries first..last (e.g. Wq|Wk|Wv as (3H,H))."""
        buf = self.flat if buf is None else buf
        a = self.offsets[first]
        b = self.offsets[last] + self._numel(last)
        return buf[a:b].view(shape)

    def to(self, device) -> "ParamArena":
        self.flat = self.flat.to(device)
        for attr in ("grad", "exp_avg", "exp_avg_sq"):
            t = getattr(self, attr)
            if t is not None:
                setattr(self, attr, t.to(device))
        return self

    def ensure_grad(self) -> torch.Tensor:
        if self.grad is None or self.grad.device != self.flat.device:
            self.grad = torch.zeros_like(self.flat)
        return self.grad

    def ensure_moments(self):
        if self.exp_avg is None or self.exp_avg.device != self.flat.device:
            self.exp_avg = torch.zeros_like(self.flat)
            self.exp_avg_sq = torch.zeros_like(self.flat)
        return self.exp_avg, self.exp_avg_sq


def glued(names: Sequence[str]) -> List[str]:
    """Mark all but the last name so that the entries are laid out back to back (no alignment gap)."""
    return [n + "+" for n in names[:-1]] + [names[-1]]


class ArenaModule(nn.Module):
    """nn.Module whose parameters are views into a :class:`ParamArena`.

    Sub-modules are plain containers that reproduce the reference's attribute paths (= state-dict keys);
    ``_apply`` (``.cuda()``, ``.to()``) re-packs the arena on the new device and re-points every parameter."""

    def _init_arena(self, specs: Sequence[Tuple[str, Tuple[int, ...]]]):
        # spec names may carry the "+" glue marker; the parameter path is the name without it
        # an optional third element "T" exposes the parameter as the TRANSPOSE of the stored matrix
        # (LinearUpscaler weights are stored (Va,H) so that the gather reads contiguous rows)
        self._arena = ParamArena(specs)
        self._param_paths = [sp[0].rstrip("+") for sp in specs if not sp[0].startswith("_ghost")]
        self._transposed = {sp[0].rstrip("+") for sp in specs if len(sp) > 2 and sp[2] == "T"}
        for sp in specs:
            name = sp[0]
            if name.startswith("_ghost"):
                continue
            path = name.rstrip("+").split(".")
            mod = self
            for part in path[:-1]:
                if not hasattr(mod, part):
                    setattr(mod, part, nn.Module())
                mod = getattr(mod, part)
            v = self._arena.view(name)
            mod.register_parameter(path[-1], nn.Parameter(v.t() if name.rstrip("+") in self._transposed else v))

    def _spec_name(self, path: str) -> str:
        return path if path in self._arena.offsets else path + "+"

    def weight(self, path: str, buf: torch.Tensor = None) -> torch.Tensor:
        return self._arena.view(self._spec_name(path), buf)

    def weights_span(self, first: str, last: str, shape, buf: torch.Tensor = None) -> torch.Tensor:
        return self._arena.span(self._spec_name(first), self._spec_name(last), shape, buf)

    def _named_arena_params(self):
        for path in self._param_paths:
            mod = self
            parts = path.split(".")
            for part in parts[:-1]:
                mod = getattr(mod, part)
            yield path, mod._parameters[parts[-1]]

    def _repack(self, device=None):
        """re-create the flat buffer on ``device`` from the current parameter values and re-point every parameter at it.  The
        gradient buffer and the Adam moments have the same layout and simply move along (a validation pass or a ``.to()`` must
        not reset the optimizer); captured CUDA graphs hold the old addresses, so ``generation`` tells them to re-capture."""
        params = list(self._named_arena_params())
        device = device if device is not None else params[0][1].device
        new_flat = torch.zeros(self._arena.numel, dtype=torch.float32, device=device)
        old_flat = self._arena.flat
        self._arena.flat = new_flat
        self._arena.bump()
        self._arena.generation = getattr(self._arena, "generation", 0) + 1
        self._arena.flat_bf16 = None
        for attr in ("grad", "exp_avg", "exp_avg_sq"):
            t = getattr(self._arena, attr)
            if t is not None and t.device != new_flat.device:
                setattr(self._arena, attr, t.to(new_flat.device))
        with torch.no_grad():
            for path, p in params:
                v = self._arena.view(self._spec_name(path))
                if path in self._transposed:
                    v = v.t()
                v.copy_(p.data.to(device=device, dtype=torch.float32))
                p.data = v
                p.grad = None
        del old_flat

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        params = list(self._named_arena_params())
        # ``.to()`` / ``.cuda()`` onto the device the arena already lives on leaves every parameter where it is: nothing to do
        # (re-packing would hand out a new buffer -- and used to drop the Adam moments -- on every validation pass)
        if params and not (self.arena_is_intact() and all(p.dtype == torch.float32 for _, p in params)):
            self._repack(params[0][1].device)
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._arena.bump()
        return out

    def weight_bf16(self, path: str) -> torch.Tensor:
        """view of a parameter inside the bf16 shadow arena (tensor-core operand)"""
        return self._arena.view(self._spec_name(path), self._arena.bf16())

    def weights_span_bf16(self, first: str, last: str, shape) -> torch.Tensor:
        return self._arena.span(self._spec_name(first), self._spec_name(last), shape, self._arena.bf16())

    def arena_is_intact(self) -> bool:
        base = self._arena.flat.data_ptr()
        for path, p in self._named_arena_params():
            off = self._arena.offsets[self._spec_name(path)]
            if p.data_ptr() != base + 4 * off or p.device != self._arena.flat.device:  # .t() keeps data_ptr
                return False
        return True

    def attach_grads(self):
        """Expose the flat gradient buffer through ``param.grad`` views (what torch optimizers read)."""
        g = self._arena.ensure_grad()
        for path, p in self._named_arena_params():
            gv = self._arena.view(self._spec_name(path), g)
            p.grad = gv.t() if path in self._transposed else gv
