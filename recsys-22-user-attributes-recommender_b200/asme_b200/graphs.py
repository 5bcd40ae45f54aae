"""Whole-step CUDA graphs.  A training step of the C2-C4 shapes is ~90 kernel launches of 5-300 us each; issued one by one
from Python the host needs ~2.7 ms per step while the kernels need ~2 ms, i.e. the GPU idles.  ``GraphedTrainStep`` captures
zero_grad -> training_step -> backward -> fused Adam for one batch SHAPE into a CUDA graph and replays it; everything
that changes from step to step lives in device memory: the dropout seed and the Adam step count (:class:`StepState`, bumped
by the first node of the graph), the learning rate (written by the host before each replay) and the number of positions that
carry a target (``ops.select_rows``: the selected-rows path has capacity-sized buffers and a device-side row count).  One
graph therefore serves every batch of a ``(B, S)`` -- cloze masking draws a different number of targets for every batch.

The captured step is exactly the eager step (same kernels, same order, same arithmetic); eager mode stays the default and the
reference API (training_step / loss.backward() / optimizer.step()) is unchanged."""
from typing import Any, Callable, Dict

import torch

from . import ops


class StepState:
    """device struct {uint64 seed; float64 lr; int64 adam_step} (include/asme_b200.h, asme_b200_step_state_advance)"""

    def __init__(self, device, seed: int = 0, adam_step: int = 0, lr: float = 0.0):
        self.tensor = torch.tensor([seed & 0x7FFFFFFFFFFFFFFF, 0, adam_step], dtype=torch.int64).to(device)
        self._lr_view = self.tensor.view(torch.float64)[1:2]
        # a ring of pinned slots: the copy is asynchronous, and a host that runs ahead of the device (a loop that reads the loss one step
        # late) must not overwrite the value of a step whose copy has not been executed yet
        self._lr_host = torch.zeros(16, dtype=torch.float64).pin_memory()
        self._lr_slot = 0
        self.set_lr(lr)

    def set_lr(self, lr: float):
        i = self._lr_slot = (self._lr_slot + 1) % self._lr_host.numel()
        self._lr_host[i] = lr
        self._lr_view.copy_(self._lr_host[i:i + 1], non_blocking=True)

    def indirect_seed(self) -> int:
        return (1 << 63) | self.tensor.data_ptr()


class GraphedTrainStep:
    """captures ``step_fn(batch)`` (zero_grad + training_step + backward + optimizer.step) once per batch signature"""

    def __init__(self, module, optimizer, scheduler=None, warmup_iters: int = 3, grad_hook: Callable[[], None] = None):
        self.module, self.optimizer, self.scheduler = module, optimizer, scheduler
        self.grad_hook = grad_hook          # e.g. the data-parallel all-reduce of the flat gradient arena (captured too)
        self.model = module.model
        dev = next(self.model.parameters()).device
        lr = optimizer.param_groups[0]["lr"]
        self.state = StepState(dev, seed=(int(self.model._seed) << 32) + self.model._step_counter, adam_step=optimizer._steps, lr=lr)
        self.model._step_state = self.state
        optimizer.step_state = self.state
        self.graphs: Dict[Any, Any] = {}
        self.warmup_iters = warmup_iters

    def _eager(self, batch):
        ops.step_state_advance(self.state.tensor)
        self.optimizer.zero_grad()
        with torch.no_grad():                       # direct mode: modules.fused_loss parks the fused backward on the model
            out = self.module.training_step(batch, 0)
            self.model._pending_backward()
            self.model._pending_backward = None
        if self.grad_hook is not None:
            self.grad_hook()
        self.optimizer.step()
        return out["loss"].detach()

    def _capture(self, key, batch):
        dev = self.state.tensor.device
        static = {k: (v.to(dev, copy=True) if torch.is_tensor(v) else v) for k, v in batch.items()}
        arena = self.model._arena
        arena.ensure_grad()
        arena.ensure_moments()
        # the warm-up iterations are REAL steps (they grow workspaces and allocator pools and fetch driver entry points): weights,
        # moments, step state and host counters are put back afterwards, so that the batch that triggers a capture is trained on
        # exactly once -- by the first replay
        snapshot = (arena.flat.clone(), arena.exp_avg.clone(), arena.exp_avg_sq.clone(), self.state.tensor.clone(),
                    self.optimizer._steps, self.model._step_counter)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup_iters):
                self._eager(static)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss = self._eager(static)
        arena.flat.copy_(snapshot[0])
        arena.exp_avg.copy_(snapshot[1])
        arena.exp_avg_sq.copy_(snapshot[2])
        self.state.tensor.copy_(snapshot[3])
        self.optimizer._steps, self.model._step_counter = snapshot[4], snapshot[5]
        arena.bump()
        self.graphs[key] = (graph, static, loss, arena.generation)
        return self.graphs[key]

    def __call__(self, batch: Dict[str, torch.Tensor], key=None) -> torch.Tensor:
        """batch tensors are copied into the graph's static inputs (device-to-device or pinned host-to-device); returns the
        loss tensor of the replayed step (static: read it before the next replay of the same graph)"""
        batch = {k: v for k, v in batch.items() if torch.is_tensor(v)}
        if key is None:
            key = tuple((k, tuple(v.shape), str(v.dtype)) for k, v in sorted(batch.items()))
        entry = self.graphs.get(key)
        if entry is not None and entry[3] != self.model._arena.generation:
            entry = None            # the parameter arena was re-created (a real device move): the graph points at freed memory
        if entry is None:
            entry = self._capture(key, batch)
        graph, static, loss, _ = entry
        for k, v in batch.items():
            if static[k].data_ptr() != v.data_ptr():
                static[k].copy_(v, non_blocking=True)
        self.state.set_lr(self.optimizer.param_groups[0]["lr"])
        graph.replay()
        self.optimizer._steps += 1
        self.model._step_counter += 1
        if self.scheduler is not None:
            self.scheduler.step()
        return loss


class GraphedEvalStep:
    """captures ``fn(batch) -> dict of tensors`` (e.g. ``model.evaluate_rank`` of one batch: embedding, encoder, catalog sweeps,
    merges -- ~40 launches that the host cannot issue as fast as the GPU retires them) once per batch signature and replays it.
    ``fn`` must not synchronise with the host.  Non-tensor entries of the result are dropped; the returned tensors are the
    graph's static outputs: consume them (e.g. ``metrics.update``) before the next replay of the same signature."""

    def __init__(self, fn: Callable[[Dict[str, torch.Tensor]], Dict[str, Any]], warmup_iters: int = 2):
        self.fn, self.warmup_iters = fn, warmup_iters
        self.graphs: Dict[Any, Any] = {}

    def _capture(self, key, batch):
        static = {k: v.clone() for k, v in batch.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(self.warmup_iters):
                self.fn(static)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph), torch.no_grad():
            out = self.fn(static)
        out = {k: v for k, v in out.items() if torch.is_tensor(v)}
        self.graphs[key] = (graph, static, out)
        return self.graphs[key]

    def __call__(self, batch: Dict[str, torch.Tensor], key=None) -> Dict[str, torch.Tensor]:
        if key is None:
            key = tuple((k, tuple(v.shape)) for k, v in sorted(batch.items()))
        entry = self.graphs.get(key)
        if entry is None:
            entry = self._capture(key, batch)
        graph, static, out = entry
        for k, v in batch.items():
            if static[k].data_ptr() != v.data_ptr():
                static[k].copy_(v, non_blocking=True)
        graph.replay()
        return out
