"""A Lightning-free fit / validate / test loop that drives the same hooks the reference's modules expose to
``pytorch_lightning.Trainer``: training_step -> loss.backward() -> optimizer.step() ->
scheduler.step(); validation_step -> validation_step_end -> validation_epoch_end.

One process per GPU.  With ``torch.distributed`` initialised, training is data parallel: after each fused backward
the FLAT gradient arena is averaged with ONE NCCL all-reduce (N1), and metric states are summed across ranks at
epoch end (N2)."""
from typing import Any, Dict, Iterable, List, Optional

import torch
import torch.distributed as dist


def _to_device(batch: Dict[str, Any], device, non_blocking=True):
    return {k: (v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v) for k, v in batch.items()}


class Trainer:
    def __init__(self, max_epochs: int = 1, device: Optional[torch.device] = None, max_steps: Optional[int] = None,
                 log_every_n_steps: int = 50, gradient_clip_val: Optional[float] = None, process_group=None):
        self.max_epochs, self.max_steps = max_epochs, max_steps
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.log_every_n_steps = log_every_n_steps
        self.gradient_clip_val = float(gradient_clip_val) if gradient_clip_val else None     # pl.Trainer: clip by global L2 norm
        self.process_group = process_group
        self.global_step = 0
        self.history: List[Dict[str, float]] = []

    # ---- distributed helpers -----------------------------------------------------------------------------------
    def _world(self) -> int:
        return dist.get_world_size(self.process_group) if dist.is_available() and dist.is_initialized() else 1

    def _allreduce_grads(self, module):
        world = self._world()
        if world == 1:
            return
        g = module.model._arena.ensure_grad()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.process_group)   # one bucket: the whole arena
        g.div_(world)

    # ---- loops ----------------------------------------------------------------------------------------------------
    def _broadcast_parameters(self, module):
        """data parallel: every replica starts from rank 0's weights (what DDP does at construction), whatever the seeding"""
        if self._world() > 1:
            arena = module.model._arena
            dist.broadcast(arena.flat, src=dist.get_global_rank(self.process_group, 0) if self.process_group is not None else 0,
                           group=self.process_group)
            arena.bump()

    def _clip(self, module):
        if self.gradient_clip_val is not None:
            from . import ops
            ops.clip_grad_norm(module.model._arena.ensure_grad(), self.gradient_clip_val)

    def fit(self, module, train_dataloaders: Iterable, val_dataloaders: Optional[Iterable] = None):
        module.to(self.device)
        self._broadcast_parameters(module)
        opt = module.configure_optimizers()
        schedulers = []
        if isinstance(opt, tuple):
            optimizers, sched = opt
            optimizer = optimizers[0]
            schedulers = [s["scheduler"] if isinstance(s, dict) else s for s in sched]
        elif isinstance(opt, list):
            optimizer = opt[0]
        else:
            optimizer = opt
        for epoch in range(self.max_epochs):
            module.train()
            for batch_idx, batch in enumerate(train_dataloaders):
                batch = _to_device(batch, self.device)
                optimizer.zero_grad(set_to_none=True)
                out = module.training_step(batch, batch_idx)
                loss = out["loss"] if isinstance(out, dict) else out
                loss.backward()
                self._allreduce_grads(module)
                self._clip(module)
                optimizer.step()
                for s in schedulers:
                    s.step()
                self.global_step += 1
                if self.max_steps is not None and self.global_step >= self.max_steps:
                    break
            record = {"epoch": epoch, "train_loss": float(loss.detach())}
            if val_dataloaders is not None:
                record.update(self.validate(module, val_dataloaders))
            self.history.append(record)
            if self.max_steps is not None and self.global_step >= self.max_steps:
                break
        return self.history

    @torch.no_grad()
    def _eval_loop(self, module, loader, step, step_end, epoch_end) -> Dict[str, float]:
        if next(module.parameters()).device != self.device:
            module.to(self.device)
        module.eval()
        for batch_idx, batch in enumerate(loader):
            batch = _to_device(batch, self.device)
            outputs = getattr(module, step)(batch, batch_idx)
            getattr(module, step_end)(outputs)
        if self._world() > 1 and hasattr(module.get_metrics(), "sync"):
            module.get_metrics().sync(self.process_group)
        result = getattr(module, epoch_end)(None)
        return {k: float(v) for k, v in (result or {}).items()}

    def validate(self, module, loader) -> Dict[str, float]:
        return self._eval_loop(module, loader, "validation_step", "validation_step_end", "validation_epoch_end")

    def test(self, module, loader) -> Dict[str, float]:
        return self._eval_loop(module, loader, "test_step", "test_step_end", "test_epoch_end")
