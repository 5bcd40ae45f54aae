"""Fused Adam over the flat parameter arena (K22): ONE kernel launch updates every parameter of the model.
torch.optim.Adam semantics (L2 weight decay added to the gradient, bias correction), reference defaults
beta = (0.99, 0.998) (modules/masked_training_module.py:37-38).  Subclasses torch.optim.Optimizer so that
LambdaLR schedulers and ``optimizer.zero_grad()`` / ``state_dict()`` keep working."""
import torch

from . import ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-3, betas=(0.99, 0.998), eps=1e-8, weight_decay=0.0):
        self.model = model
        params = [p for _, p in model._named_arena_params()]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._steps = 0
        self.step_state = None      # graphs.StepState: when set, step count and learning rate are read from device memory

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        arena = self.model._arena
        group = self.param_groups[0]
        g = arena.ensure_grad()
        m, v = arena.ensure_moments()
        self._steps += 1
        if self.step_state is not None:
            ops.adam_step_dev(arena.flat, g, m, v, self.step_state.tensor, group["betas"][0], group["betas"][1], group["eps"],
                              group["weight_decay"])
        else:
            ops.adam_step(arena.flat, g, m, v, group["lr"], group["betas"][0], group["betas"][1], group["eps"],
                          group["weight_decay"], self._steps)
        arena.bump()        # the kernel wrote the weights through raw pointers: invalidate the bf16 shadow
        return loss

    def zero_grad(self, set_to_none: bool = True):
        # the next fused backward zeroes the flat gradient buffer with one fill kernel when grads are detached
        for _, p in self.model._named_arena_params():
            p.grad = None
