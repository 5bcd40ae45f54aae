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
        if self.step_state is not None:
            # step count and learning rate live in device memory (graphs.StepState): the graph's first node advances them, and
            # GraphedTrainStep keeps ``_steps`` in step with the replays -- a capture is not a step
            ops.adam_step_dev(arena.flat, g, m, v, self.step_state.tensor, group["betas"][0], group["betas"][1], group["eps"],
                              group["weight_decay"])
        else:
            self._steps += 1
            ops.adam_step(arena.flat, g, m, v, group["lr"], group["betas"][0], group["betas"][1], group["eps"],
                          group["weight_decay"], self._steps)
        arena.bump()        # the kernel wrote the weights through raw pointers: invalidate the bf16 shadow
        return loss

    # The moments and the step count live in the arena, not in ``Optimizer.state``: checkpoints carry them explicitly, so that a
    # resumed run continues Adam where it stopped (a Lightning checkpoint of the reference holds exp_avg / exp_avg_sq / step too).
    def state_dict(self):
        out = super().state_dict()
        arena = self.model._arena
        steps = self._steps
        if self.step_state is not None:
            steps = int(self.step_state.tensor[2].item())
        out["asme_b200"] = {"steps": steps,
                            "exp_avg": None if arena.exp_avg is None else arena.exp_avg.detach().cpu().clone(),
                            "exp_avg_sq": None if arena.exp_avg_sq is None else arena.exp_avg_sq.detach().cpu().clone()}
        return out

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        extra = state_dict.pop("asme_b200", None)
        super().load_state_dict(state_dict)
        if extra is None:
            return
        arena = self.model._arena
        self._steps = int(extra["steps"])
        if extra["exp_avg"] is not None:
            m, v = arena.ensure_moments()
            if m.numel() != extra["exp_avg"].numel():
                raise RuntimeError("asme_b200: optimizer state belongs to a model with a different parameter layout")
            m.copy_(extra["exp_avg"].to(m.device))
            v.copy_(extra["exp_avg_sq"].to(v.device))
        if self.step_state is not None:
            self.step_state.tensor[2] = self._steps

    def zero_grad(self, set_to_none: bool = True):
        # the next fused backward zeroes the flat gradient buffer with one fill kernel when grads are detached
        for _, p in self.model._named_arena_params():
            p.grad = None
