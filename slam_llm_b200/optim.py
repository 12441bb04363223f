"""AdamW over the flat trainable arena as a torch.optim.Optimizer (so LambdaLR and the reference train loop drive it):
one slam_adamw kernel per step instead of torch's foreach AdamW over ~130 tensors
(reference: optim.AdamW(model.parameters(), lr, weight_decay) at src/slam_llm/pipeline/finetune.py:247-251).

Parameters that a recipe adds OUTSIDE the arena (e.g. the s2s group-decode adapter, examples/s2s/utils/projector_utils.py: a plain nn.Linear on
the logits) are ordinary torch tensors: they are stepped by an inner torch.optim.AdamW with the same hyper-parameters, so
`optim.AdamW(model.parameters(), ...)` semantics hold for the whole model."""
from __future__ import annotations

import torch


class FlatAdamW(torch.optim.Optimizer):
    def __init__(self, model, lr: float = 1e-4, weight_decay: float = 0.0, betas=(0.9, 0.999), eps: float = 1e-8):
        step_engine = getattr(getattr(model, "module", model), "b200", None)
        if step_engine is None:
            raise TypeError("FlatAdamW needs a slam_model built on the B200 step (model.b200)")
        self.engine = step_engine
        self.model = getattr(model, "module", model)
        super().__init__([step_engine.arena.param], dict(lr=lr, weight_decay=weight_decay, betas=betas, eps=eps))
        lo = step_engine.arena.param.data_ptr()
        hi = lo + step_engine.arena.param.numel() * 4
        self.foreign = [p for p in self.model.parameters() if p.requires_grad and not (lo <= p.data_ptr() < hi)]
        self.inner = torch.optim.AdamW(self.foreign, lr=lr, weight_decay=weight_decay, betas=betas, eps=eps) if self.foreign else None

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        world = getattr(self.model, "ddp_world_size", 1)
        # recorded now, applied by the engine (immediately, or - deferred data-parallel mode - after the next step's frozen front end)
        self.engine.optimizer_step(g["lr"], g["weight_decay"], grad_div=float(world), betas=g["betas"], eps=g["eps"])
        if self.inner is not None:
            for ig in self.inner.param_groups:                  # LambdaLR drives this optimizer's group: mirror it
                ig["lr"], ig["weight_decay"] = g["lr"], g["weight_decay"]
            if world > 1:
                for p in self.foreign:                          # DDP mean of the foreign gradients (tiny tensors)
                    if p.grad is not None:
                        torch.distributed.all_reduce(p.grad)
                        p.grad.div_(world)
            self.inner.step()

    def zero_grad(self, set_to_none: bool = True):
        # the backward kernels overwrite the flat gradient buffer on the first micro-step: nothing to clear there
        self.engine.micro_steps = 0
        if self.inner is not None:
            self.inner.zero_grad(set_to_none=set_to_none)
