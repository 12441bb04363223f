"""AdamW over the flat trainable arena as a torch.optim.Optimizer (so LambdaLR and the reference train loop drive it):
one slam_adamw kernel per step instead of torch's foreach AdamW over ~130 tensors
(reference: optim.AdamW(model.parameters(), lr, weight_decay) at src/slam_llm/pipeline/finetune.py:247-251)."""
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

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        world = getattr(self.model, "ddp_world_size", 1)
        # recorded now, applied by the engine (immediately, or - deferred data-parallel mode - after the next step's frozen front end)
        self.engine.optimizer_step(g["lr"], g["weight_decay"], grad_div=float(world), betas=g["betas"], eps=g["eps"])

    def zero_grad(self, set_to_none: bool = True):
        # the backward kernels overwrite the flat gradient buffer on the first micro-step: nothing to clear
        self.engine.micro_steps = 0
