"""Encoder projectors (reference: src/slam_llm/models/projector.py).  Parameter names (linear1/linear2) are part of
the checkpoint format.  The arithmetic runs in slam_llm_b200 kernels once the module is bound to a step arena by
slam_model.__init__ (tcgen05 GEMM + bias/ReLU epilogues); before binding the module only carries its initial values."""
import torch
import torch.nn as nn


def _bind(module, engine_projector, arena, names) -> None:
    """Move the parameters into the flat trainable arena (the new nn.Parameters are views of arena.param)."""
    pre = engine_projector.PREFIX
    for name in names:
        mod = getattr(module, name)
        for suffix in ("weight", "bias"):
            view = arena.view(f"{pre}{name}.{suffix}")
            old = getattr(mod, suffix)
            view.copy_(old.data.to(view.device, view.dtype))
            mod._parameters[suffix] = nn.Parameter(view, requires_grad=old.requires_grad)
    module._b200 = engine_projector


class _ProjectorFn(torch.autograd.Function):
    """The projector as an autograd node, for recipes that call `self.encoder_projector(encoder_outs)` themselves and feed the result to
    `self.llm(inputs_embeds=...)`: backward writes the weight / bias gradients into the flat arena (the encoder is frozen: no dX)."""

    @staticmethod
    def forward(ctx, anchor, x, module):
        ctx.module = module
        return module._b200.forward(x.to(torch.bfloat16).contiguous(), save=True)

    @staticmethod
    def backward(ctx, dy):
        ctx.module._b200.backward(dy.to(torch.bfloat16).contiguous())
        return None, None, None


def _forward(module, x):
    if module._b200 is None:
        raise RuntimeError(f"{type(module).__name__} is not bound to a B200 step yet (construct slam_model first); no CPU fallback")
    step = getattr(module, "_step", None)
    if step is not None:
        step.flush_update()                                      # a deferred optimizer step must land before the weights are read
    if torch.is_grad_enabled():
        return _ProjectorFn.apply(module._b200.arena.param, x, module)
    return module._b200.forward(x.to(torch.bfloat16).contiguous(), save=False)


class EncoderProjectorConcat(nn.Module):
    """concat k frames -> Linear(k*d, 2048) -> ReLU -> Linear(2048, llm_dim)   (projector.py:5-27)."""

    def __init__(self, config):
        super().__init__()
        self.k = config.encoder_projector_ds_rate
        self.encoder_dim = config.encoder_dim
        self.llm_dim = config.llm_dim
        self.linear1 = nn.Linear(self.encoder_dim * self.k, 2048)
        self.relu = nn.ReLU()
        self.linear2 = nn.Linear(2048, config.llm_dim)
        self._b200 = None  # slam_llm_b200.engine.ProjectorB200 after binding

    kind = "linear"

    def bind(self, engine_projector, arena) -> None:
        _bind(self, engine_projector, arena, ("linear1", "linear2"))

    def forward(self, x):
        return _forward(self, x)


class EncoderProjectorCov1d(nn.Module):
    """Conv1d(d, d, k, stride k) -> ReLU -> Linear(d, 2048) -> ReLU -> Linear(2048, llm_dim)   (projector.py:29-49)."""
    kind = "cov1d-linear"

    def __init__(self, config):
        super().__init__()
        self.k = config.encoder_projector_ds_rate
        self.encoder_dim = config.encoder_dim
        self.llm_dim = config.llm_dim
        self.conv1d = nn.Conv1d(in_channels=self.encoder_dim, out_channels=self.encoder_dim, kernel_size=self.k, stride=self.k, padding=0)
        self.linear1 = nn.Linear(self.encoder_dim, 2048)
        self.relu1 = nn.ReLU()
        self.linear2 = nn.Linear(2048, self.llm_dim)
        self.relu2 = nn.ReLU()
        self._b200 = None

    def bind(self, engine_projector, arena) -> None:
        _bind(self, engine_projector, arena, ("conv1d", "linear1", "linear2"))

    def forward(self, x):
        return _forward(self, x)


class EncoderProjectorQFormer(nn.Module):
    def __init__(self, config):
        super().__init__()
        raise NotImplementedError("q-former projector is out of scope of the B200 hot path (SURVEY.md §2.1)")
