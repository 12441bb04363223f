"""Encoder projectors (reference: src/slam_llm/models/projector.py).  Parameter names (linear1/linear2) are part of
the checkpoint format.  The arithmetic runs in slam_llm_b200 kernels once the module is bound to a step arena by
slam_model.__init__ (tcgen05 GEMM + bias/ReLU epilogues); before binding the module only carries its initial values."""
import torch
import torch.nn as nn


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

    def bind(self, engine_projector, arena) -> None:
        """Move the four parameters into the flat trainable arena (views share storage with arena.param)."""
        pre = engine_projector.PREFIX
        for mod, name in ((self.linear1, "linear1"), (self.linear2, "linear2")):
            for suffix in ("weight", "bias"):
                view = arena.view(f"{pre}{name}.{suffix}")
                view.copy_(getattr(mod, suffix).data.to(view.device, view.dtype))
                old = getattr(mod, suffix)
                mod._parameters[suffix] = nn.Parameter(view, requires_grad=old.requires_grad)
        self._b200 = engine_projector

    def forward(self, x):
        if self._b200 is None:
            raise RuntimeError("EncoderProjectorConcat is not bound to a B200 step yet (construct slam_model first); no CPU fallback")
        return self._b200.forward(x.to(torch.bfloat16).contiguous(), save=torch.is_grad_enabled())


class EncoderProjectorCov1d(nn.Module):
    """Conv1d(d, d, k, stride k) -> ReLU -> Linear(d, 2048) -> ReLU -> Linear(2048, llm_dim)   (projector.py:29-49)."""

    def __init__(self, config):
        super().__init__()
        raise NotImplementedError("cov1d-linear projector: not yet on the B200 path (SURVEY.md §8a a3 alt; planned next)")


class EncoderProjectorQFormer(nn.Module):
    def __init__(self, config):
        super().__init__()
        raise NotImplementedError("q-former projector is out of scope of the B200 hot path (SURVEY.md §2.1)")
