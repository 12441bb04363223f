"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import torch


def round_frozen(om):
    """The device stores FROZEN matrices in bf16: give the oracle the same rounded values so that the comparison
    isolates kernel arithmetic from one-time weight quantisation.  LayerNorm affine params, biases and the sinusoidal
    positions stay fp32 on the device, and all trainables stay fp32 in both."""
    for k in om.enc_w:
        if k.endswith(".weight") and "ln" not in k:
            om.enc_w[k] = om.enc_w[k].bfloat16().float()
    for k in om.llm_w:
        om.llm_w[k] = om.llm_w[k].bfloat16().float()
    return om
