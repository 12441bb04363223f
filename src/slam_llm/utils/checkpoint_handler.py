"""Trainable-only checkpoint in the reference's format (src/slam_llm/utils/checkpoint_handler.py:185-201):
model.pt = {name: tensor} for every parameter with requires_grad, under the reference key names
(`encoder_projector.linear1.weight`, `llm.base_model.model.model.layers.N.self_attn.q_proj.lora_A.default.weight`, ...)."""
import logging
import os

import torch

logger = logging.getLogger(__name__)


def save_model_checkpoint_peft(model, optimizer, rank, cfg, checkpoint_name="checkpoint", save_trainable_only=True):
    logger.info("--> saving model ...")
    save_dir = os.path.join(cfg.output_dir, checkpoint_name)
    os.makedirs(save_dir, exist_ok=True)
    save_full_path = os.path.join(save_dir, "model.pt")
    model = getattr(model, "module", model)
    cpu_state = model.state_dict()
    if save_trainable_only:
        trainable = {k for k, v in model.named_parameters() if v.requires_grad}
        cpu_state = {k: v.detach().cpu().clone() for k, v in cpu_state.items() if k in trainable}
    torch.save(cpu_state, save_full_path)
    logger.info(f"encoder saved at {save_full_path}")
