"""PEFT config + DataLoader kwargs (reference: src/slam_llm/utils/config_utils.py:46-114)."""
import logging

import torch.distributed as dist
from omegaconf import OmegaConf
from torch.utils.data import DistributedSampler

from slam_llm_b200.config import LoraCfg

logger = logging.getLogger(__name__)


def generate_peft_config(train_config) -> LoraCfg:
    """train_config.peft_config -> LoraCfg (the fields of peft.LoraConfig the hot path uses).  Only `lora` is on the
    B200 path (llama_adapter / prefix tuning are out of scope)."""
    params = OmegaConf.to_container(train_config.peft_config, resolve=True)
    method = params.pop("peft_method", "lora")
    if method != "lora":
        raise NotImplementedError(f"peft_method={method!r}: only LoRA adapters are implemented on the B200 path")
    if params.get("bias", "none") != "none":
        raise NotImplementedError("LoRA bias != 'none' is not supported")
    return LoraCfg(r=int(params.get("r", 8)), alpha=params.get("lora_alpha", 32), targets=tuple(params.get("target_modules", ("q_proj", "v_proj"))),
                   dropout=float(params.get("lora_dropout", 0.0)))


def get_dataloader_kwargs(train_config, dataset, tokenizer, mode):
    kwargs = {}
    batch_size = train_config.batch_size_training if mode == "train" else train_config.val_batch_size
    distributed = train_config.enable_fsdp or train_config.enable_ddp or train_config.get("enable_deepspeed", False)
    if train_config.batching_strategy in ("padding", "packing"):
        raise NotImplementedError(f"batching_strategy={train_config.batching_strategy!r} is a text-only llama-recipes strategy "
                                  "(out of scope, SURVEY.md §2.1); ASR recipes use 'custom' or 'dynamic'")
    if train_config.batching_strategy == "dynamic":
        kwargs.update(sampler=None, batch_size=None, drop_last=False, collate_fn=dataset.collator)
    else:
        if distributed:
            kwargs["sampler"] = DistributedSampler(dataset, rank=dist.get_rank(), num_replicas=dist.get_world_size(), shuffle=mode == "train")
        kwargs.update(batch_size=batch_size, drop_last=True, collate_fn=dataset.collator)
    logger.info(f"Using batching strategy: {train_config.batching_strategy}")
    return kwargs
