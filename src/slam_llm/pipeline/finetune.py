"""Training entry point (reference: src/slam_llm/pipeline/finetune.py:47-282): same config split, logging, seeding,
process-group setup, plugin-loaded model factory and dataset, LambdaLR warmup -> linear decay, then train().

B200 differences: the model is built directly on the GPU (no model.to()), DDP is the flat-arena all-reduce inside the
step (only the projector + LoRA gradients cross NVLink: one NCCL all-reduce per optimizer step), and the optimizer is
the single-kernel FlatAdamW.  FSDP / DeepSpeed wrappers are out of scope (no asr_* recipe enables them)."""
import logging
import os
import random

import hydra
import torch
from omegaconf import DictConfig, OmegaConf

from slam_llm.utils.config_utils import get_dataloader_kwargs
from slam_llm.utils.dataset_utils import get_preprocessed_dataset
from slam_llm.utils.model_utils import get_custom_model_factory
from slam_llm.utils.train_utils import clear_gpu_cache, setup, setup_environ_flags, train
from slam_llm_b200.optim import FlatAdamW


@hydra.main(config_name=None, version_base=None)
def main_hydra(cfg: DictConfig):
    kwargs = cfg
    logging.basicConfig(level=getattr(logging, kwargs.get("log_level", "INFO").upper()))
    main(kwargs)


def main(kwargs: DictConfig):
    train_config, fsdp_config, model_config, log_config, dataset_config = (kwargs.train_config, kwargs.fsdp_config, kwargs.model_config,
                                                                          kwargs.log_config, kwargs.dataset_config)
    fsdp_config.use_fp16 = train_config.use_fp16
    OmegaConf.set_struct(kwargs, False)
    for k in ("train_config", "fsdp_config", "model_config", "log_config", "dataset_config"):
        del kwargs[k]
    OmegaConf.set_struct(kwargs, True)

    log_dir = os.path.dirname(log_config.log_file)
    if log_dir and not os.path.exists(log_dir):
        os.makedirs(log_dir, exist_ok=True)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s | %(levelname)s | %(name)s | %(message)s", datefmt="%Y-%m-%d %H:%M:%S")
    logger = logging.getLogger()
    logger.setLevel(logging.INFO)
    fmt = logging.Formatter("[%(asctime)s][%(name)s][%(levelname)s] - %(message)s", datefmt="%Y-%m-%d %H:%M:%S")
    file_handler = logging.FileHandler(filename=log_config.log_file, mode="w")
    file_handler.setLevel(logging.INFO)
    file_handler.setFormatter(fmt)
    if logger.handlers:
        logger.handlers[0].setLevel(logging.INFO)
        logger.handlers[0].setFormatter(fmt)
    logger.addHandler(file_handler)

    if not torch.cuda.is_available():
        raise RuntimeError("slam_llm.pipeline.finetune: no CUDA device; the B200 training step has no CPU fallback")
    torch.cuda.manual_seed(train_config.seed)
    torch.manual_seed(train_config.seed)
    random.seed(train_config.seed)

    if train_config.enable_fsdp:
        raise NotImplementedError("enable_fsdp: FSDP sharding is out of scope (every asr_* recipe uses DDP; SURVEY.md §2.4)")
    distributed = bool(train_config.enable_ddp)
    local_rank = rank = None
    world_size = 1
    if distributed:
        setup()
        local_rank, rank, world_size = int(os.environ["LOCAL_RANK"]), int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        logger.info(f"local_rank: {local_rank}, rank: {rank}, world_size: {world_size}")
    if torch.distributed.is_initialized():
        torch.cuda.set_device(local_rank)
        clear_gpu_cache(local_rank)
        setup_environ_flags(rank)
    is_main = not distributed or rank == 0
    if is_main:
        for name, c in (("train_config", train_config), ("fsdp_config", fsdp_config), ("model_config", model_config), ("log_config", log_config)):
            logger.info("{}: {}".format(name, c))
        if log_config.use_wandb:
            import wandb
            os.makedirs(log_config.wandb_dir, exist_ok=True)
            wandb.init(dir=log_config.wandb_dir, entity=log_config.wandb_entity_name, project=log_config.wandb_project_name,
                       name=log_config.wandb_exp_name, config={"train_config": train_config, "model_config": model_config})

    model_factory = get_custom_model_factory(model_config, logger)
    model, tokenizer = model_factory(train_config, model_config, **kwargs)
    if distributed:
        # DDP (finetune.py:181-184): replicas hold identical frozen weights (same seed / same checkpoint); trainables are broadcast
        # once from rank 0, and each step all-reduces the flat projector+LoRA gradient buffer inside model._backward.
        torch.distributed.broadcast(model.b200.arena.param.data, src=0)
        model.ddp_world_size = world_size
        model.b200.defer_update = bool(train_config.get("b200_overlap_allreduce", True))   # async all-reduce + update applied behind the next front end

    logger.info("dataset_config: {}".format(dataset_config))
    dataset_train = get_preprocessed_dataset(tokenizer, dataset_config, split="train")
    if is_main and train_config.batching_strategy != "dynamic":
        logger.info(f"--> Training Set Length = {len(dataset_train)}")
    dataset_val = get_preprocessed_dataset(tokenizer, dataset_config, split="val")
    if is_main and train_config.batching_strategy != "dynamic":
        logger.info(f"--> Validation Set Length = {len(dataset_val)}")
    train_dataloader = torch.utils.data.DataLoader(dataset_train, num_workers=train_config.num_workers_dataloader, pin_memory=True,
                                                   **get_dataloader_kwargs(train_config, dataset_train, tokenizer, "train"))
    eval_dataloader = None
    if train_config.run_validation:
        eval_dataloader = torch.utils.data.DataLoader(dataset_val, num_workers=train_config.num_workers_dataloader, pin_memory=True,
                                                      **get_dataloader_kwargs(train_config, dataset_val, tokenizer, "val"))

    optimizer = FlatAdamW(model, lr=train_config.lr, weight_decay=train_config.weight_decay)
    scheduler = torch.optim.lr_scheduler.LambdaLR(
        optimizer,
        lr_lambda=lambda step: (min(step / train_config.warmup_steps, 1) if step < train_config.warmup_steps
                                else max(0.0, 1 - (step - train_config.warmup_steps) / (train_config.total_steps - train_config.warmup_steps))))

    results = train(model, train_dataloader, eval_dataloader, tokenizer, optimizer, scheduler, train_config.gradient_accumulation_steps,
                    train_config, log_config, None, local_rank if distributed else None, rank if distributed else None)
    if is_main:
        for k, v in results.items():
            logger.info(f"Key: {k}, Value: {v}")
        if log_config.use_wandb:
            import wandb
            wandb.finish()
    return results


if __name__ == "__main__":
    main_hydra()
