"""Train / eval loop (reference: src/slam_llm/utils/train_utils.py:46-498), same signatures and Python-level
semantics (Appendix D3 of SURVEY.md): loss and acc divided by gradient_accumulation_steps, optimizer step on
(step+1) % accum == 0 or the last batch of a non-dynamic loader, LR scheduler stepped per optimizer step and the loop
breaks when LR hits 0, validation every validation_interval steps, checkpoint only when eval loss improves, epoch
metrics = SUM-allreduce / world / batches, same result-dict keys.

Differences that come from the B200 design (no change in results):
  * the H2D move of a batch also attaches the label-row indices computed on the CPU copy (`_rows`, `_targets`), so the
    step never needs a device->host sync to know which rows carry a label;
  * bf16 compute needs no GradScaler: `use_fp16` keeps the reference control flow but the scale is 1;
  * logging reads loss/acc every step like the reference (tqdm f-string), i.e. one D2H read per step.
"""
from __future__ import annotations

import logging
import os
import time
from contextlib import nullcontext

import torch
import torch.distributed as dist
from tqdm import tqdm

from slam_llm.utils.checkpoint_handler import save_model_checkpoint_peft
from slam_llm.utils.memory_utils import MemoryTrace

logger = logging.getLogger(__name__)


def set_tokenizer_params(tokenizer):
    tokenizer.pad_token_id = 0
    tokenizer.padding_side = "left"


def byte2mb(x):
    return int(x / 2**20)


def _device(train_config, local_rank):
    if not torch.cuda.is_available():
        raise RuntimeError("train(): no CUDA device — the B200 step has no CPU fallback")
    return torch.device("cuda", local_rank if (train_config.enable_fsdp or train_config.enable_ddp) and local_rank is not None else 0)


def _move_batch(batch, device):
    """train_utils.py:100-111 + label rows computed on the host copy (no device sync later)."""
    labels = batch.get("labels")
    if torch.is_tensor(labels) and not labels.is_cuda:
        from slam_llm_b200.engine import SlamStepB200
        batch["_rows"], batch["_targets"] = SlamStepB200.label_rows(labels)
    for key in list(batch.keys()):
        v = batch[key]
        if isinstance(v, torch.Tensor):
            batch[key] = v.to(device, non_blocking=True)
        elif isinstance(v, dict):
            for k2 in v:
                if isinstance(v[k2], torch.Tensor):
                    v[k2] = v[k2].to(device, non_blocking=True)
    return batch


def _joined(loader, active: bool, device=None):
    """DDP `Join` for per-rank iterable loaders of unequal length (the reference wraps its epoch in
    torch.distributed.algorithms.join.Join, utils/train_utils.py:91, because dynamic-frame batching gives every rank its own
    number of batches, datasets/speech_dataset_large.py:80-86).  Every rank keeps stepping until ALL ranks are exhausted; a rank
    that has run dry is handed `None` and must then contribute zero gradients to the collectives of that step
    (`slam_model.shadow_backward`), exactly what DDP's join hook does (gradients still divided by the initial world size)."""
    it = iter(loader)
    if not active:
        yield from it
        return
    flag_dev = device if dist.get_backend() == "nccl" else "cpu"
    flag = torch.zeros(1, dtype=torch.int32, device=flag_dev)
    while True:
        batch = next(it, None)
        flag.fill_(0 if batch is None else 1)
        dist.all_reduce(flag)
        if int(flag.item()) == 0:
            return
        yield batch


def train(model, train_dataloader, eval_dataloader, tokenizer, optimizer, lr_scheduler, gradient_accumulation_steps, train_config, log_config,
          fsdp_config=None, local_rank=None, rank=None):
    """Returns the results dict of the reference (avg_train_prep / avg_train_loss / ... / avg_checkpoint_time)."""
    distributed = train_config.enable_fsdp or train_config.enable_ddp
    world_size = int(os.environ["WORLD_SIZE"]) if distributed else 1
    device = _device(train_config, local_rank)
    is_main = (not distributed) or rank == 0
    use_wandb = bool(log_config.use_wandb)
    if use_wandb:
        import wandb
    train_prep, train_loss, train_acc, val_prep, val_loss, val_acc = [], [], [], [], [], []
    epoch_times, checkpoint_times, results = [], [], {}
    best_val_loss, best_val_acc = float("inf"), 0.0
    dynamic = train_config.batching_strategy == "dynamic"
    raw_model = getattr(model, "module", model)
    if train_config.get("use_fp16", False) and is_main:
        logger.warning("train_config.use_fp16=true: the B200 step computes in bf16 with fp32 accumulation; no GradScaler is created (scale == 1)")
    if train_config.get("run_test_during_validation", False):
        raise NotImplementedError("train_config.run_test_during_validation: the in-loop model.inference(...) decode of the reference "
                                  "(utils/train_utils.py:306-320) is not wired; decode with pipeline/inference_batch.py instead")

    for epoch in range(train_config.num_epochs):
        epoch_start_time = time.perf_counter()
        with MemoryTrace() as memtrace:
            model.train()
            total_loss, total_acc = 0.0, 0.0
            total_length = None if dynamic else len(train_dataloader) // gradient_accumulation_steps
            pbar = tqdm(colour="blue", desc=f"Training Epoch: {epoch+1}", total=total_length, dynamic_ncols=True, disable=not is_main)
            stop = False
            step = -1
            real_steps = 0
            join = dynamic and distributed and world_size > 1 and dist.is_initialized()
            for step, batch in enumerate(_joined(train_dataloader, join, device)):
                last_micro = (step + 1) % gradient_accumulation_steps == 0 or (not dynamic and step == len(train_dataloader) - 1)
                if hasattr(raw_model, "ddp_sync"):
                    raw_model.ddp_sync = last_micro           # DDP no_sync() on accumulation micro-steps
                shadow = batch is None                        # Join: this rank is out of data, the others are not
                real_steps += 0 if shadow else 1
                if shadow:
                    raw_model.shadow_backward()
                    loss, acc = torch.zeros((), device=device), 0.0
                else:
                    batch = _move_batch(batch, device)
                    outputs, *rest = model(**batch)
                    acc = rest[0] if rest else -1
                    loss = outputs.loss / gradient_accumulation_steps
                    acc = acc / gradient_accumulation_steps
                gstep = (epoch * total_length + step) if not dynamic else step + 1
                if use_wandb and step % log_config.log_interval == 0 and is_main:
                    wandb.log({"train_inner/train_inner_loss": loss, "train_inner/train_inner_accuracy": acc}, step=gstep)
                total_loss = total_loss + loss.detach().float()
                total_acc = total_acc + acc
                if not shadow:
                    loss.backward()
                if last_micro:
                    optimizer.step()
                    if lr_scheduler is not None:
                        lr_scheduler.step()
                        current_lr = lr_scheduler.get_last_lr()[0]
                    else:
                        current_lr = optimizer.param_groups[0]["lr"]
                    if current_lr == 0:
                        stop = True
                    if not stop:
                        if use_wandb and step % log_config.log_interval == 0 and is_main:
                            wandb.log({"train_inner/lr": current_lr}, step=gstep)
                        optimizer.zero_grad()
                        pbar.update(1)
                if stop:
                    break
                pbar.set_description(f"Training Epoch: {epoch+1}/{train_config.num_epochs}, step {step}/{len(train_dataloader) if not dynamic else ''} "
                                     f"completed (loss: {loss.detach().float()}, acc: {acc})")

                due = ((epoch * total_length + step + 1) if not dynamic else step + 1) % train_config.validation_interval == 0
                if due and train_config.run_validation:
                    eval_ppl, eval_epoch_loss, *rest = evaluation(model, train_config, eval_dataloader, local_rank, tokenizer)
                    eval_epoch_acc = rest[0] if rest else -1
                    checkpoint_start_time = time.perf_counter()
                    if train_config.save_model and eval_epoch_loss < best_val_loss:
                        checkpoint_name = f"{train_config.model_name}_epoch_{str(epoch+1)}_step_{step+1}"
                        if distributed:
                            dist.barrier()
                        if is_main:
                            logger.info("we are about to save the PEFT modules" if train_config.use_peft else "llm is frozen, we are about to save other parts.")
                            save_model_checkpoint_peft(model, optimizer, rank, train_config, checkpoint_name=checkpoint_name)
                            logger.info(f"PEFT modules are saved in {train_config.output_dir} directory")
                        if distributed:
                            dist.barrier()
                    checkpoint_times.append(time.perf_counter() - checkpoint_start_time)
                    if eval_epoch_loss < best_val_loss:
                        best_val_loss = eval_epoch_loss
                        if is_main:
                            logger.info(f"best eval loss on epoch {epoch+1} is {best_val_loss}")
                    val_loss.append(eval_epoch_loss)
                    val_prep.append(eval_ppl)
                    if rest:
                        if eval_epoch_acc > best_val_acc:
                            best_val_acc = eval_epoch_acc
                            if is_main:
                                logger.info(f"best eval acc on epoch {epoch+1} is {best_val_acc}")
                        val_acc.append(rest[0])
                    else:
                        val_acc.append(-1)
                    if use_wandb and is_main:
                        wandb.log({"valid/val_epoch_loss": eval_epoch_loss, "valid/val_perplexity": eval_ppl, "valid/best_val_loss": best_val_loss,
                                   "valid/val_accuracy": val_acc[-1], "valid/val_best_accuracy": best_val_acc})
                    model.train()
            pbar.close()

        epoch_end_time = time.perf_counter() - epoch_start_time
        epoch_times.append(epoch_end_time)
        if not torch.is_tensor(total_loss):
            total_loss = torch.tensor(float(total_loss), device=device)
        if not torch.is_tensor(total_acc):
            total_acc = torch.tensor(float(total_acc), device=device)
        if world_size > 1 and distributed:
            dist.all_reduce(total_loss, op=dist.ReduceOp.SUM)
            dist.all_reduce(total_acc, op=dist.ReduceOp.SUM)
        n_batches = len(train_dataloader) if not dynamic else max(real_steps, 1)   # (shadow Join steps carry no loss)
        train_epoch_loss = total_loss / n_batches / world_size
        train_epoch_acc = total_acc / n_batches / world_size
        train_perplexity = torch.exp(train_epoch_loss)
        train_prep.append(train_perplexity)
        train_loss.append(train_epoch_loss)
        train_acc.append(train_epoch_acc)
        if use_wandb and is_main:
            wandb.log({"train/train_perplexity": train_perplexity, "train/train_epoch_loss": train_epoch_loss, "train/train_epoch_acc": train_epoch_acc})
        if is_main:
            logger.info(f"Epoch {epoch+1}: train_perplexity={train_perplexity:.4f}, train_epoch_loss={train_epoch_loss:.4f}, epoch time {epoch_end_time}s")
            logger.info(f"Max CUDA memory allocated was {memtrace.peak} GB")
            logger.info(f"Max CUDA memory reserved was {memtrace.max_reserved} GB")
            logger.info(f"Peak active CUDA memory was {memtrace.peak_active_gb} GB")
            logger.info(f"Cuda Malloc retires : {memtrace.cuda_malloc_retires}")
            logger.info(f"CPU Total Peak Memory consumed during the train (max): {memtrace.cpu_peaked + memtrace.cpu_begin} GB")

    results["avg_train_prep"] = sum(train_prep) / len(train_prep)
    results["avg_train_loss"] = sum(train_loss) / len(train_loss)
    results["avg_train_acc"] = sum(train_acc) / len(train_acc)
    if train_config.run_validation and val_loss:
        results["avg_eval_prep"] = sum(val_prep) / len(val_prep)
        results["avg_eval_loss"] = sum(val_loss) / len(val_loss)
        results["avg_eval_acc"] = sum(val_acc) / len(val_acc)
    results["avg_epoch_time"] = sum(epoch_times) / len(epoch_times)
    results["avg_checkpoint_time"] = sum(checkpoint_times) / len(checkpoint_times) if checkpoint_times else 0
    return results


def evaluation(model, train_config, eval_dataloader, local_rank, tokenizer):
    """train_utils.py:396-469: eval loss / accuracy over the eval loader; returns (ppl, loss, acc)."""
    distributed = train_config.enable_fsdp or train_config.enable_ddp
    world_size = int(os.environ["WORLD_SIZE"]) if distributed else 1
    device = _device(train_config, local_rank)
    model.eval()
    eval_preds = []
    eval_loss, eval_acc = 0.0, 0.0
    with MemoryTrace():
        total_length = len(eval_dataloader) if train_config.batching_strategy != "dynamic" else None
        pbar = tqdm(colour="green", desc="Evaluating Epoch", total=total_length, dynamic_ncols=True)
        step = -1
        for step, batch in enumerate(eval_dataloader):
            batch = _move_batch(batch, device)
            with torch.no_grad():
                outputs, *rest = model(**batch)
                acc = rest[0] if rest else -1
                eval_loss = eval_loss + outputs.loss.detach().float()
                eval_acc = eval_acc + acc
            try:
                preds = torch.argmax(outputs.logits, -1)
                eval_preds.extend(tokenizer.batch_decode(preds.detach().cpu().numpy(), skip_special_tokens=True))
            except Exception:
                pass
            pbar.update(1)
            pbar.set_description(f"step: {step+1}/{total_length}, eval_loss: {eval_loss/(step+1):.4f}, eval_acc: {eval_acc/(step+1):.4f}")
    if not torch.is_tensor(eval_loss):
        eval_loss = torch.tensor(float(eval_loss), device=device)
    if not torch.is_tensor(eval_acc):
        eval_acc = torch.tensor(float(eval_acc), device=device)
    if world_size > 1 and distributed:   # (reference precedence slip Q7: all-reduce whenever DDP is on)
        dist.all_reduce(eval_loss, op=dist.ReduceOp.SUM)
        dist.all_reduce(eval_acc, op=dist.ReduceOp.SUM)
    n = len(eval_dataloader) if train_config.batching_strategy != "dynamic" else (step + 1)
    eval_epoch_loss = eval_loss / n / world_size
    eval_epoch_acc = eval_acc / n / world_size
    eval_ppl = torch.exp(eval_epoch_loss)
    if local_rank in (0, None) or not distributed:
        logger.info(f" {eval_ppl=} {eval_epoch_loss=} {eval_epoch_acc=}")
    return eval_ppl, eval_epoch_loss, eval_epoch_acc


def setup():
    """train_utils.py:484-486."""
    dist.init_process_group("nccl")


def setup_environ_flags(rank):
    """train_utils.py:489-498."""
    os.environ["TORCH_SHOW_CPP_STACKTRACES"] = str(1)
    os.environ["NCCL_ASYNC_ERROR_HANDLING"] = str(1)
    if rank == 0:
        logger.info("--> Running with torch dist debug set to detail")


def cleanup():
    dist.destroy_process_group()


def clear_gpu_cache(rank=None):
    if rank == 0:
        logger.info("Clearing GPU cache for all ranks")
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


def _count(module):
    n_train = sum(p.numel() for p in module.parameters() if p.requires_grad) if module is not None else 0
    n_all = sum(p.numel() for p in module.parameters()) if module is not None else 0
    n_all += sum(getattr(m, "num_frozen_params", 0) for m in module.modules()) if module is not None else 0
    return n_all, n_train


def print_model_size(model, config, rank: int = 0) -> None:
    """train_utils.py:520-533."""
    if rank == 0:
        logger.info(f"--> Model {config.model_name}")
        total, trainable = _count(model)
        logger.info(f"--> {config.model_name} has {total / 1e6} Million params ({trainable / 1e6} trainable)\n")


def print_module_size(module, module_name, rank: int = 0) -> None:
    """train_utils.py:536-548."""
    if rank == 0:
        logger.info(f"--> Module {module_name}")
        total, trainable = _count(module)
        logger.info(f"--> {module_name} has {total / 1e6} Million params ({trainable / 1e6} trainable)\n")
