"""CUDA-graph replay of the training step (static shapes): ~1130 kernel launches per step become two graph launches.

Why: every kernel of the step is enqueued from Python through ctypes (pointer marshalling, a fresh output tensor, 2-4 host-encoded
CUtensorMaps per GEMM).  One GPU hides that behind its own 60 ms of work, but eight ranks sharing the host cores of one box do not
(SCALE_r01: 0.879 weak-scaling efficiency at 8 GPUs with a sub-millisecond collective).  The step is split where the deferred
data-parallel update has to land:

    G_front  log-mel + Whisper encoder                  (frozen: independent of the trainables)
    -- eager: wait for the gradient all-reduce of the previous step, AdamW (one launch, host-computed bias corrections / LR)
    G_rest   LoRA packing, projector, merge, decoder forward, lm_head + CE, full backward into the flat gradient arena
    -- eager: NCCL all-reduce of the arena (async in deferred mode), optimizer_step() records the update

Kernels are captured through torch.cuda.graph (stream capture of the ctypes launches on torch's capture stream, intermediates in the
graph's private pool); TMA descriptors are kernel parameters, so they are baked into the graph nodes together with the (stable) addresses.
Inputs are copied into static buffers before each replay.  Shapes are static per graph: callers keep one GraphedTrainStep per shape
bucket (the BASELINE jsonl recipe pads every utterance to 30 s, so only S varies with the text lengths).

Not graphed (the eager path is used instead): LoRA dropout > 0 in training (the mask seed is a launch parameter), gradient accumulation
(micro_steps > 0), batches without precomputed label rows (`_rows` / `_targets` — their count is a shape).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

INPUT_KEYS = ("input_ids", "labels", "attention_mask", "modality_mask", "audio_pcm", "audio_mel", "audio_pcm_lengths", "_rows", "_targets")


def signature(batch: Dict[str, torch.Tensor]):
    """Shape bucket of a batch: every tensor shape that enters the step (incl. the number of label rows)."""
    return tuple((k, tuple(batch[k].shape), str(batch[k].dtype)) for k in INPUT_KEYS if batch.get(k) is not None)


class GraphedTrainStep:
    def __init__(self, engine, example_batch: Dict[str, torch.Tensor], warmup: int = 2):
        eng = self.eng = engine
        if eng.llm.dropout_p > 0.0 and eng.lora_dropout_enabled:
            raise ValueError("GraphedTrainStep: LoRA dropout is active (the mask seed is a launch parameter); use the eager step")
        if "_rows" not in example_batch or "_targets" not in example_batch:
            raise ValueError("GraphedTrainStep needs precomputed label rows (`_rows`, `_targets`: SlamStepB200.label_rows on the host copy)")
        dev = eng.device
        self.sig = signature(example_batch)
        self.static = {k: example_batch[k].to(dev).clone() for k in INPUT_KEYS if example_batch.get(k) is not None}
        eng.flush_update()
        keep = (eng.arena.param.detach().clone(), eng.arena.exp_avg.clone(), eng.arena.exp_avg_sq.clone(), eng.arena.step_count, eng.llm.dropout_step)
        # warm-up on a side stream (allocator pools, rope tables, cudaFuncSetAttribute, scratch pools reach their steady size)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                enc = eng.forward_front(self.static)
                eng.forward_rest(self.static, enc, train=True)
                eng.micro_steps = 0
                eng.backward()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from . import engine as _engine, ops
        pool = torch.cuda.graph_pool_handle()
        self.g_front, self.g_rest = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        n0 = ops.launch_count()
        with torch.cuda.graph(self.g_front, pool=pool, capture_error_mode="thread_local"):
            self.enc_out = eng.forward_front(self.static)
        eng.micro_steps = 0
        with torch.cuda.graph(self.g_rest, pool=pool, capture_error_mode="thread_local"):
            self.loss, self.acc, _ = eng.forward_rest(self.static, self.enc_out, train=True)
            eng.backward()
        eng.micro_steps = 0
        self.kernels_per_step = ops.launch_count() - n0            # libslam_b200 kernels recorded into the two graphs (replayed every step)
        self._keep = [_engine._THIN_POOL.buf]                       # scratch the captured kernels address: must outlive any later re-allocation
        # the warm-up steps must not count as training: restore parameters / moments (backward never touches them, but be explicit)
        with torch.no_grad():
            eng.arena.param.copy_(keep[0]); eng.arena.exp_avg.copy_(keep[1]); eng.arena.exp_avg_sq.copy_(keep[2])
        eng.arena.step_count, eng.llm.dropout_step = keep[3], keep[4]
        torch.cuda.synchronize(dev)

    def load(self, batch: Dict[str, torch.Tensor]) -> None:
        if signature(batch) != self.sig:
            raise ValueError(f"GraphedTrainStep: batch shapes {signature(batch)} differ from the captured bucket {self.sig}")
        for k, buf in self.static.items():
            buf.copy_(batch[k], non_blocking=True)

    def train_step(self, batch: Optional[Dict[str, torch.Tensor]], lr: float = 1e-4, weight_decay: float = 0.0, world_size: int = 1):
        """Same contract as SlamStepB200.train_step; batch=None replays on whatever the static buffers hold."""
        eng = self.eng
        if eng.micro_steps != 0:
            raise RuntimeError("GraphedTrainStep does not support gradient accumulation (micro_steps > 0)")
        if batch is not None:
            self.load(batch)
        self.g_front.replay()
        eng.flush_update()                       # deferred mode: all-reduce wait + AdamW of the previous step, behind the frozen front end
        self.g_rest.replay()
        eng.micro_steps = 1
        if world_size > 1:
            eng.allreduce_grads(async_op=eng.defer_update)
        eng.optimizer_step(lr, weight_decay, grad_div=float(world_size))
        return self.loss, self.acc
