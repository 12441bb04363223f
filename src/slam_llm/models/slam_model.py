"""Model assembly — the drop-in boundary of the hot path (reference: src/slam_llm/models/slam_model.py:21-456).

Same factories and the same `slam_model(encoder, llm, encoder_projector, tokenizer, train_config, model_config, **kw)`
class as the reference, so recipes (examples/asr_librispeech/model/slam_model_asr.py) import and subclass it unchanged;
`forward(**batch) -> (outputs, acc)` runs the whole step on slam_llm_b200 kernels:

  audio_pcm -> log-mel (GPU) | audio_mel -> Whisper encoder -> projector -> fused embedding gather+merge
  -> Llama decoder (frozen bf16 base + fused LoRA tiles) -> lm_head on labelled rows -> fused CE/argmax
  `outputs.loss.backward()` -> dgrad through frozen weights, grads for projector + LoRA into one flat fp32 arena.
"""
from __future__ import annotations

import json
import logging
import os
import types
from typing import Dict, Optional

import torch
import torch.nn as nn

from slam_llm.utils.config_utils import generate_peft_config
from slam_llm.utils.metric import compute_accuracy  # noqa: F401  (re-exported for recipes)
from slam_llm.utils.train_utils import print_model_size, print_module_size
from slam_llm_b200.config import ATTN_LINEARS, LlmCfg, ProjCfg
from slam_llm_b200.engine import LlamaLoRAB200, ProjectorB200, SlamStepB200, TrainableArena

logger = logging.getLogger(__name__)


def _rank(train_config) -> int:
    return int(os.environ["RANK"]) if train_config.enable_fsdp or train_config.enable_ddp else 0


def model_factory(train_config, model_config, **kwargs):
    """slam_model.py:21-51."""
    tokenizer = setup_tokenizer(train_config, model_config, **kwargs)
    encoder = setup_encoder(train_config, model_config, **kwargs)
    llm = setup_llm(train_config, model_config, **kwargs)
    encoder_projector = setup_encoder_projector(train_config, model_config, **kwargs)
    model = slam_model(encoder, llm, encoder_projector, tokenizer, train_config, model_config, **kwargs)
    ckpt_path = kwargs.get("ckpt_path", None)
    if ckpt_path is not None:
        logger.info("loading other parts from: {}".format(ckpt_path))
        ckpt_dict = torch.load(ckpt_path, map_location="cpu")
        model.load_state_dict(ckpt_dict, strict=False)
    print_model_size(model, train_config, _rank(train_config))
    return model, tokenizer


def setup_tokenizer(train_config, model_config, **kwargs):
    """slam_model.py:54-65."""
    from transformers import AutoTokenizer
    tokenizer = AutoTokenizer.from_pretrained(model_config.llm_path)
    tokenizer.pad_token_id = tokenizer.eos_token_id
    return tokenizer


def setup_encoder(train_config, model_config, **kwargs):
    """slam_model.py:68-116 (Whisper path; alternate encoders are out of scope of the B200 hot path)."""
    encoder_list = model_config.encoder_name.split(",") if model_config.encoder_name else []
    if len(encoder_list) == 0:
        return None
    encoder_name = encoder_list[0]
    from slam_llm.models.encoder import WhisperWrappedEncoder, foreign_encoder
    if len(encoder_list) == 1 and foreign_encoder(encoder_name) is not None:
        # a frozen torch encoder from the user's environment (EAT via fairseq, or anything added with register_encoder): it runs as it is,
        # the B200 step takes over at its output (projector, merge, decoder, loss, optimizer)
        if not train_config.freeze_encoder:
            raise NotImplementedError("freeze_encoder=false: foreign encoders are run frozen (no gradient reaches them)")
        encoder = foreign_encoder(encoder_name)[0](model_config)
        for _, param in encoder.named_parameters():
            param.requires_grad = False
        encoder.eval()
        if torch.cuda.is_available():
            encoder.to(torch.device("cuda", torch.cuda.current_device()))
        print_module_size(encoder, encoder_name, _rank(train_config))
        return encoder
    if len(encoder_list) != 1 or encoder_name not in ("whisper", "qwen-audio"):
        raise NotImplementedError(f"encoder_name={model_config.encoder_name!r}: built in are the Whisper encoder (B200 kernels) and EAT (fairseq, foreign); "
                                  "other frozen torch encoders can be plugged in with slam_llm.models.encoder.register_encoder")
    encoder = WhisperWrappedEncoder.load(model_config)
    print_module_size(encoder, encoder_name, _rank(train_config))
    if train_config.freeze_encoder:
        for _, param in encoder.named_parameters():
            param.requires_grad = False
        encoder.eval()
    else:
        raise NotImplementedError("freeze_encoder=false: the B200 path trains projector + LoRA only (encoder frozen, as every asr_* recipe sets)")
    return encoder


# ---------------------------------------------------------------------------------------------------------------------
# LLM module: HF LlamaForCausalLM + peft LoRA surface over LlamaLoRAB200
# ---------------------------------------------------------------------------------------------------------------------
def random_init_allowed(model_config=None) -> bool:
    """Frozen weights are drawn at random ONLY on explicit request (benchmarks / offline tests): `model_config.b200_random_init=true` or
    SLAM_B200_RANDOM_INIT=1.  Otherwise a missing checkpoint is an error, like from_pretrained / whisper.load_model in the reference."""
    if os.environ.get("SLAM_B200_RANDOM_INIT", "0") == "1":
        return True
    return bool(model_config.get("b200_random_init", False)) if model_config is not None else False


def _load_llm_cfg(llm_path: str) -> LlmCfg:
    with open(os.path.join(llm_path, "config.json")) as f:
        c = json.load(f)
    mt = c.get("model_type", "llama")
    if mt not in ("llama", "mistral", "qwen2"):
        raise NotImplementedError(f"model_type={mt!r}: the B200 decoder implements the Llama architecture (+ Qwen2's q/k/v biases and tied embeddings)")
    heads, hidden = c["num_attention_heads"], c["hidden_size"]
    # fields that change the arithmetic and are NOT implemented by the kernels: refuse instead of computing something else silently
    if c.get("rope_scaling") not in (None, {}):
        raise NotImplementedError(f"rope_scaling={c['rope_scaling']!r} (Llama-3.1/3.2 style scaled RoPE) is not implemented by the B200 decoder")
    if c.get("head_dim") not in (None, hidden // heads):
        raise NotImplementedError(f"head_dim={c['head_dim']} != hidden_size/num_attention_heads={hidden // heads}")
    if c.get("attention_bias", False) or c.get("mlp_bias", False):
        raise NotImplementedError("attention_bias / mlp_bias checkpoints are not implemented by the B200 decoder (Llama linears have no bias)")
    if mt == "qwen2" and c.get("use_sliding_window", False):
        raise NotImplementedError("Qwen2 use_sliding_window=true: the B200 attention kernels are full causal")
    if c.get("hidden_act", "silu") != "silu":
        raise NotImplementedError(f"hidden_act={c['hidden_act']!r}: the fused MLP implements SwiGLU (silu)")
    if c.get("sliding_window") not in (None, 0) and mt == "mistral":
        raise NotImplementedError(f"sliding_window={c['sliding_window']}: the B200 attention kernels are full causal")
    return LlmCfg(vocab=c["vocab_size"], d=hidden, layers=c["num_hidden_layers"], heads=heads,
                  kv_heads=c.get("num_key_value_heads", heads), ffn=c["intermediate_size"],
                  rope_theta=float(c.get("rope_theta", 10000.0)), eps=float(c.get("rms_norm_eps", 1e-5)),
                  qkv_bias=(mt == "qwen2"), tie_embeddings=bool(c.get("tie_word_embeddings", False)))


def _load_llm_weights(llm_path: str) -> Optional[Dict[str, torch.Tensor]]:
    """HF checkpoint directory -> {name: tensor}: *.safetensors shards, else pytorch_model*.bin shards.  None when the directory holds no weights."""
    names = sorted(os.listdir(llm_path))
    out: Dict[str, torch.Tensor] = {}
    st = [f for f in names if f.endswith(".safetensors")]
    if st:
        from safetensors.torch import load_file
        for f in st:
            out.update(load_file(os.path.join(llm_path, f)))
    else:
        bins = [f for f in names if f.startswith("pytorch_model") and f.endswith(".bin")]
        for f in bins:
            out.update(torch.load(os.path.join(llm_path, f), map_location="cpu", weights_only=True))
    if not out:
        return None
    return out                                                                   # (tied checkpoints carry no lm_head.weight: the engine shares the table)


class _Embedding(nn.Module):
    """`llm.model.embed_tokens` (probed at slam_model.py:375-380).  The step itself uses the fused gather+merge kernel."""

    def __init__(self, owner):
        super().__init__()
        self._owner = [owner]

    @property
    def weight(self):
        return self._owner[0].b200.embed

    def forward(self, input_ids):
        return torch.nn.functional.embedding(input_ids.to(self.weight.device), self.weight)


class _TrainableEmbedding(nn.Module):
    """`llm.model.embed_tokens` of a full fine-tune: `weight` is the fp32 arena master (registered by register_views).  The fused step reads the
    engine's bf16 copy and writes the embedding gradient itself; a recipe that calls this module directly gets an fp32, autograd-connected lookup."""

    def forward(self, input_ids):
        return torch.nn.functional.embedding(input_ids.to(self.weight.device), self.weight)


def _set_param(root: nn.Module, dotted: str, param: nn.Parameter) -> None:
    parts = dotted.split(".")
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, nn.Module())
        mod = mod._modules[p]
    mod.register_parameter(parts[-1], param)


class LlamaB200ForCausalLM(nn.Module):
    """Frozen Llama decoder (+ optional LoRA adapters under peft-0.6 key names).  Weights are materialised when the
    module is bound to a step arena by slam_model.__init__."""

    def __init__(self, cfg: LlmCfg, llm_path: str, lora_cfg, use_peft: bool, allow_random_init: bool = False, peft_ckpt: Optional[str] = None,
                 train_base: bool = False):
        super().__init__()
        self.train_base = train_base          # train_config.freeze_llm=false: every decoder parameter trains (fp32 masters in the arena)
        self.cfg, self.llm_path, self.lora_cfg, self.use_peft = cfg, llm_path, lora_cfg if use_peft else None, use_peft
        self.allow_random_init, self.peft_ckpt = allow_random_init, peft_ckpt
        self.b200: Optional[LlamaLoRAB200] = None
        object.__setattr__(self, "_step", None)
        D, Dkv, F, V, L = cfg.d, cfg.dkv, cfg.ffn, cfg.vocab, cfg.layers
        self.num_frozen_params = 2 * V * D + L * (2 * D * D + 2 * D * Dkv + 3 * D * F + 2 * D) + D

    # PeftModel.__getattr__ forwards unknown attributes to base_model; `llm.model` must reach the causal-LM wrapper
    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            if name == "model":
                mods = self.__dict__.get("_modules", {})
                if "base_model" in mods:
                    return mods["base_model"]._modules["model"]
            if name == "lm_head":                     # frozen / tied head: an object with `.weight` [vocab, hidden] (the engine's bf16 table)
                b200 = self.__dict__.get("b200")
                w = b200.lm_head if b200 is not None else torch.empty(self.cfg.vocab, self.cfg.d, device="meta")
                return types.SimpleNamespace(weight=w)
            raise

    def bind(self, arena: TrainableArena, device, seed: int = 42) -> None:
        weights = _load_llm_weights(self.llm_path)
        if weights is None:
            if not self.allow_random_init:
                raise FileNotFoundError(f"no *.safetensors / pytorch_model*.bin under llm_path={self.llm_path!r} (set model_config.b200_random_init=true "
                                        "or SLAM_B200_RANDOM_INIT=1 to benchmark with random frozen weights)")
            logger.warning(f"no weights under {self.llm_path}: RANDOM-INIT {self.cfg} (b200_random_init)")
        if weights is not None and "lm_head.weight" not in weights and not self.cfg.tie_embeddings:
            weights["lm_head.weight"] = weights["model.embed_tokens.weight"]      # checkpoint saved with tied weights but an untied config
        if weights is not None:
            g = torch.Generator().manual_seed(seed + 7)
            for k in ("model.embed_tokens.weight", "lm_head.weight"):             # grown vocabulary: new rows ~ N(0, 0.02) (HF _init_weights)
                if k in weights and weights[k].shape[0] < self.cfg.vocab:
                    extra = torch.randn(self.cfg.vocab - weights[k].shape[0], weights[k].shape[1], generator=g) * 0.02
                    weights[k] = torch.cat([weights[k].float(), extra], 0)
        self.b200 = LlamaLoRAB200(self.cfg, self.lora_cfg, arena, device, weights, seed=seed + 1, train_base=self.train_base)

    def register_views(self, arena: TrainableArena) -> None:
        """Expose the embedding module and the LoRA adapters under the reference / peft key names."""
        causal = nn.Module()          # LlamaForCausalLM
        inner = nn.Module()           # LlamaModel
        inner.add_module("embed_tokens", _TrainableEmbedding() if self.train_base else _Embedding(self))
        causal.add_module("model", inner)
        if self.use_peft:
            lora_model = nn.Module()  # peft LoraModel
            lora_model.add_module("model", causal)
            self.add_module("base_model", lora_model)
            for key, view in self.b200.lora_state().items():
                _set_param(self, key[len("llm."):], nn.Parameter(view, requires_grad=True))
        else:
            self.add_module("model", inner)       # un-wrapped: this module IS the HF causal LM (`llm.model` = LlamaModel, `llm.lm_head`)
        if self.train_base:           # full fine-tune: HF parameter names over the fp32 arena masters (`llm.model.layers.N...`, `llm.lm_head.weight`)
            for key, view in self.b200.base_state().items():
                _set_param(self, key[len("llm."):], nn.Parameter(view, requires_grad=True))

    # ---- HF surface the s2s recipe touches (examples/s2s/model/slam_model_s2s.py:146-151)
    # (`self.llm.lm_head.weight.size(0)` resolves through __getattr__ below when no trainable lm_head parameter is registered)
    def resize_token_embeddings(self, new_num_tokens: int):
        """The B200 decoder sizes its tables at construction: setup_llm reads `model_config.vocab_config.total_vocabsize` (the value the s2s recipe
        passes here) and grows the vocabulary up front, new rows ~ N(0, 0.02) like HF's _init_weights.  A later resize is therefore a no-op
        check."""
        if int(new_num_tokens) != self.cfg.vocab:
            raise NotImplementedError(f"resize_token_embeddings({new_num_tokens}) after construction (vocab {self.cfg.vocab}): set "
                                      "model_config.vocab_config.total_vocabsize so that the decoder is built at the final size")
        return self

    def print_trainable_parameters(self):
        trainable = sum(p.numel() for p in self.parameters() if p.requires_grad)
        total = trainable + self.num_frozen_params
        logger.info(f"trainable params: {trainable:,d} || all params: {total:,d} || trainable%: {100 * trainable / total:.4f}")

    def forward(self, inputs_embeds=None, attention_mask=None, labels=None, **kw):
        if self._step is None:
            raise RuntimeError("LLM is not bound to a B200 step yet (construct slam_model first); no CPU fallback")
        return self._step.llm_forward(inputs_embeds, attention_mask, labels)

    @torch.no_grad()
    def generate(self, inputs_embeds=None, attention_mask=None, max_new_tokens=200, num_beams=4, do_sample=False, min_length=1, top_p=1.0,
                 repetition_penalty=1.0, length_penalty=1.0, temperature=1.0, bos_token_id=None, eos_token_id=None, pad_token_id=None, **kw):
        """`llm.generate(inputs_embeds=..., ...)` as slam_model.generate calls it (slam_model.py:439-454): greedy / beam search / sampling control
        flow in slam_llm_b200.generation (transformers v4.35.2 semantics), next-token logits from the B200 decoder with a KV cache (prefill
        once, then one token per sequence per step; beam re-ordering index-selects the cache rows)."""
        from slam_llm_b200 import generation
        if self._step is None:
            raise RuntimeError("LLM is not bound to a B200 step yet (construct slam_model first); no CPU fallback")
        if inputs_embeds is None:
            raise NotImplementedError("generate() is driven by inputs_embeds, as slam_model.generate does")
        eng = self._step.b200
        dev = eng.device
        prompt = inputs_embeds.to(dev, torch.bfloat16)
        B, S0, _ = prompt.shape
        mask = torch.ones(B, S0, dtype=torch.uint8, device=dev) if attention_mask is None else attention_mask.to(dev).to(torch.uint8)
        session = {}

        def next_logits(tokens, beam_src):
            """KV-cache decode: the prompt is run once (prefill, expanded to one copy per beam on the first beam step), afterwards each call feeds
            the newest token of every sequence; beam_src re-orders the cache rows.  SLAM_DECODE_NO_CACHE=1 re-runs the whole sequence instead."""
            n, t = tokens.shape
            reps = n // B
            if os.environ.get("SLAM_DECODE_NO_CACHE", "0") == "1":
                x = prompt.repeat_interleave(reps, dim=0) if reps > 1 else prompt
                m = mask.repeat_interleave(reps, dim=0) if reps > 1 else mask
                if t > 0:
                    x = torch.cat([x, torch.nn.functional.embedding(tokens.to(dev), eng.llm.embed)], dim=1)
                    m = torch.cat([m, torch.ones(n, t, dtype=torch.uint8, device=dev)], dim=1)
                return eng.decoder_last_logits(x, m)
            if t == 0:
                logits, session["state"] = eng.decode_prefill(prompt, mask)
                if reps > 1:                                   # beams start as copies of the prompt (their scores differ, not their state)
                    rep = torch.arange(B, device=dev).repeat_interleave(reps)
                    st = session["state"]
                    st["cache"] = [(K.index_select(0, rep), V.index_select(0, rep)) for K, V in st["cache"]]
                    st["mask"] = st["mask"].index_select(0, rep)
                    logits = logits.index_select(0, rep)
                return logits
            return eng.decode_next(tokens[:, -1], session["state"], beam_src)

        return generation.generate(next_logits, B, max_new_tokens=max_new_tokens, num_beams=num_beams, do_sample=do_sample, min_length=min_length,
                                   top_p=top_p, repetition_penalty=repetition_penalty, length_penalty=length_penalty, temperature=temperature,
                                   eos_token_id=eos_token_id, pad_token_id=pad_token_id).to(dev)


def setup_llm(train_config, model_config, **kwargs):
    """slam_model.py:118-221: frozen base LLM + LoRA adapters from train_config.peft_config."""
    if train_config.quantization:
        raise NotImplementedError("8-bit quantised loading is out of scope of the B200 path")
    train_base = not train_config.freeze_llm                        # full fine-tune (examples/s2s): slam_model.py:205-208 not taken
    if train_base and (train_config.use_peft or kwargs.get("peft_ckpt", None)):
        raise NotImplementedError("freeze_llm=false together with peft adapters is not implemented (the reference recipes use one or the other)")
    cfg = _load_llm_cfg(model_config.llm_path)
    peft_ckpt = kwargs.get("peft_ckpt", None)
    if peft_ckpt:                                                    # slam_model.py:210-213: PeftModel.from_pretrained(model, peft_ckpt, is_trainable=True)
        logger.info("loading peft_ckpt from: {}".format(peft_ckpt))
        lora_cfg = _peft_dir_config(peft_ckpt)
    else:
        lora_cfg = generate_peft_config(train_config) if train_config.use_peft else None
    use_peft = bool(peft_ckpt) or bool(train_config.use_peft)
    vc = model_config.get("vocab_config", None)
    total_vocab = int(vc.get("total_vocabsize", 0)) if vc is not None else 0
    if total_vocab and total_vocab != cfg.vocab:                    # s2s: text vocab + code_layer audio vocabularies (slam_model_s2s.py:146-151)
        if total_vocab < cfg.vocab:
            raise ValueError(f"vocab_config.total_vocabsize={total_vocab} is smaller than the checkpoint vocabulary {cfg.vocab}")
        logger.info(f"growing the LLM vocabulary {cfg.vocab} -> {total_vocab} (model_config.vocab_config.total_vocabsize)")
        cfg.vocab_ckpt, cfg.vocab = cfg.vocab, total_vocab
    model = LlamaB200ForCausalLM(cfg, model_config.llm_path, lora_cfg, use_peft, allow_random_init=random_init_allowed(model_config),
                                 peft_ckpt=peft_ckpt or None, train_base=train_base)
    print_module_size(model, model_config.llm_name, _rank(train_config))
    model.eval()
    if train_config.use_peft and not peft_ckpt:
        logger.info("setup peft...")
    return model


def _peft_dir_config(peft_dir: str):
    """adapter_config.json of a peft checkpoint directory -> the LoRA config the engine needs (r, alpha, targets, dropout)."""
    from slam_llm_b200.config import LoraCfg
    with open(os.path.join(peft_dir, "adapter_config.json")) as f:
        c = json.load(f)
    if c.get("peft_type", "LORA") != "LORA":
        raise NotImplementedError(f"peft_type={c.get('peft_type')!r}: only LoRA adapters are implemented")
    if c.get("bias", "none") != "none" or c.get("modules_to_save"):
        raise NotImplementedError("peft checkpoints with trainable biases / modules_to_save are not implemented")
    return LoraCfg(int(c["r"]), int(c["lora_alpha"]), tuple(c["target_modules"]), float(c.get("lora_dropout", 0.0)))


def _peft_dir_state(peft_dir: str) -> Dict[str, torch.Tensor]:
    """adapter_model.{safetensors,bin} -> reference state-dict names: peft saves `...q_proj.lora_A.weight` (adapter name stripped);
    the live module calls it `...q_proj.lora_A.default.weight`, under the `llm.` attribute of slam_model."""
    st = os.path.join(peft_dir, "adapter_model.safetensors")
    if os.path.exists(st):
        from safetensors.torch import load_file
        sd = load_file(st)
    else:
        sd = torch.load(os.path.join(peft_dir, "adapter_model.bin"), map_location="cpu", weights_only=True)
    out = {}
    for k, v in sd.items():
        k = k.replace(".lora_A.weight", ".lora_A.default.weight").replace(".lora_B.weight", ".lora_B.default.weight")
        out["llm." + k] = v
    return out


def setup_encoder_projector(train_config, model_config, **kwargs):
    """slam_model.py:223-236."""
    if model_config.encoder_projector == "linear":
        from slam_llm.models.projector import EncoderProjectorConcat
        encoder_projector = EncoderProjectorConcat(model_config)
    elif model_config.encoder_projector == "cov1d-linear":
        from slam_llm.models.projector import EncoderProjectorCov1d
        encoder_projector = EncoderProjectorCov1d(model_config)
    elif model_config.encoder_projector == "q-former":
        from slam_llm.models.projector import EncoderProjectorQFormer
        encoder_projector = EncoderProjectorQFormer(model_config)
    else:
        return None
    print_module_size(encoder_projector, model_config.encoder_projector, _rank(train_config))
    return encoder_projector


# ---------------------------------------------------------------------------------------------------------------------
class _StepLoss(torch.autograd.Function):
    """Connects the device-computed loss to autograd: `loss.backward()` (utils/train_utils.py:130,152) runs the B200
    backward, which writes the projector/LoRA gradients straight into the flat arena (p.grad are views of it)."""

    @staticmethod
    def forward(ctx, anchor, loss, owner):
        ctx.owner = owner
        return loss.clone()

    @staticmethod
    def backward(ctx, grad_out):
        ctx.owner._backward(grad_out)
        return None, None, None


class _LlmLoss(torch.autograd.Function):
    """Decoder + lm_head + CE as ONE autograd node over inputs_embeds (see slam_model.llm_forward)."""

    @staticmethod
    def forward(ctx, inputs_embeds, owner, key_mask, labels, full):
        eng = owner.b200
        eng.begin_decoder_pass(True)
        x = inputs_embeds.detach().to(eng.device, torch.bfloat16).contiguous()
        loss, acc, logits = eng.decoder_loss(x, key_mask, labels, train=True, full_logits=full)
        ctx.owner, ctx.in_dtype = owner, inputs_embeds.dtype
        ctx.set_materialize_grads(False)
        if logits is None:
            logits = torch.empty(0, device=eng.device)
            ctx.mark_non_differentiable(logits)                    # labelled-rows mode: only the loss is differentiable
        return loss.clone(), logits                                # full-logits mode: BOTH outputs are differentiable (s2s builds its own loss)

    @staticmethod
    def backward(ctx, grad_loss, grad_logits):
        owner = ctx.owner
        eng = owner.b200
        eng.backward_begin()
        if grad_loss is None:                                      # the recipe ignored outputs.loss: no CE term
            grad_loss = torch.zeros((), device=eng.device)
        dx = eng.decoder_backward(grad_loss, grad_logits=grad_logits)
        torch.autograd.Variable._execution_engine.queue_callback(owner._finish_split_backward)
        return dx.to(ctx.in_dtype), None, None, None, None


class _Outputs(types.SimpleNamespace):
    pass


class slam_model(nn.Module):
    def __init__(self, encoder: nn.Module, llm: nn.Module, encoder_projector: nn.Module, tokenizer, train_config, model_config, **kwargs):
        super().__init__()
        self.encoder = encoder
        self.llm = llm
        self.encoder_projector = encoder_projector
        self.tokenizer = tokenizer
        self.metric = kwargs.get("metric", "acc")
        self.train_config = train_config
        self.model_config = model_config
        self.dataset_config = kwargs.get("dataset_config", None)   # reference quirk Q4
        if not torch.cuda.is_available():
            raise RuntimeError("slam_model needs a CUDA (B200) device: the hot path has no CPU fallback")
        if encoder is None or encoder_projector is None or not isinstance(llm, LlamaB200ForCausalLM):
            raise NotImplementedError("the B200 step needs a Whisper encoder, a linear projector and a Llama-architecture LLM")
        device = torch.device("cuda", torch.cuda.current_device())
        arena = TrainableArena()
        proj_cfg = ProjCfg(encoder_projector.kind, encoder_projector.k, encoder_projector.linear1.out_features)
        self._foreign_call = None
        if getattr(encoder, "b200", None) is None:                  # frozen torch encoder (slam_llm.models.encoder.register_encoder)
            from slam_llm.models.encoder import foreign_encoder
            from slam_llm_b200.config import EncoderCfg
            entry = foreign_encoder(model_config.encoder_name)
            if entry is None:
                raise NotImplementedError(f"encoder of type {type(encoder).__name__} is neither the B200 Whisper encoder nor a registered foreign encoder")
            self._foreign_call = entry[1]
            enc_cfg = EncoderCfg(n_mels=0, n_ctx=0, d=int(model_config.encoder_dim), heads=1, layers=0)
        else:
            enc_cfg = encoder.b200.cfg
        eng_proj = ProjectorB200(enc_cfg, llm.cfg, proj_cfg, arena)
        seed = int(train_config.get("seed", 42))                    # adapters + LoRA dropout follow train_config.seed like the reference's global RNG
        llm.bind(arena, device, seed=seed)
        arena.finalize(device)
        encoder_projector.bind(eng_proj, arena)
        llm.b200.init_lora(None, seed=seed + 2)                     # peft init: A kaiming-uniform, B zeros
        llm.b200.init_base()                                        # full fine-tune: checkpoint weights -> fp32 arena masters
        llm.register_views(arena)
        if llm.peft_ckpt:
            sd = _peft_dir_state(llm.peft_ckpt)
            mine = llm.b200.lora_state()
            unknown = sorted(set(sd) - set(mine))
            if unknown:
                raise KeyError(f"peft_ckpt holds adapters this model does not have: {unknown[:4]} ...")
            for k, v in sd.items():
                mine[k].copy_(v.to(device, torch.float32))
        self.b200 = SlamStepB200.from_parts(getattr(encoder, "b200", None), eng_proj, llm.b200, arena, device, enc_cfg=enc_cfg)
        object.__setattr__(llm, "_step", self)       # plain attribute: registering the parent as a sub-module would create a cycle
        object.__setattr__(encoder_projector, "_step", self.b200)
        arena.param.requires_grad_(True)
        self._grad_views_set = False
        self.ddp_world_size = 1
        self.ddp_sync = True

    # nn.Module.to()/cuda() would re-materialise the arena views as independent tensors: the step is already on the GPU
    def _apply(self, fn, recurse=True):
        return self

    # ---- gradient plumbing
    def _bind_grad_views(self):
        grads = self.b200.trainable_state("grad")
        for name, p in self.named_parameters():
            if p.requires_grad and name in grads:
                p.grad = grads[name]

    def _backward(self, grad_out):
        self.b200.backward(grad_out)
        if self.ddp_world_size > 1 and self.ddp_sync:
            # DDP: the one data-path collective (finetune.py:181-184).  In deferred mode (train_config.b200_overlap_allreduce, default on for
            # DDP) it is asynchronous: FlatAdamW.step() records the update and the next forward applies it after the frozen encoder has been
            # launched, so the all-reduce and the wait for the slowest rank overlap the next step's front end.
            self.b200.allreduce_grads(async_op=self.b200.defer_update)
        self._bind_grad_views()

    def state_dict(self, *args, **kwargs):
        self.b200.flush_update()                                   # a deferred optimizer step must land before parameters are read
        return super().state_dict(*args, **kwargs)

    def shadow_backward(self):
        """DDP `Join` (utils/train_utils.py:91): this rank has no batch for the micro-step while other ranks still train - contribute
        zero gradients to the step's all-reduce so that every replica applies the same update."""
        self.b200.flush_update()
        if self.b200.micro_steps == 0:
            self.b200.arena.grad.zero_()
        self.b200.micro_steps += 1
        if self.ddp_world_size > 1 and self.ddp_sync:
            self.b200.allreduce_grads(async_op=self.b200.defer_update)
        self._bind_grad_views()

    # ---- decoder entry used by recipes that override forward() and call self.llm(...) themselves (slam_model.py:400)
    def llm_forward(self, inputs_embeds, attention_mask, labels):
        """`self.llm(inputs_embeds=..., attention_mask=..., labels=...)` -> object with .loss / .logits, autograd-connected to inputs_embeds:
        `loss.backward()` runs the B200 decoder backward (LoRA gradients into the arena) and hands d loss / d inputs_embeds to whatever torch
        graph produced the embeddings (typically the projector module, itself an autograd node over the arena).  Without labels: eval logits."""
        eng = self.b200
        if inputs_embeds is None:
            raise NotImplementedError("the B200 decoder is driven by inputs_embeds (as slam_model.forward does); input_ids-only calls are not implemented")
        dev = eng.device
        B, S = inputs_embeds.shape[:2]
        key_mask = (torch.ones(B, S, dtype=torch.uint8, device=dev) if attention_mask is None else attention_mask.to(dev).to(torch.uint8).contiguous())
        train = torch.is_grad_enabled() and labels is not None
        eng.lora_dropout_enabled = self.training
        if labels is None:
            eng.begin_decoder_pass(False)
            x = inputs_embeds.detach().to(dev, torch.bfloat16).contiguous()
            dummy = torch.full((B, S), -100, dtype=torch.int64, device=dev)
            _, _, logits = eng.decoder_loss(x, key_mask, dummy, train=False, full_logits=True)
            return _Outputs(loss=None, logits=logits)
        full = (not train) or bool(self.train_config.get("b200_full_logits", False))
        if not train:
            eng.begin_decoder_pass(False)
            loss, acc, logits = eng.decoder_loss(inputs_embeds.detach().to(dev, torch.bfloat16).contiguous(), key_mask, labels, train=False, full_logits=full)
            return _Outputs(loss=loss, logits=logits)
        for p in self.parameters():                                # arena views are re-bound after the backward (_finish_split_backward)
            p.grad = None
        loss, logits = _LlmLoss.apply(inputs_embeds, self, key_mask, labels, full)
        return _Outputs(loss=loss, logits=logits if logits.numel() else None)

    def _finish_split_backward(self):
        """Runs once the autograd pass that contained _LlmLoss.backward has finished (projector / embedding nodes included)."""
        grads = self.b200.trainable_state("grad")
        for name, p in self.named_parameters():                    # a torch node (e.g. the trainable embedding lookup) left a gradient tensor of
            g = p.grad                                             # its own: fold it into the arena, which is what the optimizer reads
            if g is not None and name in grads and g.data_ptr() != grads[name].data_ptr():
                grads[name].add_(g.to(grads[name].dtype).view_as(grads[name]))
        self.b200.backward_end()
        if self.ddp_world_size > 1 and self.ddp_sync:
            self.b200.allreduce_grads(async_op=self.b200.defer_update)
        self._bind_grad_views()

    # ---- forward (slam_model.py:283-407)
    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None, labels=None,
                use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=None, **kwargs):
        audio_mel = kwargs.get("audio_mel", None)
        audio_pcm = kwargs.get("audio_pcm", None)
        modality_mask = kwargs.get("modality_mask", None)
        if self._foreign_call is not None:
            return self._forward_foreign(input_ids, attention_mask, labels, modality_mask, kwargs)
        if audio_mel is None and audio_pcm is None:
            raise NotImplementedError("the B200 step needs audio_mel or audio_pcm in the batch (Whisper recipes)")
        if self.train_config.freeze_encoder:
            self.encoder.eval()
        batch = dict(input_ids=input_ids, labels=labels, attention_mask=attention_mask, modality_mask=modality_mask)
        if audio_mel is not None:
            batch["audio_mel"] = audio_mel
        else:
            batch["audio_pcm"] = audio_pcm
        for k in ("_rows", "_targets", "audio_pcm_lengths"):
            if kwargs.get(k, None) is not None:
                batch[k] = kwargs[k]
        if kwargs.get("inference_mode", False):
            return self._inputs_embeds(batch), attention_mask
        train = torch.is_grad_enabled() and labels is not None
        self.b200.lora_dropout_enabled = self.training        # model.train() re-enables LoRA dropout each epoch (reference quirk Q6)
        full = (not train) or bool(self.train_config.get("b200_full_logits", False))
        loss, acc, logits = self.b200.forward(batch, train=train, full_logits=full)
        if train:
            loss = _StepLoss.apply(self.b200.arena.param, loss, self)
        outputs = _Outputs(loss=loss, logits=logits)
        if not self.metric:
            acc = -1
        return outputs, acc

    def _forward_foreign(self, input_ids, attention_mask, labels, modality_mask, kwargs):
        """slam_model.py:314-407 with a frozen torch encoder: its features (computed here, no grad) enter the B200 step at the projector."""
        eng = self.b200
        self.encoder.eval()
        with torch.no_grad():
            feats = self._foreign_call(self.encoder, {k: (v.to(eng.device) if torch.is_tensor(v) else v) for k, v in kwargs.items()})
        feats = feats.to(eng.device, torch.bfloat16).contiguous()
        batch = dict(input_ids=input_ids, labels=labels, attention_mask=attention_mask, modality_mask=modality_mask)
        for k in ("_rows", "_targets"):
            if kwargs.get(k, None) is not None:
                batch[k] = kwargs[k]
        if kwargs.get("inference_mode", False):
            from slam_llm_b200 import ops
            eng.begin_decoder_pass(False)
            aud = eng.projector.forward(feats, save=False)
            return ops.embed_merge(input_ids.to(eng.device).contiguous(), modality_mask.to(eng.device).to(torch.uint8).contiguous(), aud, eng.llm.embed), attention_mask
        train = torch.is_grad_enabled() and labels is not None
        eng.lora_dropout_enabled = self.training
        full = (not train) or bool(self.train_config.get("b200_full_logits", False))
        eng.flush_update()
        loss, acc, logits = eng.forward_rest(batch, feats, train=train, full_logits=full)
        if train:
            loss = _StepLoss.apply(eng.arena.param, loss, self)
        return _Outputs(loss=loss, logits=logits), (acc if self.metric else -1)

    def _inputs_embeds(self, batch):
        from slam_llm_b200 import ops
        dev = self.b200.device
        self.b200.begin_decoder_pass(False)                         # deferred update / fp32 masters -> bf16 operands before anything reads them
        mel = batch.get("audio_mel")
        mel = self.b200.log_mel(batch["audio_pcm"].to(dev, torch.float32), batch.get("audio_pcm_lengths")) if mel is None else mel.to(dev, torch.float32)
        aud = self.b200.projector.forward(self.b200.encoder.forward(mel), save=False)
        return ops.embed_merge(batch["input_ids"].to(dev).contiguous(), batch["modality_mask"].to(dev).to(torch.uint8).contiguous(), aud,
                               self.b200.llm.embed)

    @torch.no_grad()
    def generate(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None, labels=None, use_cache=None,
                 output_attentions=None, output_hidden_states=None, return_dict=None, **kwargs):
        """slam_model.py:409-456: inputs_embeds through forward(inference_mode=True), then llm.generate with the reference's defaults."""
        kwargs["inference_mode"] = True
        inputs_embeds, attention_mask = self.forward(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                                                     past_key_values=past_key_values, inputs_embeds=inputs_embeds, labels=labels, use_cache=use_cache,
                                                     output_attentions=output_attentions, output_hidden_states=output_hidden_states,
                                                     return_dict=return_dict, **kwargs)
        tok = self.tokenizer
        return self.llm.generate(inputs_embeds=inputs_embeds, max_new_tokens=kwargs.get("max_new_tokens", 200), num_beams=kwargs.get("num_beams", 4),
                                 do_sample=kwargs.get("do_sample", False), min_length=kwargs.get("min_length", 1), top_p=kwargs.get("top_p", 1.0),
                                 repetition_penalty=kwargs.get("repetition_penalty", 1.0), length_penalty=kwargs.get("length_penalty", 1.0),
                                 temperature=kwargs.get("temperature", 1.0), attention_mask=attention_mask,
                                 bos_token_id=getattr(tok, "bos_token_id", None), eos_token_id=getattr(tok, "eos_token_id", None),
                                 pad_token_id=getattr(tok, "pad_token_id", None))
