"""The SLAM-LLM training step on B200: Whisper encoder -> projector -> merge -> Llama(+LoRA) -> CE,
backward through projector/LoRA only, flat-buffer AdamW — every FLOP in libslam_b200.so kernels.

Mirrors (file:line under /root/reference):
  encoder      src/slam_llm/models/encoder.py:13-30 (+ openai-whisper AudioEncoder modules)
  projector    src/slam_llm/models/projector.py:5-27
  merge        src/slam_llm/models/slam_model.py:370-392
  decoder      transformers LlamaForCausalLM + peft LoRA, called at src/slam_llm/models/slam_model.py:400
  loss / acc   HF loss block + src/slam_llm/utils/metric.py:3-20 (slam_model.py:402-405)
  optimizer    src/slam_llm/pipeline/finetune.py:247-251

Data layout in HBM (180 GB/GPU): frozen weights live in bf16 twice — W [out,in] for the forward GEMM and
W^T [in,out] for the dgrad GEMM — so every GEMM is the same K-major tcgen05 kernel with TMA-loaded operands.
All trainables (projector + LoRA) live in ONE flat fp32 arena (param / grad / exp_avg / exp_avg_sq) so the
data-parallel exchange is one NCCL all-reduce and the optimizer is one kernel.
"""
from __future__ import annotations

import functools
import math
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .config import ATTN_LINEARS, MLP_LINEARS, EncoderCfg, LlmCfg, LoraCfg, ProjCfg, linear_shape
from .frontend import mel_filterbank, sinusoids

BF16, F32 = torch.bfloat16, torch.float32


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class _ZeroPool:
    """fp32 scratch for the split-K thin products of one forward or backward pass: ONE memset per pass instead of one fill kernel per
    product (64 of them per step); slices are handed out in call order and the pool grows to the largest pass seen."""

    def __init__(self) -> None:
        self.buf: Optional[torch.Tensor] = None
        self.used = 0
        self.need = 0

    def begin(self, device) -> None:
        if self.buf is None or self.buf.numel() < self.need:
            self.buf = torch.empty(max(self.need, 1), device=device, dtype=F32)
        self.buf.zero_()
        self.used = 0

    def take(self, m: int, n: int, device) -> torch.Tensor:
        k = m * n
        self.need = max(self.need, self.used + k)
        if self.buf is None or self.used + k > self.buf.numel():
            self.used += k
            return torch.zeros((m, n), device=device, dtype=F32)           # first pass / larger batch: the pool is resized at the next begin()
        out = self.buf[self.used: self.used + k].view(m, n)
        self.used += k
        return out


_THIN_POOL = _ZeroPool()
_THIN_CLUSTER = os.environ.get("SLAM_THIN_CLUSTER", "1") != "0"      # A/B switch for the thin-product cluster kernel


def _gemm_few_tiles(a: torch.Tensor, b: torch.Tensor, a2=None, b2=None) -> torch.Tensor:
    """bf16 out = a @ b.T (+ a2 @ b2.T) for products with few output tiles but a long K (LoRA rank-64 products, lm_head dgrad):
    wide outputs (N >= 256) use the GEMM's own tail split (k-slices on idle SMs, deterministic exchange, bf16 epilogue);
    thin ones (rank-64 LoRA products, N <= 64) go to the cluster kernel of the library (8 CTAs per tile split K and reduce through distributed
    shared memory: csrc/gemm_thin.cuh); what is left (a second K segment) splits K with fp32 atomics into a zeroed buffer + one cast."""
    M, K = a.shape
    N = b.shape[0]
    if N >= 256 or (N <= 64 and a2 is None and _THIN_CLUSTER):
        return ops.gemm(a, b, a2=a2, b2=b2)
    tiles = ((M + 127) // 128) * ((N + 255) // 256 if N > 64 else 1)
    nkb = (K + 63) // 64 + ((a2.shape[1] + 63) // 64 if a2 is not None else 0)
    split = min(8, nkb // 4, max(1, 148 // tiles))
    if nkb < 128:      # measured (tools/thin_probe.py, M=1604 N=64): K=4096 16.5 us unsplit vs 22.8 us split-8 + cast; K=6144 22.4 vs 23.8; K=14336 46.1 vs 26.9
        split = 1
    if split <= 1:
        return ops.gemm(a, b, a2=a2, b2=b2, block_n=64 if N <= 64 else 0)          # (explicit 128 x 64 tile: `auto` would pick the cluster kernel)
    return ops.cast_bf16(ops.gemm(a, b, a2=a2, b2=b2, out=_THIN_POOL.take(M, N, a.device), out_f32=True, split_k=split))


_SWAP_AB = os.environ.get("SLAM_SWAP_AB", "1") != "0"
_STATIC_W = os.environ.get("SLAM_STATIC_W", "1") != "0"     # frozen weights are declared to the GEMM (slam_gemm_args.static_operands); 0 = A/B switch


def _swap_ab(tokens: int, features: int) -> bool:
    """Swap-AB for the decoder linears (slam_gemm_args.transpose_out): the weight becomes the 256-row M operand of CTA-pair tiles and the token
    dimension the 192-wide N.  Pays when the token count does not fill 256-row pair tiles (M = 1604: 7 pair tiles = 10.5 % padding the other way
    round) while the feature count does.  Measured at M = 1604 (tools/swap_probe.py, steady state): +1.4 ... +4.0 % over the best non-swapped tile on
    qkv / o / down / d_gate_up / d_qkv."""
    if not _SWAP_AB or features % 256 != 0 or tokens < 512:
        return False
    pad = (-tokens) % 256
    return pad * 25 > tokens                      # more than 4 % of a pair grid would be padding rows


_SWAP_DDOWN = os.environ.get("SLAM_SWAP_DDOWN", "0") != "0"      # swap-AB for the fused-SwiGLU-backward GEMM (d_down): measured +0.3 ... +0.6 ms per step
                                                                  # (profiles/r02_ab_swap_ddown.log), so off; the kernel path stays (tested, ABI-documented)
_PAIR512 = os.environ.get("SLAM_PAIR512", "long")               # "0" never, "long" K >= 8192, "all" whenever the GEMM is one round of 512-row tiles


@functools.lru_cache(maxsize=None)
def _swap_tile(tokens: int, features: int, k: int) -> int:
    """slam_gemm_args.block_n for a swap-AB GEMM: the 512 x 192 pair tile (one accumulator set, 21 % fewer operand bytes per MMA cycle) when the
    whole product is ONE round of such tiles - e.g. 4096 weight rows x 1604 tokens = 8 x 9 = 72 tiles on 74 SM pairs - else 0 (automatic)."""
    if _PAIR512 == "0" or features % 512 != 0 or (_PAIR512 != "all" and k < 8192):
        return 0
    pairs = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count // 2
    return 4000192 if (features // 512) * ((tokens + 191) // 192) <= pairs else 0


def _require_cuda(device) -> torch.device:
    device = torch.device(device)
    if device.type != "cuda" or not torch.cuda.is_available():
        raise RuntimeError("slam_llm_b200 engine needs a CUDA (B200) device; there is no CPU fallback")
    return device


# =====================================================================================================
# flat arena of trainables
# =====================================================================================================
class TrainableArena:
    """One flat fp32 buffer each for parameters, gradients and the two Adam moments."""

    def __init__(self) -> None:
        self._specs: List[Tuple[str, Tuple[int, ...]]] = []
        self._offsets: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
        self.n = 0
        self.param = self.grad = self.exp_avg = self.exp_avg_sq = None
        self.step_count = 0

    def add(self, name: str, shape: Tuple[int, ...]) -> None:
        assert self.param is None and name not in self._offsets
        numel = int(math.prod(shape))
        self._offsets[name] = (self.n, tuple(shape))
        self.n += _round_up(numel, 8)  # keep every tensor 32-byte aligned for vector kernels

    def finalize(self, device) -> None:
        self.param = torch.zeros(self.n, device=device, dtype=F32)
        self.grad = torch.zeros(self.n, device=device, dtype=F32)
        self.exp_avg = torch.zeros(self.n, device=device, dtype=F32)
        self.exp_avg_sq = torch.zeros(self.n, device=device, dtype=F32)

    def names(self) -> List[str]:
        return list(self._offsets)

    def offset(self, name: str) -> int:
        return self._offsets[name][0]

    def view(self, name: str, which: str = "param") -> torch.Tensor:
        off, shape = self._offsets[name]
        return getattr(self, which)[off: off + int(math.prod(shape))].view(shape)

    def adamw_step(self, lr: float, weight_decay: float = 0.0, betas=(0.9, 0.999), eps: float = 1e-8, grad_div: float = 1.0) -> None:
        self.step_count += 1
        ops.adamw_(self.param, self.grad, self.exp_avg, self.exp_avg_sq, lr=lr, beta1=betas[0], beta2=betas[1], eps=eps,
                   weight_decay=weight_decay, step=self.step_count, grad_div=grad_div)


# =====================================================================================================
# Whisper encoder (frozen, forward only)
# =====================================================================================================
class WhisperEncoderB200:
    def __init__(self, cfg: EncoderCfg, weights: Optional[Dict[str, torch.Tensor]], device, seed: int = 42, std: float = 0.02):
        self.cfg = cfg
        self.device = _require_cuda(device)
        d = cfg.d
        self.k1 = _round_up(3 * cfg.n_mels, 64)
        if weights is None:
            weights = self._random(cfg, seed, std)
        dev = self.device

        def g(name):
            return weights[name].to(dev, F32)

        w1 = torch.zeros(d, self.k1, device=dev, dtype=F32)
        w1[:, : 3 * cfg.n_mels] = g("conv1.weight").permute(0, 2, 1).reshape(d, 3 * cfg.n_mels)  # [d, (kk, c)]
        self.conv1_w, self.conv1_b = w1.to(BF16), g("conv1.bias").contiguous()
        self.conv2_w = g("conv2.weight").permute(0, 2, 1).reshape(d, 3 * d).to(BF16).contiguous()
        self.conv2_b = g("conv2.bias").contiguous()
        self.pos = g("positional_embedding").contiguous()
        self.layers = []
        for i in range(cfg.layers):
            p = f"blocks.{i}."
            wqkv = torch.cat([g(p + "attn.query.weight"), g(p + "attn.key.weight"), g(p + "attn.value.weight")], 0).to(BF16).contiguous()
            bqkv = torch.cat([g(p + "attn.query.bias"), torch.zeros(d, device=dev), g(p + "attn.value.bias")]).contiguous()
            self.layers.append(dict(
                ln1_w=g(p + "attn_ln.weight").contiguous(), ln1_b=g(p + "attn_ln.bias").contiguous(), wqkv=wqkv, bqkv=bqkv,
                wo=g(p + "attn.out.weight").to(BF16).contiguous(), bo=g(p + "attn.out.bias").contiguous(),
                ln2_w=g(p + "mlp_ln.weight").contiguous(), ln2_b=g(p + "mlp_ln.bias").contiguous(),
                w1=g(p + "mlp.0.weight").to(BF16).contiguous(), b1=g(p + "mlp.0.bias").contiguous(),
                w2=g(p + "mlp.2.weight").to(BF16).contiguous(), b2=g(p + "mlp.2.bias").contiguous()))
        self.lnp_w, self.lnp_b = g("ln_post.weight").contiguous(), g("ln_post.bias").contiguous()

    @staticmethod
    def _random(cfg: EncoderCfg, seed: int, std: float) -> Dict[str, torch.Tensor]:
        """Random-init AudioEncoder weights generated directly on the device (bench / smoke: no checkpoints offline)."""
        g = torch.Generator(device="cuda").manual_seed(seed)
        d = cfg.d

        def n(*s):
            return torch.randn(*s, generator=g, device="cuda") * std

        w = {"conv1.weight": n(d, cfg.n_mels, 3) * 3, "conv1.bias": n(d), "conv2.weight": n(d, d, 3), "conv2.bias": n(d),
             "positional_embedding": sinusoids(cfg.n_ctx, d), "ln_post.weight": 1 + n(d), "ln_post.bias": n(d)}
        for i in range(cfg.layers):
            p = f"blocks.{i}."
            w.update({p + "attn.query.weight": n(d, d), p + "attn.query.bias": n(d), p + "attn.key.weight": n(d, d),
                      p + "attn.value.weight": n(d, d), p + "attn.value.bias": n(d), p + "attn.out.weight": n(d, d), p + "attn.out.bias": n(d),
                      p + "attn_ln.weight": 1 + n(d), p + "attn_ln.bias": n(d), p + "mlp.0.weight": n(4 * d, d), p + "mlp.0.bias": n(4 * d),
                      p + "mlp.2.weight": n(d, 4 * d), p + "mlp.2.bias": n(d), p + "mlp_ln.weight": 1 + n(d), p + "mlp_ln.bias": n(d)})
        return w

    def forward(self, mel: torch.Tensor) -> torch.Tensor:
        """mel f32 [B, T, n_mels] (time-major) -> bf16 [B, ceil(T/2), d]."""
        cfg = self.cfg
        B, T, _ = mel.shape
        d, H = cfg.d, cfg.heads
        dh = d // H
        col1 = ops.conv_im2col(mel.contiguous(), 1, self.k1)
        x1 = ops.gemm(col1, self.conv1_w, bias=self.conv1_b, act=1, static_w=_STATIC_W)                       # GELU(conv1)
        col2 = ops.conv_im2col(x1.view(B, T, d), 2, 3 * d)
        Tp = (T + 1) // 2
        if Tp > self.pos.shape[0]:
            raise ValueError(f"audio too long for positional_embedding: {Tp} > {self.pos.shape[0]}")
        x = ops.gemm(col2, self.conv2_w, bias=self.conv2_b, act=1, static_w=_STATIC_W)                        # GELU(conv2), [B*Tp, d]
        ops.add_pos_(x.view(B, Tp, d), self.pos)
        M = B * Tp
        scale = dh ** -0.5                                                                 # (q dh^-.25)(k dh^-.25)
        for L in self.layers:
            h = ops.layernorm(x, L["ln1_w"], L["ln1_b"])
            qkv = ops.gemm(h, L["wqkv"], bias=L["bqkv"], static_w=_STATIC_W)
            q = qkv[:, :d].view(B, Tp, H, dh)
            k = qkv[:, d:2 * d].view(B, Tp, H, dh)
            v = qkv[:, 2 * d:].view(B, Tp, H, dh)
            a, _ = ops.attn_fwd(q, k, v, causal=False, scale=scale)                        # no mask on padded frames (ref Q9)
            ops.gemm(a.view(M, d), L["wo"], bias=L["bo"], residual=x, out=x, static_w=_STATIC_W)
            h = ops.layernorm(x, L["ln2_w"], L["ln2_b"])
            f = ops.gemm(h, L["w1"], bias=L["b1"], act=1, static_w=_STATIC_W)
            ops.gemm(f, L["w2"], bias=L["b2"], residual=x, out=x, static_w=_STATIC_W)
        return ops.layernorm(x, self.lnp_w, self.lnp_b).view(B, Tp, d)


# =====================================================================================================
# projector (trainable): EncoderProjectorConcat
# =====================================================================================================
class ProjectorB200:
    """EncoderProjectorConcat (projector.py:5-27) and EncoderProjectorCov1d (projector.py:29-49).

    Both are a chain of Linear(+ReLU) layers on the k-frame-concatenated encoder output: Conv1d(d, d, k, stride=k, pad=0) on
    [B, d, T'] is exactly a Linear(k*d -> d) on the non-overlapping windows with the kernel re-laid-out as [co, (kk, c)]."""
    PREFIX = "encoder_projector."

    def __init__(self, enc: EncoderCfg, llm: LlmCfg, proj: ProjCfg, arena: TrainableArena):
        self.cfg, self.k, self.d_enc, self.hidden, self.d_out = proj, proj.k, enc.d, proj.hidden, llm.d
        self.d_in = enc.d * proj.k
        self.arena = arena
        P = self.PREFIX
        if proj.kind == "linear":
            self.chain = [("linear1", self.hidden, self.d_in, True), ("linear2", self.d_out, self.hidden, False)]
        elif proj.kind == "cov1d-linear":
            arena.add(P + "conv1d.weight", (enc.d, enc.d, proj.k))
            arena.add(P + "conv1d.bias", (enc.d,))
            self.chain = [("conv1d", enc.d, self.d_in, True), ("linear1", self.hidden, enc.d, True), ("linear2", self.d_out, self.hidden, False)]
        else:
            raise NotImplementedError(f"projector kind {proj.kind!r} is not on the B200 path (q-former is out of scope)")
        for name, out_f, in_f, _ in self.chain:
            if name != "conv1d":
                arena.add(P + f"{name}.weight", (out_f, in_f))
                arena.add(P + f"{name}.bias", (out_f,))
        self.saved = None

    def param_names(self) -> List[str]:
        return [f"{name}.{sfx}" for name, _, _, _ in self.chain for sfx in ("weight", "bias")]

    def init_weights(self, weights: Optional[Dict[str, torch.Tensor]], seed: int = 45) -> None:
        a = self.arena
        if weights is not None:
            for k in self.param_names():
                a.view(self.PREFIX + k).copy_(weights[k].to(a.param.device, F32))
            return
        g = torch.Generator(device="cuda").manual_seed(seed)
        for name, _, in_f, _ in self.chain:                              # nn.Linear / nn.Conv1d default init
            b = 1.0 / math.sqrt(in_f)
            for suffix in ("weight", "bias"):
                v = a.view(self.PREFIX + f"{name}.{suffix}")
                v.copy_((torch.rand(v.shape, generator=g, device="cuda") * 2 - 1) * b)

    def _weight_bf16(self, name: str, out_f: int, in_f: int) -> torch.Tensor:
        w = self.arena.view(self.PREFIX + f"{name}.weight")
        if name != "conv1d":
            return ops.cast_bf16(w)
        d, k = self.d_enc, self.k                                        # [co, c, kk] -> [co, (kk, c)]
        out = torch.empty((d, k * d), device=w.device, dtype=BF16)
        return ops.pack2d(w, out, batch=d, rows=d, cols=k, src_bs=d * k, src_ld=k, dst_bs=k * d, dst_ld=d, transpose=True)

    def forward(self, enc_out: torch.Tensor, save: bool) -> torch.Tensor:
        """enc_out bf16 [B, T', d] -> bf16 [B, T'//k, D]."""
        B, Tp, d = enc_out.shape
        Ta = Tp // self.k
        if Tp % self.k:
            enc_out = enc_out[:, : Ta * self.k].contiguous()             # projector.py:17-19 / conv stride drops the tail frames
        h = enc_out.reshape(B * Ta, self.k * d)
        acts, wts = [h], []
        for name, out_f, in_f, relu in self.chain:
            w = self._weight_bf16(name, out_f, in_f)
            h = ops.gemm(h, w, bias=self.arena.view(self.PREFIX + f"{name}.bias"), act=2 if relu else 0)
            acts.append(h)
            wts.append(w)
        if save:
            self.saved = (acts, wts)
        return h.view(B, Ta, self.d_out)

    def backward(self, dy: torch.Tensor) -> None:
        """dy bf16 [B, Ta, D]; writes the weight/bias grads into the arena (encoder frozen: no dX)."""
        acts, wts = self.saved
        self.saved = None
        a = self.arena
        M = acts[0].shape[0]
        Mp = _round_up(M, 8)

        def tpose(x):  # [M, C] -> [C, M] with a 16-byte aligned leading dimension
            buf = torch.zeros((x.shape[1], Mp), device=x.device, dtype=BF16) if Mp != M else torch.empty((x.shape[1], M), device=x.device, dtype=BF16)
            return ops.transpose(x, out=buf[:, :M])

        g = dy.reshape(M, self.d_out)
        for li in range(len(self.chain) - 1, -1, -1):
            name, out_f, in_f, relu = self.chain[li]
            if relu:
                g = ops.relu_bwd(g, acts[li + 1], out=g)
            gw = a.view(self.PREFIX + f"{name}.weight", "grad")
            if name == "conv1d":
                tmp = torch.empty((out_f, in_f), device=g.device, dtype=F32)
                ops.gemm(tpose(g), tpose(acts[li]), out=tmp, out_f32=True)                          # [co, (kk, c)]
                ops.transpose_f32_batched(tmp, gw, batch=self.d_enc, rows=self.k, cols=self.d_enc)  # -> [co, c, kk]
            else:
                ops.gemm(tpose(g), tpose(acts[li]), out=gw, out_f32=True)                           # dW = dY^T X
            ops.colsum(g, a.view(self.PREFIX + f"{name}.bias", "grad"))
            if li > 0:
                g = ops.gemm(g, ops.transpose(wts[li]))                                             # dX = dY W


# =====================================================================================================
# Llama decoder with frozen base weights + LoRA adapters
# =====================================================================================================
GROUPS = {"qkv": ("q_proj", "k_proj", "v_proj"), "o": ("o_proj",), "gu": ("gate_proj", "up_proj"), "down": ("down_proj",)}


class LlamaLoRAB200:
    LORA_PREFIX = "llm.base_model.model."

    BASE_PREFIX = "llm."      # reference state-dict prefix of an un-wrapped (non-peft) HF causal LM held as `slam_model.llm`

    def __init__(self, cfg: LlmCfg, lora: Optional[LoraCfg], arena: TrainableArena, device, weights: Optional[Dict[str, torch.Tensor]] = None,
                 seed: int = 43, std: float = 0.02, train_base: bool = False):
        """train_base (train_config.freeze_llm=false, full fine-tune as in examples/s2s): every decoder parameter is an fp32 master tensor in
        the arena under its HF name; bf16 GEMM operands (W and W^T) are re-derived from the masters at the start of every pass and the
        backward adds a weight-gradient GEMM (dW = dY^T X on the tcgen05 kernel) per linear, bias / RMSNorm-weight / embedding gradients."""
        self.cfg, self.lora, self.arena = cfg, lora, arena
        self.device = _require_cuda(device)
        self.train_base = bool(train_base)
        self._static_w = _STATIC_W and not self.train_base          # the base weights are frozen (LoRA / projector-only training)
        if self.train_base and lora is not None:
            raise NotImplementedError("full fine-tune together with LoRA adapters is not implemented (the reference recipes use one or the other)")
        dev = self.device
        L = cfg.layers
        gen = torch.Generator(device="cuda").manual_seed(seed)
        self._init_src: Dict[str, torch.Tensor] = {}      # train_base: initial values, copied into the arena by init_base() after finalize()

        def base(name: str, shape) -> torch.Tensor:
            if weights is not None:
                t = weights[name].to(dev, F32 if self.train_base else BF16).contiguous()
            else:
                t = torch.randn(*shape, generator=gen, device=dev, dtype=F32) * std
            if self.train_base:
                self._init_src[name] = t
            return t.to(BF16)

        def norm_w(name: str) -> torch.Tensor:
            if weights is not None:
                t = weights[name].to(dev, F32 if self.train_base else BF16).contiguous()
            else:
                t = 1.0 + torch.randn(cfg.d, generator=gen, device=dev) * std
            if self.train_base:
                self._init_src[name] = t
            return t.to(BF16)

        def bias_f32(name: str, n: int) -> torch.Tensor:
            if weights is not None:
                t = weights[name].to(dev, F32).contiguous()
            else:
                t = torch.randn(n, generator=gen, device=dev, dtype=F32) * std
            if self.train_base:
                self._init_src[name] = t
            return t

        self.embed = base("model.embed_tokens.weight", (cfg.vocab, cfg.d))
        # Fused SwiGLU (GEMM epilogues act 3 / 4): without LoRA on gate/up the concatenated gate/up weight is stored "blocked-64"
        # (64 gate rows, then the 64 up rows of the same features, ...) so that a GEMM tile holds both halves of a feature pair.
        # With adapters on gate/up the HF [gate | up] order is kept (their gradient slices address whole projections).
        lora_targets = set(lora.targets) if lora is not None else set()
        self.fuse_swiglu = (not ({"gate_proj", "up_proj"} & lora_targets) and cfg.ffn % 64 == 0 and not self.train_base
                            and os.environ.get("SLAM_FUSE_SWIGLU", "1") != "0")      # trainable gate/up keep the HF [gate | up] order
        self.fuse_swiglu_bwd = self.fuse_swiglu and os.environ.get("SLAM_FUSE_SWIGLU_BWD", "1") != "0"
        self.layers = []
        for i in range(L):
            p = f"model.layers.{i}."
            q, k, v = (base(p + f"self_attn.{n}.weight", linear_shape(cfg, n)) for n in ("q_proj", "k_proj", "v_proj"))
            wqkv = torch.cat([q, k, v], 0).contiguous()
            del q, k, v
            wo = base(p + "self_attn.o_proj.weight", linear_shape(cfg, "o_proj"))
            g_, u_ = (base(p + f"mlp.{n}.weight", linear_shape(cfg, n)) for n in ("gate_proj", "up_proj"))
            if self.fuse_swiglu:
                wgu = torch.stack([g_.view(cfg.ffn // 64, 64, cfg.d), u_.view(cfg.ffn // 64, 64, cfg.d)], 1).reshape(2 * cfg.ffn, cfg.d).contiguous()
            else:
                wgu = torch.cat([g_, u_], 0).contiguous()
            del g_, u_
            wd = base(p + "mlp.down_proj.weight", linear_shape(cfg, "down_proj"))
            bqkv = None
            if cfg.qkv_bias:                                                     # Qwen2: q/k/v biases, added in the GEMM epilogue (fp32)
                bqkv = torch.cat([bias_f32(p + f"self_attn.{n}.bias", linear_shape(cfg, n)[0]) for n in ("q_proj", "k_proj", "v_proj")]).contiguous()
            self.layers.append(dict(wqkv=wqkv, wqkvT=ops.transpose(wqkv), wo=wo, woT=ops.transpose(wo), wgu=wgu, wguT=ops.transpose(wgu),
                                    wd=wd, wdT=ops.transpose(wd), ln1=norm_w(p + "input_layernorm.weight"),
                                    ln2=norm_w(p + "post_attention_layernorm.weight"), bqkv=bqkv))
        self.norm = norm_w("model.norm.weight")
        if cfg.tie_embeddings and (weights is None or "lm_head.weight" not in weights or self.train_base):
            self.lm_head = self.embed                                            # tied: ONE table (and one gradient) for both uses
        else:
            self.lm_head = base("lm_head.weight", (cfg.vocab, cfg.d))
        self.lm_headT = ops.transpose(self.lm_head)
        if self.train_base:
            self._register_base(arena)
        self._rope_cache: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}

        # ---- LoRA bookkeeping: per fused-GEMM group, rank-padded packed operands for all layers
        self.groups: Dict[str, dict] = {}
        if lora is not None:
            for gname, members in GROUPS.items():
                tg = [m for m in members if m in lora.targets]
                if not tg:
                    continue
                in_f = linear_shape(cfg, members[0])[1]
                col, cols = 0, {}
                for m in members:
                    cols[m] = col
                    col += linear_shape(cfg, m)[0]
                out_total = col
                rpad = _round_up(len(tg) * lora.r, 64)
                info = dict(targets=tg, in_f=in_f, out_total=out_total, rpad=rpad, cols=cols,
                            roff={m: j * lora.r for j, m in enumerate(tg)},
                            a_cat=torch.zeros(L, rpad, in_f, device=dev, dtype=BF16), a_catT=torch.zeros(L, in_f, rpad, device=dev, dtype=BF16),
                            b_cat=torch.zeros(L, out_total, rpad, device=dev, dtype=BF16), b_catT=torch.zeros(L, rpad, out_total, device=dev, dtype=BF16))
                self.groups[gname] = info
                for m in tg:
                    out_f = linear_shape(cfg, m)[0]
                    arena.add(self._pname(m, "A"), (L, lora.r, in_f))      # lora_A.weight [r, in] stacked over layers
                    arena.add(self._pname(m, "Bt"), (L, lora.r, out_f))    # lora_B.weight^T [r, out] stacked over layers
        self.saved: Optional[dict] = None
        # LoRA-branch dropout (peft lora_dropout): active only for training forwards when the owner enables it
        self.dropout_p = float(lora.dropout) if lora is not None else 0.0
        self.dropout_active = False
        self.dropout_seed = seed
        self.dropout_step = 0

    # ---------------------------------------------------------------------------------------- full fine-tune: fp32 masters in the arena
    def _base_names(self) -> List[Tuple[str, Tuple[int, ...]]]:
        """HF parameter names in ARENA ORDER: q,k,v (and their biases) and gate,up are adjacent so that the fused GEMM operands are plain
        views of consecutive arena tensors (every size is a multiple of 8 elements, so the arena inserts no padding between them)."""
        cfg = self.cfg
        out = [("model.embed_tokens.weight", (cfg.vocab, cfg.d))]
        for i in range(cfg.layers):
            p = f"model.layers.{i}."
            out += [(p + f"self_attn.{n}.weight", linear_shape(cfg, n)) for n in ("q_proj", "k_proj", "v_proj")]
            if cfg.qkv_bias:
                out += [(p + f"self_attn.{n}.bias", (linear_shape(cfg, n)[0],)) for n in ("q_proj", "k_proj", "v_proj")]
            out += [(p + "self_attn.o_proj.weight", linear_shape(cfg, "o_proj"))]
            out += [(p + f"mlp.{n}.weight", linear_shape(cfg, n)) for n in ("gate_proj", "up_proj")]
            out += [(p + "mlp.down_proj.weight", linear_shape(cfg, "down_proj")), (p + "input_layernorm.weight", (cfg.d,)),
                    (p + "post_attention_layernorm.weight", (cfg.d,))]
        out.append(("model.norm.weight", (cfg.d,)))
        if self.lm_head is not self.embed:
            out.append(("lm_head.weight", (cfg.vocab, cfg.d)))
        return out

    def _register_base(self, arena: TrainableArena) -> None:
        cfg = self.cfg
        if cfg.d % 8 or cfg.dkv % 8 or cfg.ffn % 8:
            raise ValueError("full fine-tune needs hidden / kv / ffn sizes that are multiples of 8 (fused arena views)")
        for name, shape in self._base_names():
            arena.add(self.BASE_PREFIX + name, shape)

    def _fused(self, first: str, rows: int, cols: int, which: str) -> torch.Tensor:
        """[rows, cols] (or [rows]) view over consecutive arena tensors starting at `first`."""
        off = self.arena.offset(self.BASE_PREFIX + first)
        buf = getattr(self.arena, which)
        return buf[off: off + rows * max(cols, 1)].view(rows, cols) if cols else buf[off: off + rows]

    def init_base(self) -> None:
        """Copy the initial decoder weights (checkpoint or random) into the fp32 arena masters (after arena.finalize())."""
        if not self.train_base:
            return
        for name, _ in self._base_names():
            self.arena.view(self.BASE_PREFIX + name).copy_(self._init_src[name].to(F32).view(self.arena.view(self.BASE_PREFIX + name).shape))
        self._init_src = {}
        self.refresh_base()

    def base_state(self, which: str = "param") -> Dict[str, torch.Tensor]:
        return {self.BASE_PREFIX + n: self.arena.view(self.BASE_PREFIX + n, which) for n, _ in self._base_names()} if self.train_base else {}

    def refresh_base(self) -> None:
        """fp32 masters -> the bf16 GEMM operands W and W^T, bf16 norm weights, fp32 fused biases (once per pass: AdamW just changed them)."""
        cfg = self.cfg
        D, Dq, Dkv, Fd = cfg.d, cfg.dq, cfg.dkv, cfg.ffn

        def both(src: torch.Tensor, w: torch.Tensor, wt: torch.Tensor) -> None:
            r, c = src.shape
            ops.cast_bf16(src, out=w)
            ops.pack2d(src, wt, batch=1, rows=r, cols=c, src_bs=r * c, src_ld=c, dst_bs=r * c, dst_ld=r, transpose=True)

        if self.lm_head is self.embed:
            both(self.arena.view(self.BASE_PREFIX + "model.embed_tokens.weight"), self.embed, self.lm_headT)
        else:
            ops.cast_bf16(self.arena.view(self.BASE_PREFIX + "model.embed_tokens.weight"), out=self.embed)
        for i, Lw in enumerate(self.layers):
            p = f"model.layers.{i}."
            both(self._fused(p + "self_attn.q_proj.weight", Dq + 2 * Dkv, D, "param"), Lw["wqkv"], Lw["wqkvT"])
            both(self.arena.view(self.BASE_PREFIX + p + "self_attn.o_proj.weight"), Lw["wo"], Lw["woT"])
            both(self._fused(p + "mlp.gate_proj.weight", 2 * Fd, D, "param"), Lw["wgu"], Lw["wguT"])
            both(self.arena.view(self.BASE_PREFIX + p + "mlp.down_proj.weight"), Lw["wd"], Lw["wdT"])
            ops.cast_bf16(self.arena.view(self.BASE_PREFIX + p + "input_layernorm.weight"), out=Lw["ln1"])
            ops.cast_bf16(self.arena.view(self.BASE_PREFIX + p + "post_attention_layernorm.weight"), out=Lw["ln2"])
            if cfg.qkv_bias:
                Lw["bqkv"] = self._fused(p + "self_attn.q_proj.bias", Dq + 2 * Dkv, 0, "param")   # the fp32 master itself (epilogue reads fp32)
        ops.cast_bf16(self.arena.view(self.BASE_PREFIX + "model.norm.weight"), out=self.norm)
        if self.lm_head is not self.embed:
            both(self.arena.view(self.BASE_PREFIX + "lm_head.weight"), self.lm_head, self.lm_headT)

    def _wgrad(self, dy: torch.Tensor, x: torch.Tensor, first: str, rows: int, cols: int) -> None:
        """dW [rows = out, cols = in] (fp32, into the gradient arena) = dY^T X on the tcgen05 GEMM (K = tokens, padded to a multiple of 8)."""
        M = dy.shape[0]
        Mp = _round_up(M, 8)

        def tp(t):
            buf = torch.zeros((t.shape[1], Mp), device=t.device, dtype=BF16) if Mp != M else torch.empty((t.shape[1], M), device=t.device, dtype=BF16)
            ops.transpose(t, out=buf[:, :M])
            return buf
        ops.gemm(tp(dy), tp(x), out=self._fused(first, rows, cols, "grad"), out_f32=True)

    # names of the stacked arena tensors (the slam_llm mirror exposes them under the peft key names)
    @staticmethod
    def _pname(target: str, which: str) -> str:
        return f"lora.{target}.{which}"

    def init_lora(self, weights: Optional[Dict[str, torch.Tensor]], seed: int = 44, b_std: float = 0.0) -> None:
        """weights: peft-named dict ('model.layers.i.self_attn.q_proj.lora_A.default.weight').  Without weights:
        A ~ kaiming_uniform(a=sqrt(5)), B = 0 (peft reset_lora_parameters) or N(0, b_std) when b_std > 0."""
        if self.lora is None:
            return
        cfg, lora, a = self.cfg, self.lora, self.arena
        g = torch.Generator(device="cuda").manual_seed(seed)
        for m in lora.targets:
            out_f, in_f = linear_shape(cfg, m)
            mod = "self_attn" if m in ATTN_LINEARS else "mlp"
            A, Bt = a.view(self._pname(m, "A")), a.view(self._pname(m, "Bt"))
            for i in range(cfg.layers):
                if weights is not None:
                    p = f"model.layers.{i}.{mod}.{m}."
                    A[i].copy_(weights[p + "lora_A.default.weight"].to(A.device, F32))
                    Bt[i].copy_(weights[p + "lora_B.default.weight"].to(A.device, F32).t())
                else:
                    A[i].copy_((torch.rand(lora.r, in_f, generator=g, device="cuda") * 2 - 1) / math.sqrt(in_f))
                    if b_std > 0:
                        Bt[i].copy_(torch.randn(out_f, lora.r, generator=g, device="cuda").t() * b_std)
                    else:
                        Bt[i].zero_()

    def lora_state(self, which: str = "param") -> Dict[str, torch.Tensor]:
        """peft-0.6-named views (lora_B as a transposed view of the stacked [r,out] storage)."""
        out = {}
        if self.lora is None:
            return out
        for m in self.lora.targets:
            mod = "self_attn" if m in ATTN_LINEARS else "mlp"
            A, Bt = self.arena.view(self._pname(m, "A"), which), self.arena.view(self._pname(m, "Bt"), which)
            for i in range(self.cfg.layers):
                p = f"{self.LORA_PREFIX}model.layers.{i}.{mod}.{m}."
                out[p + "lora_A.default.weight"] = A[i]
                out[p + "lora_B.default.weight"] = Bt[i].t()
        return out

    def pack_lora(self) -> None:
        """fp32 adapters -> rank-padded bf16 GEMM operands for all layers (8 launches per target, once per step)."""
        if self.lora is None:
            return
        L, r, s = self.cfg.layers, self.lora.r, self.lora.scaling
        for info in self.groups.values():
            in_f, out_total, rpad = info["in_f"], info["out_total"], info["rpad"]
            for m in info["targets"]:
                out_f = linear_shape(self.cfg, m)[0]
                off, col = info["roff"][m], info["cols"][m]
                A, Bt = self.arena.view(self._pname(m, "A")), self.arena.view(self._pname(m, "Bt"))
                ops.pack2d(A, info["a_cat"], batch=L, rows=r, cols=in_f, src_bs=r * in_f, src_ld=in_f, dst_bs=rpad * in_f, dst_ld=in_f,
                           dst_off=off * in_f)
                ops.pack2d(A, info["a_catT"], batch=L, rows=r, cols=in_f, src_bs=r * in_f, src_ld=in_f, dst_bs=in_f * rpad, dst_ld=rpad,
                           dst_off=off, transpose=True)
                ops.pack2d(Bt, info["b_catT"], batch=L, rows=r, cols=out_f, src_bs=r * out_f, src_ld=out_f, dst_bs=rpad * out_total,
                           dst_ld=out_total, dst_off=off * out_total + col, scale=s)
                ops.pack2d(Bt, info["b_cat"], batch=L, rows=r, cols=out_f, src_bs=r * out_f, src_ld=out_f, dst_bs=out_total * rpad, dst_ld=rpad,
                           dst_off=col * rpad + off, scale=s, transpose=True)

    def rope_tables(self, S: int) -> Tuple[torch.Tensor, torch.Tensor]:
        if S not in self._rope_cache:
            dh = self.cfg.dh
            inv = 1.0 / (self.cfg.rope_theta ** (torch.arange(0, dh, 2, dtype=F32, device=self.device) / dh))
            fr = torch.outer(torch.arange(S, dtype=F32, device=self.device), inv)
            self._rope_cache[S] = (fr.cos().contiguous(), fr.sin().contiguous())
        return self._rope_cache[S]

    # ---------------------------------------------------------------------------------------- linear (+LoRA) helpers
    def _drop_seed(self, li: int, gname: str) -> int:
        gi = list(GROUPS).index(gname)
        return (self.dropout_seed * 1000003 + self.dropout_step * 4099 + li * 8 + gi) & 0x7FFFFFFFFFFFFFFF

    def _lin_fwd(self, x, w, gname: str, li: int, residual=None, out=None, bias=None):
        """y = x W^T (+ bias) (+ residual) + (dropout(x) A_cat^T)(s B_cat)^T  ->  (y, saved) with saved = (x_lora, T, seed) for the backward."""
        info = self.groups.get(gname)
        swap = bias is None and _swap_ab(x.shape[0], w.shape[0])
        if info is None:
            if swap:
                return ops.gemm(w, x, residual=residual, out=out, transpose_out=True, static_w=self._static_w,
                                block_n=_swap_tile(x.shape[0], w.shape[0], w.shape[1])), None
            return ops.gemm(x, w, residual=residual, out=out, bias=bias, static_w=self._static_w), None
        p, seed = self.dropout_p if self.dropout_active else 0.0, 0
        x_lora = x
        if p > 0.0:
            seed = self._drop_seed(li, gname)
            x_lora = ops.dropout(x, p, seed)                                             # lora_A(dropout(x)): LoRA branch only
        t = _gemm_few_tiles(x_lora, info["a_cat"][li])                                   # T = x A_cat^T  [M, rpad]
        if swap:                                                                         # y^T tiles = W x^T + (s B_cat) T^T: same fused tile, operands swapped
            y = ops.gemm(w, x, a2=info["b_cat"][li], b2=t, residual=residual, out=out, transpose_out=True, static_w=self._static_w,
                         block_n=_swap_tile(x.shape[0], w.shape[0], w.shape[1]))
        else:
            y = ops.gemm(x, w, a2=t, b2=info["b_cat"][li], residual=residual, out=out, bias=bias, static_w=self._static_w)   # fused base + LoRA tile
        return y, (x_lora, t, p, seed)

    def _lin_bwd(self, dy, wT, gname: str, li: int, saved):
        """dX = dY W + mask o ((dY sB) A)  and LoRA grads (dA = U^T x_lora, dB^T = s T^T dY) into the arena."""
        info = self.groups.get(gname)
        swap = _swap_ab(dy.shape[0], wT.shape[0])
        if info is None:
            return (ops.gemm(wT, dy, transpose_out=True, static_w=self._static_w, block_n=_swap_tile(dy.shape[0], wT.shape[0], wT.shape[1])) if swap
                    else ops.gemm(dy, wT, static_w=self._static_w))
        x_lora, t, p, seed = saved
        u = _gemm_few_tiles(dy, info["b_catT"][li])                                      # U = dY (s B)  [M, rpad]
        if p > 0.0:
            dx = ops.dropout_bwd_add(ops.gemm(dy, wT), ops.gemm(u, info["a_catT"][li]), p, seed)
        elif swap:
            dx = ops.gemm(wT, dy, a2=info["a_catT"][li], b2=u, transpose_out=True, static_w=self._static_w,
                          block_n=_swap_tile(dy.shape[0], wT.shape[0], wT.shape[1]))
        else:
            dx = ops.gemm(dy, wT, a2=u, b2=info["a_catT"][li], static_w=self._static_w)                           # fused: one accumulator tile
        r, s = self.lora.r, self.lora.scaling
        for m in info["targets"]:
            off, col = info["roff"][m], info["cols"][m]
            out_f = linear_shape(self.cfg, m)[0]
            gA = self.arena.view(self._pname(m, "A"), "grad")[li]
            gBt = self.arena.view(self._pname(m, "Bt"), "grad")[li]
            ops.wgrad_thin(u[:, off: off + r], x_lora, gA, accumulate=True)          # (the flat gradient buffer was zeroed once: backward())
            ops.wgrad_thin(t[:, off: off + r], dy[:, col: col + out_f], gBt, scale=s, accumulate=True)
        return dx

    # ---------------------------------------------------------------------------------------- forward
    def forward(self, x: torch.Tensor, key_mask: torch.Tensor, save: bool, kv_out: Optional[list] = None) -> torch.Tensor:
        """x bf16 [B,S,D] (inputs_embeds), key_mask u8 [B,S] -> final normed hidden bf16 [B*S, D].
        kv_out (decode prefill): receives one (K [B,S,Hkv,dh], V [B,S,Hkv,dh]) pair of post-RoPE keys / values per layer."""
        cfg = self.cfg
        B, S, D = x.shape
        M = B * S
        H, Hkv, dh, Dq, Dkv = cfg.heads, cfg.kv_heads, cfg.dh, cfg.dq, cfg.dkv
        cos, sin = self.rope_tables(S)
        scale = 1.0 / math.sqrt(dh)
        _THIN_POOL.begin(x.device)
        x = x.reshape(M, D)
        saved_layers = []
        for li, Lw in enumerate(self.layers):
            xn1, rstd1 = ops.rmsnorm_fwd(x, Lw["ln1"], cfg.eps, need_rstd=save)
            qkv, sv_qkv = self._lin_fwd(xn1, Lw["wqkv"], "qkv", li, bias=Lw["bqkv"])
            ops.rope_(qkv[:, : Dq + Dkv], H + Hkv, dh, S, cos, sin)                  # q and k heads are adjacent in the fused buffer: one launch
            q = qkv[:, :Dq].view(B, S, H, dh)
            k = qkv[:, Dq: Dq + Dkv].view(B, S, Hkv, dh)
            v = qkv[:, Dq + Dkv:].view(B, S, Hkv, dh)
            attn, lse = ops.attn_fwd(q, k, v, causal=True, scale=scale, key_mask=key_mask, need_lse=save)
            if kv_out is not None:
                kv_out.append((k.contiguous(), v.contiguous()))
            attn2 = attn.view(M, Dq)
            x2, sv_o = self._lin_fwd(attn2, Lw["wo"], "o", li, residual=x)
            xn2, rstd2 = ops.rmsnorm_fwd(x2, Lw["ln2"], cfg.eps, need_rstd=save)
            if self.fuse_swiglu:
                hmid = torch.empty((M, cfg.ffn), device=x.device, dtype=BF16)
                gu, sv_gu = ops.gemm(xn2, Lw["wgu"], act=3, aux=hmid, static_w=self._static_w), None        # gu (blocked-64) and silu(g) * u from one epilogue
            else:
                gu, sv_gu = self._lin_fwd(xn2, Lw["wgu"], "gu", li)
                hmid = ops.swiglu_fwd(gu)
            x3, sv_d = self._lin_fwd(hmid, Lw["wd"], "down", li, residual=x2)
            if save:
                saved_layers.append(dict(x=x, rstd1=rstd1, qkv=qkv, attn=attn, lse=lse, x2=x2, rstd2=rstd2, gu=gu,
                                         sv_qkv=sv_qkv, sv_o=sv_o, sv_gu=sv_gu, sv_d=sv_d))
                if self.train_base:                                                   # inputs of the linears: the wgrad GEMMs need them
                    saved_layers[-1].update(xn1=xn1, xn2=xn2, hmid=hmid)
            x = x3
        xf, rstd_f = ops.rmsnorm_fwd(x, self.norm, cfg.eps, need_rstd=save)
        if save:
            self.saved = dict(layers=saved_layers, x_last=x, rstd_f=rstd_f, B=B, S=S, key_mask=key_mask)
        return xf

    # ---------------------------------------------------------------------------------------- decode (KV cache)
    def decode_step(self, x: torch.Tensor, cache: list, key_mask: torch.Tensor, pos: int) -> torch.Tensor:
        """One new token per sequence: x bf16 [n, D] (its embedding), cache = [(K, V)] per layer with K/V [n, t, Hkv, dh] (post-RoPE, grown in
        place of the list), key_mask u8 [n, t + 1] over cached + new positions, pos = the new token's position (RoPE index).
        Returns the final normed hidden bf16 [n, D].  Same kernels as the training forward: fused QKV(+LoRA) GEMM, RoPE, flash attention
        (q_len 1 against the cache, non-causal + key mask), O / SwiGLU GEMMs."""
        cfg = self.cfg
        n = x.shape[0]
        H, Hkv, dh, Dq, Dkv = cfg.heads, cfg.kv_heads, cfg.dh, cfg.dq, cfg.dkv
        cos, sin = self.rope_tables(pos + 1)
        cos_t, sin_t = cos[pos: pos + 1].contiguous(), sin[pos: pos + 1].contiguous()
        scale = 1.0 / math.sqrt(dh)
        _THIN_POOL.begin(x.device)
        for li, Lw in enumerate(self.layers):
            xn1, _ = ops.rmsnorm_fwd(x, Lw["ln1"], cfg.eps, need_rstd=False)
            qkv, _ = self._lin_fwd(xn1, Lw["wqkv"], "qkv", li, bias=Lw["bqkv"])
            ops.rope_(qkv[:, : Dq + Dkv], H + Hkv, dh, 1, cos_t, sin_t)                # every row sits at position `pos`
            K, V = cache[li]
            K = torch.cat([K, qkv[:, Dq: Dq + Dkv].reshape(n, 1, Hkv, dh)], dim=1)
            V = torch.cat([V, qkv[:, Dq + Dkv:].reshape(n, 1, Hkv, dh)], dim=1)
            cache[li] = (K, V)
            attn, _ = ops.attn_fwd(qkv[:, :Dq].reshape(n, 1, H, dh).contiguous(), K, V, causal=False, scale=scale, key_mask=key_mask)
            x2, _ = self._lin_fwd(attn.view(n, Dq), Lw["wo"], "o", li, residual=x)
            xn2, _ = ops.rmsnorm_fwd(x2, Lw["ln2"], cfg.eps, need_rstd=False)
            if self.fuse_swiglu:
                hmid = torch.empty((n, cfg.ffn), device=x.device, dtype=BF16)
                ops.gemm(xn2, Lw["wgu"], act=3, aux=hmid, static_w=self._static_w)
            else:
                gu, _ = self._lin_fwd(xn2, Lw["wgu"], "gu", li)
                hmid = ops.swiglu_fwd(gu)
            x, _ = self._lin_fwd(hmid, Lw["wd"], "down", li, residual=x2)
        xf, _ = ops.rmsnorm_fwd(x, self.norm, cfg.eps, need_rstd=False)
        return xf

    # ---------------------------------------------------------------------------------------- backward
    def backward(self, dxf: torch.Tensor) -> torch.Tensor:
        """dxf bf16 [B*S, D]: grad wrt the final normed hidden -> grad wrt inputs_embeds bf16 [B,S,D]."""
        cfg, sv = self.cfg, self.saved
        self.saved = None
        B, S = sv["B"], sv["S"]
        M = B * S
        H, Hkv, dh, Dq, Dkv = cfg.heads, cfg.kv_heads, cfg.dh, cfg.dq, cfg.dkv
        cos, sin = self.rope_tables(S)
        scale = 1.0 / math.sqrt(dh)
        _THIN_POOL.begin(dxf.device)
        dx = ops.rmsnorm_bwd(dxf, sv["x_last"], self.norm, sv["rstd_f"])
        tb = self.train_base
        Fd, Dm = cfg.ffn, cfg.d
        if tb:
            ops.rmsnorm_wgrad(dxf, sv["x_last"], sv["rstd_f"], self.arena.view(self.BASE_PREFIX + "model.norm.weight", "grad"))
        for li in range(cfg.layers - 1, -1, -1):
            Lw, kp = self.layers[li], sv["layers"][li]
            pn = f"model.layers.{li}."
            # ---- MLP block: x3 = x2 + down(silu(g) * u)
            if self.fuse_swiglu_bwd and "down" not in self.groups:
                dgu = (ops.gemm(Lw["wdT"], dx, act=4, aux=kp["gu"], transpose_out=True, static_w=self._static_w) if _SWAP_DDOWN and _swap_ab(dx.shape[0], Lw["wdT"].shape[0])
                       else ops.gemm(dx, Lw["wdT"], act=4, aux=kp["gu"], static_w=self._static_w))                  # dh = dY W_down stays in TMEM: the epilogue emits d(gu)
            else:
                dhmid = self._lin_bwd(dx, Lw["wdT"], "down", li, kp["sv_d"])
                if tb:
                    self._wgrad(dx, kp["hmid"], pn + "mlp.down_proj.weight", Dm, Fd)
                dgu = ops.swiglu_bwd(kp["gu"], dhmid, block=64 if self.fuse_swiglu else 0)
            dxn2 = self._lin_bwd(dgu, Lw["wguT"], "gu", li, kp["sv_gu"])
            if tb:
                self._wgrad(dgu, kp["xn2"], pn + "mlp.gate_proj.weight", 2 * Fd, Dm)      # gate and up are adjacent in the arena: one GEMM
                ops.rmsnorm_wgrad(dxn2, kp["x2"], kp["rstd2"], self.arena.view(self.BASE_PREFIX + pn + "post_attention_layernorm.weight", "grad"))
            dx2 = ops.rmsnorm_bwd(dxn2, kp["x2"], Lw["ln2"], kp["rstd2"], dres=dx)
            # ---- attention block: x2 = x + o(attn(rope(qkv(norm(x)))))
            dattn = self._lin_bwd(dx2, Lw["woT"], "o", li, kp["sv_o"])
            if tb:
                self._wgrad(dx2, kp["attn"].view(M, Dq), pn + "self_attn.o_proj.weight", Dm, Dq)
            qkv = kp["qkv"]
            q = qkv[:, :Dq].view(B, S, H, dh)
            k = qkv[:, Dq: Dq + Dkv].view(B, S, Hkv, dh)
            v = qkv[:, Dq + Dkv:].view(B, S, Hkv, dh)
            dqkv = torch.empty_like(qkv)
            ops.attn_bwd(q, k, v, kp["attn"], kp["lse"], dattn.view(B, S, H, dh), causal=True, scale=scale, key_mask=sv["key_mask"],
                         dq=dqkv[:, :Dq].view(B, S, H, dh), dk=dqkv[:, Dq: Dq + Dkv].view(B, S, Hkv, dh), dv=dqkv[:, Dq + Dkv:].view(B, S, Hkv, dh),
                         rope=(cos, sin))                                               # inverse RoPE of dQ / dK fused into the finishing kernel
            dxn1 = self._lin_bwd(dqkv, Lw["wqkvT"], "qkv", li, kp["sv_qkv"])
            if tb:
                self._wgrad(dqkv, kp["xn1"], pn + "self_attn.q_proj.weight", Dq + 2 * Dkv, Dm)   # q, k, v adjacent: one GEMM
                if cfg.qkv_bias:
                    ops.colsum(dqkv, self._fused(pn + "self_attn.q_proj.bias", Dq + 2 * Dkv, 0, "grad"))
                ops.rmsnorm_wgrad(dxn1, kp["x"], kp["rstd1"], self.arena.view(self.BASE_PREFIX + pn + "input_layernorm.weight", "grad"))
            dx = ops.rmsnorm_bwd(dxn1, kp["x"], Lw["ln1"], kp["rstd1"], dres=dx2)
            sv["layers"][li] = None
        return dx.view(B, S, cfg.d)


# =====================================================================================================
# the step
# =====================================================================================================
class SlamStepB200:
    """log-mel -> encoder -> projector -> merge -> decoder -> CE (+acc); backward; AdamW."""

    def __init__(self, enc_cfg: EncoderCfg, llm_cfg: LlmCfg, lora_cfg: Optional[LoraCfg], proj_cfg: ProjCfg, device="cuda:0",
                 enc_weights=None, llm_weights=None, lora_weights=None, proj_weights=None, seed: int = 42, lora_b_std: float = 0.0,
                 train_llm: bool = False):
        device = _require_cuda(device)
        torch.cuda.set_device(device)
        arena = TrainableArena()
        encoder = WhisperEncoderB200(enc_cfg, enc_weights, device, seed=seed)
        projector = ProjectorB200(enc_cfg, llm_cfg, proj_cfg, arena)
        llm = LlamaLoRAB200(llm_cfg, lora_cfg, arena, device, llm_weights, seed=seed + 1, train_base=train_llm)
        arena.finalize(device)
        projector.init_weights(proj_weights, seed=seed + 3)
        llm.init_lora(lora_weights, seed=seed + 2, b_std=lora_b_std)
        llm.init_base()
        self._assemble(encoder, projector, llm, arena, device)

    @classmethod
    def from_parts(cls, encoder: Optional[WhisperEncoderB200], projector: ProjectorB200, llm: LlamaLoRAB200, arena: TrainableArena, device,
                   enc_cfg: Optional[EncoderCfg] = None) -> "SlamStepB200":
        """Assemble a step from already-built components sharing one (finalized) arena (used by the slam_llm mirror).
        encoder=None: the modality encoder is a foreign (frozen, torch) module run by the caller, who enters at forward_rest() with its output;
        enc_cfg then only carries the feature width `d`."""
        self = cls.__new__(cls)
        self._assemble(encoder, projector, llm, arena, _require_cuda(device), enc_cfg)
        return self

    def _assemble(self, encoder, projector, llm, arena, device, enc_cfg=None) -> None:
        self.device = device
        self.enc_cfg = encoder.cfg if encoder is not None else enc_cfg
        self.llm_cfg, self.lora_cfg, self.proj_cfg = llm.cfg, llm.lora, projector.cfg
        self.arena, self.encoder, self.projector, self.llm = arena, encoder, projector, llm
        self.filters_t = mel_filterbank(self.enc_cfg.n_mels).t().contiguous().to(device) if encoder is not None else None
        self._ctx = None
        self.micro_steps = 0   # backward() calls since the last optimizer step (gradient accumulation)
        self.lora_dropout_enabled = True   # module.train()/eval() of the host mirror toggles this (reference quirk Q6)
        # Deferred update (data-parallel overlap): optimizer_step() only RECORDS the update; it is applied by flush_update(), which the next
        # forward() calls AFTER launching the frozen front end (log-mel + Whisper encoder, independent of the trainables).  The gradient
        # all-reduce issued with async_op=True therefore overlaps ~a quarter of the next step, and a rank that arrives early at the
        # collective keeps computing instead of idling (per-step jitter between power-capped GPUs is absorbed).
        self.defer_update = False
        self._pending_update = None        # (lr, weight_decay, grad_div)
        self._pending_work = None          # torch.distributed Work of the in-flight gradient all-reduce
        self._carry = None                 # gradients of earlier micro-steps (accumulation) while a backward is in progress

    # ------------------------------------------------------------------ state dict in the reference's key names
    def trainable_state(self, which: str = "param") -> Dict[str, torch.Tensor]:
        self.flush_update()
        out = {n: self.arena.view(n, which) for n in self.arena.names() if n.startswith(ProjectorB200.PREFIX)}
        out.update(self.llm.lora_state(which))
        out.update(self.llm.base_state(which))
        return out

    def load_trainable_state(self, sd: Dict[str, torch.Tensor]) -> List[str]:
        mine = self.trainable_state()
        loaded = []
        for k, v in sd.items():
            if k in mine:
                mine[k].copy_(v.to(self.device, F32))
                loaded.append(k)
        return loaded

    # ------------------------------------------------------------------ front end
    def log_mel(self, pcm: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        """pcm f32 [B, n_samples] on device (+ optional i32 [B] real lengths) -> [B, n_samples//160, n_mels]."""
        if lengths is not None:
            lengths = lengths.to(self.device, torch.int32).contiguous()
        return ops.logmel(pcm.contiguous(), self.filters_t, lengths=lengths)

    # ------------------------------------------------------------------ forward
    @staticmethod
    def label_rows(labels: torch.Tensor, full: bool = False):
        """Rows (b*S + s) of the hidden states whose NEXT-token label is not -100, and those targets
        (HF shift: logits[:, :-1] vs labels[:, 1:]).  Works on CPU or GPU tensors; the train loop calls it on the
        CPU copy before the H2D so no device sync is needed."""
        B, S = labels.shape
        tgt = torch.full_like(labels, -100)
        tgt[:, :-1] = labels[:, 1:]
        if full:
            idx = torch.arange(B * S, device=labels.device)
        else:
            idx = torch.nonzero(tgt.reshape(-1) != -100).squeeze(1)
        return idx.to(torch.int32), tgt.reshape(-1)[idx.long()].contiguous()

    def forward(self, batch: Dict[str, torch.Tensor], train: bool = True, full_logits: bool = False):
        """batch: collator contract (input_ids, labels, attention_mask, modality_mask, audio_mel | audio_pcm) on device.
        Optional precomputed '_rows' (int32) / '_targets' (int64) avoid a device->host sync.
        Returns (loss, acc, logits_or_None); loss/acc are 0-dim device tensors."""
        enc_out = self.forward_front(batch)                                                # frozen: does not read the trainables
        self.flush_update()                                                                # (deferred mode) all-reduce wait + AdamW of the previous step
        return self.forward_rest(batch, enc_out, train=train, full_logits=full_logits)

    def forward_front(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        """log-mel (unless the batch carries audio_mel) + Whisper encoder: the part of the step that is independent of the trainables."""
        dev = self.device
        if self.encoder is None:
            raise RuntimeError("this step has no B200 encoder (foreign modality encoder): run it yourself and enter at forward_rest(batch, enc_out)")
        mel = batch.get("audio_mel")
        if mel is None:
            mel = self.log_mel(batch["audio_pcm"].to(dev, F32), batch.get("audio_pcm_lengths"))
        else:
            mel = mel.to(dev, F32)
        return self.encoder.forward(mel)

    def forward_rest(self, batch: Dict[str, torch.Tensor], enc_out: torch.Tensor, train: bool = True, full_logits: bool = False):
        dev = self.device
        ids = batch["input_ids"].to(dev).contiguous()
        labels = batch["labels"].to(dev)
        key_mask = batch["attention_mask"].to(dev).to(torch.uint8).contiguous()
        mod_mask = batch["modality_mask"].to(dev).to(torch.uint8).contiguous()
        B, S = ids.shape
        self.begin_decoder_pass(train)
        aud = self.projector.forward(enc_out, save=train)
        x = ops.embed_merge(ids, mod_mask, aud, self.llm.embed)
        out = self.decoder_loss(x, key_mask, labels, rows=batch.get("_rows"), targets=batch.get("_targets"), train=train, full_logits=full_logits)
        if train:
            self._ctx.update(mod_mask=mod_mask, Ta=aud.shape[1], ids=ids)
        return out

    def begin_decoder_pass(self, train: bool) -> None:
        """Once per forward, before anything reads the trainables: land a deferred update, re-pack the adapters, arm LoRA dropout."""
        self.flush_update()
        if self.llm.train_base:
            self.llm.refresh_base()                                                        # full fine-tune: bf16 operands from the fp32 masters
        self.llm.pack_lora()                                                               # adapters change every optimizer step
        self.llm.dropout_active = bool(train and self.lora_dropout_enabled and self.llm.dropout_p > 0.0)
        self.llm.dropout_step += 1

    def decoder_loss(self, x: torch.Tensor, key_mask: torch.Tensor, labels: torch.Tensor, rows=None, targets=None, train: bool = True,
                     full_logits: bool = False):
        """inputs_embeds bf16 [B,S,D] + key mask u8 [B,S] + labels -> (loss, acc, logits | None): the call `self.llm(inputs_embeds=...,
        attention_mask=..., labels=...)` of slam_model.py:400 incl. the accuracy of :402-405.  Recipes that build inputs_embeds themselves
        enter here (slam_model.llm_forward); decoder_backward() returns the gradient w.r.t. x."""
        dev = self.device
        B, S, _ = x.shape
        xf = self.llm.forward(x, key_mask, save=train)
        if rows is not None and not full_logits:
            rows, tgts = rows.to(dev), targets.to(dev)
        else:
            rows, tgts = self.label_rows(labels.to(dev), full=full_logits)
        R = rows.numel()
        hsel = xf if full_logits else ops.gather_rows(xf, rows)
        logits = ops.gemm(hsel, self.llm.lm_head, out_f32=True)                            # fp32 logits (HF .float())
        loss_sum = torch.zeros(1, device=dev, dtype=F32)
        n_valid = torch.zeros(1, device=dev, dtype=torch.int32)
        n_correct = torch.zeros(1, device=dev, dtype=torch.int32)
        ops.cross_entropy(logits, tgts, (loss_sum, n_valid, n_correct))
        nv = n_valid.to(F32)
        loss = (loss_sum / nv).squeeze(0)
        acc = (n_correct.to(F32) / nv).squeeze(0)
        if train:
            self._ctx = dict(logits=logits, tgts=tgts, rows=rows, full=full_logits, nv=nv, B=B, S=S, R=R, hsel=hsel if self.llm.train_base else None)
        return loss, acc, (logits.view(B, S, -1) if full_logits else None)

    def decoder_last_logits(self, x: torch.Tensor, key_mask: torch.Tensor) -> torch.Tensor:
        """Next-token logits of a batch of sequences (decode path, slam_model.generate): decoder forward on bf16 [n,S,D] with the key mask,
        lm_head on the LAST position of every sequence only -> f32 [n, V].  No KV cache: the caller passes the whole sequence each step."""
        n, S, _ = x.shape
        self.begin_decoder_pass(False)
        xf = self.llm.forward(x.to(self.device, BF16).contiguous(), key_mask.to(self.device).to(torch.uint8).contiguous(), save=False)
        rows = (torch.arange(n, device=self.device, dtype=torch.int32) + 1) * S - 1
        return ops.gemm(ops.gather_rows(xf, rows), self.llm.lm_head, out_f32=True)

    def decode_prefill(self, x: torch.Tensor, key_mask: torch.Tensor):
        """Prompt pass of a decode session: -> (next-token logits f32 [n, V], state).  state = dict(cache=[(K, V)] per layer, mask u8 [n, S])."""
        n, S, _ = x.shape
        self.begin_decoder_pass(False)
        key_mask = key_mask.to(self.device).to(torch.uint8).contiguous()
        cache: list = []
        xf = self.llm.forward(x.to(self.device, BF16).contiguous(), key_mask, save=False, kv_out=cache)
        rows = (torch.arange(n, device=self.device, dtype=torch.int32) + 1) * S - 1
        logits = ops.gemm(ops.gather_rows(xf, rows), self.llm.lm_head, out_f32=True)
        return logits, dict(cache=cache, mask=key_mask, pos=S)

    def decode_next(self, tokens: torch.Tensor, state: dict, beam_src: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Feed one token per sequence (i64 [n]) -> next-token logits f32 [n, V]; the KV cache in `state` grows by one position.
        beam_src (i64 [n]): row i continues the sequence that was row beam_src[i] of the previous call (beam re-ordering / expansion)."""
        dev = self.device
        if beam_src is not None:
            idx = beam_src.to(dev)
            state["cache"] = [(K.index_select(0, idx), V.index_select(0, idx)) for K, V in state["cache"]]
            state["mask"] = state["mask"].index_select(0, idx)
        n = tokens.shape[0]
        state["mask"] = torch.cat([state["mask"], torch.ones(n, 1, dtype=torch.uint8, device=dev)], dim=1).contiguous()
        x = torch.nn.functional.embedding(tokens.to(dev), self.llm.embed)
        xf = self.llm.decode_step(x.contiguous(), state["cache"], state["mask"], state["pos"])
        state["pos"] += 1
        return ops.gemm(xf, self.llm.lm_head, out_f32=True)

    # ------------------------------------------------------------------ backward
    def backward(self, grad_out: Optional[torch.Tensor] = None) -> None:
        """Backward of the last forward(train=True); grad_out is d(total)/d(loss) (device scalar, default 1)."""
        mod_mask, ta, ids = self._ctx["mod_mask"], self._ctx["Ta"], self._ctx["ids"]
        self.backward_begin()
        dx = self.decoder_backward(grad_out)
        if self.llm.train_base:                                                           # embedding rows used by the merge (slam_model.py:392)
            ops.embed_grad(ids, mod_mask, dx.contiguous(), self.arena.view(LlamaLoRAB200.BASE_PREFIX + "model.embed_tokens.weight", "grad"))
        daud = ops.embed_merge_bwd(mod_mask, dx.contiguous(), ta)
        self.projector.backward(daud)
        self.backward_end()

    def backward_begin(self) -> None:
        self.flush_update()                                                               # never overwrite gradients an update still needs
        self._carry = self.arena.grad.clone() if self.micro_steps > 0 else None           # gradient accumulation: kernels overwrite
        self.arena.grad.zero_()                                                           # one memset: the LoRA wgrad products accumulate

    def backward_end(self) -> None:
        if self._carry is not None:
            self.arena.grad.add_(self._carry)
            self._carry = None
        self.micro_steps += 1

    def decoder_backward(self, grad_out: Optional[torch.Tensor] = None, grad_logits: Optional[torch.Tensor] = None) -> torch.Tensor:
        """CE + lm_head + decoder backward of the last decoder_loss(train=True): LoRA gradients into the arena, returns d loss / d inputs_embeds
        bf16 [B,S,D] (call between backward_begin() and backward_end()).  grad_logits (f32 [B,S,V] or [B*S,V], full-logits mode only): an
        upstream gradient w.r.t. the returned logits (recipes that compute their own loss from `outputs.logits`, e.g. the s2s group CE);
        it is added to the CE term (grad_out = 0 drops the CE term)."""
        c = self._ctx
        self._ctx = None
        dev = self.device
        gs = (1.0 / c["nv"]) if grad_out is None else (grad_out.to(dev, F32).reshape(1) / c["nv"])
        gs = gs.contiguous()
        logits = c["logits"]
        R, V = logits.shape
        dlogits = torch.empty((R, V), device=dev, dtype=BF16)
        scratch = (torch.zeros(1, device=dev), torch.zeros(1, device=dev, dtype=torch.int32), torch.zeros(1, device=dev, dtype=torch.int32))
        ops.cross_entropy(logits, c["tgts"], scratch, dlogits, gs)
        if grad_logits is not None:
            if not c["full"]:
                raise RuntimeError("a gradient w.r.t. logits needs the full-logits mode (train_config.b200_full_logits=true)")
            dlogits = ops.add(dlogits, ops.cast_bf16(grad_logits.to(dev, F32).reshape(R, V).contiguous()))
        if self.llm.train_base:                                                            # lm_head weight gradient = dlogits^T h  (tied: lands in dE)
            head = "model.embed_tokens.weight" if self.llm.lm_head is self.llm.embed else "lm_head.weight"
            self.llm._wgrad(dlogits, c["hsel"], head, V, self.llm_cfg.d)
        dh = _gemm_few_tiles(dlogits, self.llm.lm_headT)                                   # [R, D], K = vocab: split-K
        M = c["B"] * c["S"]
        if c["full"]:
            dxf = dh
        else:
            dxf = torch.zeros((M, self.llm_cfg.d), device=dev, dtype=BF16)
            ops.scatter_rows(dh, c["rows"], dxf)
        return self.llm.backward(dxf)

    def allreduce_grads(self, async_op: bool = False):
        """The one data-path collective (DDP's bucket all-reduce, pipeline/finetune.py:181-184): SUM over ranks of the flat
        gradient arena (AdamW divides by the world size).  async_op: NCCL runs it on its own stream behind the kernels already queued;
        the Work is kept and waited for (stream-side, no host block) by flush_update()."""
        if async_op:
            self._pending_work = torch.distributed.all_reduce(self.arena.grad, async_op=True)
        else:
            torch.distributed.all_reduce(self.arena.grad)

    def optimizer_step(self, lr: float, weight_decay: float = 0.0, grad_div: float = 1.0, betas=(0.9, 0.999), eps: float = 1e-8) -> None:
        self._pending_update = (lr, weight_decay, grad_div, betas, eps)
        self.micro_steps = 0
        if not self.defer_update:
            self.flush_update()

    def flush_update(self) -> None:
        """Apply the recorded optimizer step (no-op when none is pending).  Everything that reads the trainables calls this first."""
        if self._pending_work is not None:
            self._pending_work.wait()                                                      # current stream waits for the NCCL stream
            self._pending_work = None
        if self._pending_update is not None:
            lr, weight_decay, grad_div, betas, eps = self._pending_update
            self._pending_update = None
            self.arena.adamw_step(lr, weight_decay, betas=betas, eps=eps, grad_div=grad_div)

    def train_step(self, batch, lr: float = 1e-4, weight_decay: float = 0.0, world_size: int = 1):
        loss, acc, _ = self.forward(batch, train=True)
        self.backward()
        if world_size > 1:
            self.allreduce_grads(async_op=self.defer_update)
        self.optimizer_step(lr, weight_decay, grad_div=float(world_size))
        return loss, acc
