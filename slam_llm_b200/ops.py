"""torch.Tensor-level wrappers over the C ABI (pointers + sizes + current CUDA stream).

PyTorch is used here only for device memory and streams; every function launches kernels from
libslam_b200.so.  Nothing in this module computes on the CPU or falls back to torch ops.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import os

import torch

from . import lib as _l

BF16 = torch.bfloat16
F32 = torch.float32


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"slam_b200.{name}: tensor must be on a CUDA device (no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"slam_b200.{name}: expected {dtype}, got {t.dtype}")


def _row_major_2d(t: torch.Tensor, name: str) -> int:
    """Return the leading dimension of a 2-D row-major (possibly strided-rows) tensor."""
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"slam_b200.{name}: need a 2-D tensor with contiguous last dim, got {tuple(t.shape)}/{t.stride()}")
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def launch_count() -> int:
    return int(_l.load().slam_launch_count())


_GEMM_LOG = None


def set_gemm_event_log(log) -> None:
    """bench.py instrumentation: when `log` is a list, every gemm() appends (algorithmic_flops, start_event, end_event)
    recorded on the launching stream (no synchronisation here)."""
    global _GEMM_LOG
    _GEMM_LOG = log


# ------------------------------------------------------------------------------------------- GEMM
_GEMM_WS: dict = {}
_TAIL_SPLIT_DEFAULT = int(os.environ.get("SLAM_TAIL_SPLIT", "0"))     # experiments: -1 disables the GEMM tail split globally


def _gemm_workspace(device: torch.device) -> torch.Tensor:
    """Tail-split exchange buffer: zeroed once, one per (device, stream) - the library owns it between calls on that stream."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream())
    ws = _GEMM_WS.get(key)
    if ws is None:
        ws = torch.zeros(int(_l.load().slam_gemm_workspace_bytes()), dtype=torch.uint8, device=device)
        _GEMM_WS[key] = ws
    return ws


def gemm(a: torch.Tensor, b: torch.Tensor, *, out: Optional[torch.Tensor] = None, a2: Optional[torch.Tensor] = None,
         b2: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
         act: int = 0, alpha: float = 1.0, out_f32: bool = False, block_n: int = 0, split_k: int = 1, tail_split: int = 0,
         aux: Optional[torch.Tensor] = None, transpose_out: bool = False, static_w: bool = False) -> torch.Tensor:
    """out[M,N] = act(alpha*(a @ b.T + a2 @ b2.T) + bias) + residual ; a [M,K], b [N,K] bf16.
    tail_split: 0 = automatic, -1 = off, n > 1 = at most n k-slices per tail tile.
    act 3 / 4 = fused SwiGLU forward / backward on the blocked-64 gate/up layout (see slam_gemm_args.aux): act 3 returns gu and
    writes h = silu(g) * u to aux [M, N/2]; act 4 reads gu from aux [M, 2N] and returns d(gu) [M, 2N] (the product dh is not stored).
    static_w: the weight operand (`b`; `a` with transpose_out) is frozen - never written by a kernel that may still be in flight
    (slam_gemm_args.static_operands): its first tiles are then requested under the tail of the preceding kernel."""
    _req(a, BF16, "gemm.a"); _req(b, BF16, "gemm.b")
    M, K1 = a.shape
    N = b.shape[0]
    assert b.shape[1] == K1, (a.shape, b.shape)
    if transpose_out:
        return _gemm_swapped(a, b, out, a2, b2, residual, block_n, static_w, act, aux)
    if split_k > 1:
        assert out_f32 and bias is None and residual is None and act == 0, "split_k needs a plain f32 output"
        if out is None:
            out = torch.zeros((M, N), device=a.device, dtype=F32)          # k-slices are merged with atomics
    out_cols = 2 * N if act == 4 else N
    if out is None:
        out = torch.empty((M, out_cols), device=a.device, dtype=F32 if out_f32 else BF16)
    g = _l.GemmArgs()
    g.a, g.lda = a.data_ptr(), _row_major_2d(a, "gemm.a")
    g.b, g.ldb = b.data_ptr(), _row_major_2d(b, "gemm.b")
    g.k1 = K1
    if a2 is not None:
        _req(a2, BF16, "gemm.a2"); _req(b2, BF16, "gemm.b2")
        assert a2.shape[0] == M and b2.shape[0] == N and a2.shape[1] == b2.shape[1]
        g.k2 = a2.shape[1]
        g.a2, g.lda2 = a2.data_ptr(), _row_major_2d(a2, "gemm.a2")
        g.b2, g.ldb2 = b2.data_ptr(), _row_major_2d(b2, "gemm.b2")
    else:
        g.k2 = 0
    _req(out, F32 if out_f32 else BF16, "gemm.out")
    assert tuple(out.shape) == (M, out_cols)
    g.out, g.ldo = out.data_ptr(), _row_major_2d(out, "gemm.out")
    g.out_f32 = 1 if out_f32 else 0
    g.act = act
    if bias is not None:
        _req(bias, F32, "gemm.bias")
        assert bias.numel() == N and bias.is_contiguous()
    g.bias = _p(bias)
    if residual is not None:
        _req(residual, BF16, "gemm.residual")
        assert tuple(residual.shape) == (M, N)
        g.residual, g.ldr = residual.data_ptr(), _row_major_2d(residual, "gemm.residual")
    else:
        g.residual, g.ldr = None, 0
    g.alpha = alpha
    g.m, g.n = M, N
    if act >= 3:
        _req(aux, BF16, "gemm.aux")
        assert tuple(aux.shape) == ((M, N // 2) if act == 3 else (M, 2 * N)), (aux.shape, M, N, act)
        g.aux, g.ld_aux = aux.data_ptr(), _row_major_2d(aux, "gemm.aux")
    else:
        g.aux, g.ld_aux = None, 0
    g.block_n = block_n
    g.transpose_out = 0
    g.static_operands = 2 if static_w else 0
    g.split_k = split_k
    if tail_split == 0:
        tail_split = _TAIL_SPLIT_DEFAULT
    g.tail_split = tail_split
    if tail_split >= 0 and split_k <= 1:
        ws = _gemm_workspace(a.device)
        g.workspace, g.workspace_bytes = ws.data_ptr(), ws.numel()
    else:
        g.workspace, g.workspace_bytes = None, 0
    if _GEMM_LOG is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _l.check(_l.load().slam_gemm_bf16(C.byref(g), _stream()), "slam_gemm_bf16")
        e1.record()
        _GEMM_LOG.append((2.0 * M * N * (K1 + g.k2), e0, e1))
        return out
    _l.check(_l.load().slam_gemm_bf16(C.byref(g), _stream()), "slam_gemm_bf16")
    return out


def _gemm_swapped(w, x, out, w2, x2, residual, block_n, static_w=False, act=0, aux=None):
    """transpose_out (swap-AB): returns y[Mx, Nw] = x @ w.T (+ x2 @ w2.T) (+ residual[Mx, Nw]) computed as tiles of (w @ x.T): the weight `w` [Nw, K]
    is the M operand (its rows fill 256-row CTA-pair tiles exactly), the activations `x` [Mx, K] the N operand.
    act 4 (fused SwiGLU backward): the product is dh [Mx, Nw = F] (never stored), aux = gu [Mx, 2F], returns d(gu) [Mx, 2F]."""
    Nw, K1 = w.shape
    Mx = x.shape[0]
    assert act in (0, 4), "transpose_out supports act 0 and 4"
    out_cols = 2 * Nw if act == 4 else Nw
    if out is None:
        out = torch.empty((Mx, out_cols), device=w.device, dtype=BF16)
    _req(out, BF16, "gemm.out")
    assert tuple(out.shape) == (Mx, out_cols) and out.stride(1) == 1
    g = _l.GemmArgs()
    g.a, g.lda = w.data_ptr(), _row_major_2d(w, "gemm.a")
    g.b, g.ldb = x.data_ptr(), _row_major_2d(x, "gemm.b")
    g.k1 = K1
    if w2 is not None:
        _req(w2, BF16, "gemm.a2"); _req(x2, BF16, "gemm.b2")
        assert w2.shape[0] == Nw and x2.shape[0] == Mx and w2.shape[1] == x2.shape[1]
        g.k2 = w2.shape[1]
        g.a2, g.lda2 = w2.data_ptr(), _row_major_2d(w2, "gemm.a2")
        g.b2, g.ldb2 = x2.data_ptr(), _row_major_2d(x2, "gemm.b2")
    else:
        g.k2 = 0
    g.out, g.ldo = out.data_ptr(), _row_major_2d(out, "gemm.out")
    g.out_f32, g.act, g.bias, g.alpha = 0, act, None, 1.0
    if residual is not None:
        _req(residual, BF16, "gemm.residual")
        assert tuple(residual.shape) == (Mx, Nw)
        g.residual, g.ldr = residual.data_ptr(), _row_major_2d(residual, "gemm.residual")
    else:
        g.residual, g.ldr = None, 0
    g.m, g.n = Nw, Mx
    if act == 4:
        _req(aux, BF16, "gemm.aux")
        assert tuple(aux.shape) == (Mx, 2 * Nw) and residual is None and w2 is None
        g.aux, g.ld_aux = aux.data_ptr(), _row_major_2d(aux, "gemm.aux")
    else:
        g.aux, g.ld_aux = None, 0
    g.block_n, g.split_k, g.tail_split, g.transpose_out = block_n, 1, -1, 1
    g.static_operands = 1 if static_w else 0
    g.workspace, g.workspace_bytes = None, 0
    if _GEMM_LOG is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _l.check(_l.load().slam_gemm_bf16(C.byref(g), _stream()), "slam_gemm_bf16")
        e1.record()
        _GEMM_LOG.append((2.0 * Mx * Nw * (K1 + g.k2), e0, e1))
        return out
    _l.check(_l.load().slam_gemm_bf16(C.byref(g), _stream()), "slam_gemm_bf16")
    return out


def wgrad_thin(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, scale: float = 1.0, accumulate: bool = False) -> torch.Tensor:
    """out[P,Q] (f32) = (out if accumulate else 0) + scale * a[M,P].T @ b[M,Q]   (P <= 64)."""
    _req(a, BF16, "wgrad_thin.a"); _req(b, BF16, "wgrad_thin.b"); _req(out, F32, "wgrad_thin.out")
    M, P = a.shape
    Q = b.shape[1]
    assert b.shape[0] == M and tuple(out.shape) == (P, Q)
    _l.check(_l.load().slam_wgrad_thin(a.data_ptr(), _row_major_2d(a, "wgrad.a"), P, b.data_ptr(), _row_major_2d(b, "wgrad.b"), Q, M,
                                       scale, out.data_ptr(), _row_major_2d(out, "wgrad.out"), 1 if accumulate else 0, _stream()), "slam_wgrad_thin")
    return out


# ------------------------------------------------------------------------------------------- front end
def logmel(wav: torch.Tensor, filters_t: torch.Tensor, out: Optional[torch.Tensor] = None, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
    """wav f32 [B, n_samples] -> log-mel f32 [B, n_samples//160, n_mels] (whisper.log_mel_spectrogram, time-major).
    lengths (i32 [B], optional): real sample counts of a zero-padded variable-length batch."""
    _req(wav, F32, "logmel.wav"); _req(filters_t, F32, "logmel.filters_t")
    assert wav.dim() == 2 and wav.is_contiguous() and filters_t.is_contiguous() and filters_t.shape[0] == 201
    B, n = wav.shape
    n_mels = filters_t.shape[1]
    if out is None:
        out = torch.empty((B, n // 160, n_mels), device=wav.device, dtype=F32)
    scratch = torch.empty((B,), device=wav.device, dtype=F32)
    if lengths is not None:
        assert lengths.dtype == torch.int32 and lengths.is_cuda and lengths.numel() == B and lengths.is_contiguous()
    _l.check(_l.load().slam_logmel(wav.data_ptr(), B, n, _p(lengths), filters_t.data_ptr(), n_mels, out.data_ptr(), scratch.data_ptr(), _stream()),
             "slam_logmel")
    return out


def conv_im2col(x: torch.Tensor, stride: int, ldk: int) -> torch.Tensor:
    """x [B,T,C] (f32|bf16) -> bf16 [B*T_out, ldk] patches for Conv1d(k=3, pad=1, stride)."""
    assert x.dim() == 3 and x.is_contiguous() and x.is_cuda
    B, T, Cc = x.shape
    t_out = (T + 2 - 3) // stride + 1
    col = torch.empty((B * t_out, ldk), device=x.device, dtype=BF16)
    _l.check(_l.load().slam_conv_im2col(x.data_ptr(), 1 if x.dtype == F32 else 0, B, T, Cc, stride, col.data_ptr(), ldk, _stream()),
             "slam_conv_im2col")
    return col


def add_pos_(x: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
    _req(x, BF16, "add_pos.x"); _req(pos, F32, "add_pos.pos")
    B, T, D = x.shape
    assert x.is_contiguous() and pos.is_contiguous() and pos.shape[0] >= T and pos.shape[1] == D
    _l.check(_l.load().slam_add_pos(x.data_ptr(), pos.data_ptr(), B, T, D, _stream()), "slam_add_pos")
    return x


def layernorm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-5, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(x, BF16, "layernorm.x"); _req(w, F32, "layernorm.w"); _req(b, F32, "layernorm.b")
    assert x.is_contiguous()
    d = x.shape[-1]
    rows = x.numel() // d
    if out is None:
        out = torch.empty_like(x)
    _l.check(_l.load().slam_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), rows, d, eps, _stream()), "slam_layernorm")
    return out


# ------------------------------------------------------------------------------------------- attention
def _attn_args(q, k, v, out, lse, key_mask, causal, scale):
    # q [B,Sq,Hq,dh] / k,v [B,Sk,Hkv,dh]; last two dims contiguous, token stride arbitrary (fused QKV buffers)
    B, Sq, Hq, dh = q.shape
    Sk, Hkv = k.shape[1], k.shape[2]
    for t, nm in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        _req(t, BF16, f"attn.{nm}")
        assert t.stride(3) == 1 and t.stride(2) == dh, f"attn.{nm}: heads must be packed"
        assert t.stride(0) == t.stride(1) * t.shape[1], f"attn.{nm}: batch must be packed over tokens"
    a = _l.AttnArgs()
    a.q, a.ldq = q.data_ptr(), q.stride(1)
    a.k, a.ldk = k.data_ptr(), k.stride(1)
    a.v, a.ldv = v.data_ptr(), v.stride(1)
    a.out, a.ldo = out.data_ptr(), out.stride(1)
    a.lse = _p(lse)
    if key_mask is not None:
        assert key_mask.dtype == torch.uint8 and key_mask.is_contiguous() and tuple(key_mask.shape) == (B, Sk)
    a.key_mask = _p(key_mask)
    a.batch, a.sq, a.sk, a.hq, a.hkv, a.dh = B, Sq, Sk, Hq, Hkv, dh
    a.causal = 1 if causal else 0
    a.scale = scale
    return a


def attn_fwd(q, k, v, *, causal: bool, scale: float, key_mask=None, out=None, need_lse: bool = False):
    B, Sq, Hq, dh = q.shape
    if out is None:
        out = torch.empty((B, Sq, Hq, dh), device=q.device, dtype=BF16)
    lse = torch.empty((B, Hq, Sq), device=q.device, dtype=F32) if need_lse else None
    a = _attn_args(q, k, v, out, lse, key_mask, causal, scale)
    _l.check(_l.load().slam_attn_fwd(C.byref(a), _stream()), "slam_attn_fwd")
    return out, lse


def attn_bwd(q, k, v, out, lse, dout, *, causal: bool, scale: float, key_mask=None, dq=None, dk=None, dv=None, rope=None):
    """rope = (cos, sin) f32 [S, dh/2]: q / k are the ROTATED projections; dq / dk come back w.r.t. the un-rotated ones (inverse RoPE fused
    into the finishing kernel together with the GQA group sum and the dQ fp32 -> bf16 conversion)."""
    B, Sq, Hq, dh = q.shape
    Sk, Hkv = k.shape[1], k.shape[2]
    if dq is None:
        dq = torch.empty((B, Sq, Hq, dh), device=q.device, dtype=BF16)
    if dk is None:
        dk = torch.empty((B, Sk, Hkv, dh), device=q.device, dtype=BF16)
    if dv is None:
        dv = torch.empty((B, Sk, Hkv, dh), device=q.device, dtype=BF16)
    a = _attn_args(q, k, v, out, lse, key_mask, causal, scale)
    for t, nm in ((dout, "dout"), (dq, "dq"), (dk, "dk"), (dv, "dv")):
        _req(t, BF16, f"attn.{nm}")
        assert t.stride(3) == 1 and t.stride(2) == dh and t.stride(0) == t.stride(1) * t.shape[1]
    a.dout, a.lddo = dout.data_ptr(), dout.stride(1)
    a.dq, a.lddq = dq.data_ptr(), dq.stride(1)
    a.dk, a.lddk = dk.data_ptr(), dk.stride(1)
    a.dv, a.lddv = dv.data_ptr(), dv.stride(1)
    delta = torch.empty((B, Hq, Sq), device=q.device, dtype=F32)
    dq_accum = torch.empty((B, Sq, Hq, dh), device=q.device, dtype=F32)
    a.delta, a.dq_accum = delta.data_ptr(), dq_accum.data_ptr()
    dkv_part = torch.empty((2, B, Sk, Hq, dh), device=q.device, dtype=BF16) if Hq != Hkv else None   # GQA: split CTAs per Q head
    a.dkv_part = _p(dkv_part)
    if rope is not None:
        cos, sin = rope
        _req(cos, F32, "attn_bwd.rope_cos"); _req(sin, F32, "attn_bwd.rope_sin")
        assert cos.is_contiguous() and sin.is_contiguous() and cos.shape[0] >= max(Sq, Sk) and cos.shape[1] == dh // 2 and sin.shape == cos.shape
        a.rope_cos, a.rope_sin = cos.data_ptr(), sin.data_ptr()
    else:
        a.rope_cos = a.rope_sin = None
    _l.check(_l.load().slam_attn_bwd(C.byref(a), _stream()), "slam_attn_bwd")
    return dq, dk, dv


# ------------------------------------------------------------------------------------------- merge / decoder element-wise
_DEBUG_CHECKS = os.environ.get("SLAM_DEBUG_CHECKS", "0") == "1"


def embed_merge(ids, modality_mask, audio, embed, out=None):
    """SLAM_DEBUG_CHECKS=1 adds the host-side validation the reference gets for free from torch indexing (costs a device sync): token ids must
    be < vocab and the audio span of every row must fit the projector output — otherwise the kernel clamps the span / reads what it is given."""
    assert ids.dtype == torch.int64 and modality_mask.dtype == torch.uint8
    if _DEBUG_CHECKS:
        if int(ids.max()) >= embed.shape[0]:
            raise IndexError(f"embed_merge: token id {int(ids.max())} >= vocabulary {embed.shape[0]}")
        if int(modality_mask.sum(dim=1).max()) > audio.shape[1]:
            raise ValueError(f"embed_merge: a modality span of {int(modality_mask.sum(dim=1).max())} positions exceeds the {audio.shape[1]} audio tokens "
                             "the projector produced (fix_length_audio / modality_mask mismatch)")
    _req(audio, BF16, "embed_merge.audio"); _req(embed, BF16, "embed_merge.embed")
    B, S = ids.shape
    Ta, D = audio.shape[1], audio.shape[2]
    assert ids.is_contiguous() and modality_mask.is_contiguous() and audio.is_contiguous() and embed.is_contiguous()
    if out is None:
        out = torch.empty((B, S, D), device=audio.device, dtype=BF16)
    _l.check(_l.load().slam_embed_merge(ids.data_ptr(), modality_mask.data_ptr(), audio.data_ptr(), Ta, embed.data_ptr(), out.data_ptr(),
                                        B, S, D, _stream()), "slam_embed_merge")
    return out


def embed_merge_bwd(modality_mask, dx, ta: int):
    B, S, D = dx.shape
    _req(dx, BF16, "embed_merge_bwd.dx")
    assert dx.is_contiguous()
    daudio = torch.empty((B, ta, D), device=dx.device, dtype=BF16)
    _l.check(_l.load().slam_embed_merge_bwd(modality_mask.data_ptr(), dx.data_ptr(), daudio.data_ptr(), ta, B, S, D, _stream()),
             "slam_embed_merge_bwd")
    return daudio


def rmsnorm_fwd(x, w, eps: float, out=None, need_rstd: bool = True):
    _req(x, BF16, "rmsnorm.x"); _req(w, BF16, "rmsnorm.w")
    assert x.is_contiguous()
    d = x.shape[-1]
    rows = x.numel() // d
    if out is None:
        out = torch.empty_like(x)
    rstd = torch.empty((rows,), device=x.device, dtype=F32) if need_rstd else None
    _l.check(_l.load().slam_rmsnorm_fwd(x.data_ptr(), w.data_ptr(), out.data_ptr(), _p(rstd), rows, d, eps, _stream()), "slam_rmsnorm_fwd")
    return out, rstd


def rmsnorm_bwd(dy, x, w, rstd, dres=None, out=None):
    _req(dy, BF16, "rmsnorm_bwd.dy"); _req(x, BF16, "rmsnorm_bwd.x")
    assert dy.is_contiguous() and x.is_contiguous() and (dres is None or dres.is_contiguous())
    d = x.shape[-1]
    rows = x.numel() // d
    if out is None:
        out = torch.empty_like(x)
    _l.check(_l.load().slam_rmsnorm_bwd(dy.data_ptr(), x.data_ptr(), w.data_ptr(), rstd.data_ptr(), _p(dres), out.data_ptr(), rows, d,
                                        _stream()), "slam_rmsnorm_bwd")
    return out


def rope_(x: torch.Tensor, n_heads: int, dh: int, seq_len: int, cos: torch.Tensor, sin: torch.Tensor, inverse: bool = False):
    """In-place RoPE on x viewed as [rows, n_heads, dh] with row stride x.stride(0) (x is a 2-D column slice)."""
    _req(x, BF16, "rope.x"); _req(cos, F32, "rope.cos"); _req(sin, F32, "rope.sin")
    assert x.dim() == 2 and x.stride(1) == 1 and x.shape[1] == n_heads * dh
    assert cos.is_contiguous() and sin.is_contiguous() and cos.shape[0] >= seq_len and cos.shape[1] == dh // 2
    _l.check(_l.load().slam_rope(x.data_ptr(), x.stride(0), x.shape[0], seq_len, n_heads, dh, cos.data_ptr(), sin.data_ptr(),
                                 1 if inverse else 0, _stream()), "slam_rope")
    return x


def swiglu_fwd(gu, out=None, block: int = 0):
    _req(gu, BF16, "swiglu.gu")
    assert gu.dim() == 2 and gu.is_contiguous()
    rows, f2 = gu.shape
    if out is None:
        out = torch.empty((rows, f2 // 2), device=gu.device, dtype=BF16)
    _l.check(_l.load().slam_swiglu_fwd(gu.data_ptr(), out.data_ptr(), rows, f2 // 2, block, _stream()), "slam_swiglu_fwd")
    return out


def swiglu_bwd(gu, dh, out=None, block: int = 0):
    _req(gu, BF16, "swiglu_bwd.gu"); _req(dh, BF16, "swiglu_bwd.dh")
    assert gu.is_contiguous() and dh.is_contiguous()
    rows, f2 = gu.shape
    if out is None:
        out = torch.empty_like(gu)
    _l.check(_l.load().slam_swiglu_bwd(gu.data_ptr(), dh.data_ptr(), out.data_ptr(), rows, f2 // 2, block, _stream()), "slam_swiglu_bwd")
    return out


def dropout(x, p: float, seed: int, out=None):
    """y = x * keep/(1-p); keep regenerated from (seed, element index) — see dropout_bwd_add."""
    _req(x, BF16, "dropout.x")
    assert x.is_contiguous()
    if out is None:
        out = torch.empty_like(x)
    _l.check(_l.load().slam_dropout(x.data_ptr(), out.data_ptr(), x.numel(), p, seed, _stream()), "slam_dropout")
    return out


def dropout_bwd_add(base, lora, p: float, seed: int, out=None):
    """out = base + lora * keep/(1-p) with the mask of dropout(seed)."""
    _req(base, BF16, "dropout_bwd_add.base"); _req(lora, BF16, "dropout_bwd_add.lora")
    assert base.is_contiguous() and lora.is_contiguous() and base.numel() == lora.numel()
    if out is None:
        out = torch.empty_like(base)
    _l.check(_l.load().slam_dropout_bwd_add(base.data_ptr(), lora.data_ptr(), out.data_ptr(), base.numel(), p, seed, _stream()),
             "slam_dropout_bwd_add")
    return out


# ------------------------------------------------------------------------------------------- loss / optimizer
def cross_entropy(logits, targets, stats, dlogits=None, grad_scale=None):
    """stats: (loss_sum f32[1], n_valid i32[1], n_correct i32[1]) device tensors, accumulated atomically."""
    _req(logits, F32, "ce.logits")
    assert targets.dtype == torch.int64 and targets.is_contiguous()
    R, V = logits.shape
    loss_sum, n_valid, n_correct = stats
    _l.check(_l.load().slam_cross_entropy(logits.data_ptr(), _row_major_2d(logits, "ce.logits"), targets.data_ptr(), R, V,
                                          loss_sum.data_ptr(), n_valid.data_ptr(), n_correct.data_ptr(), _p(dlogits),
                                          0 if dlogits is None else _row_major_2d(dlogits, "ce.dlogits"), _p(grad_scale), _stream()),
             "slam_cross_entropy")
    return dlogits


def adamw_(param, grad, exp_avg, exp_avg_sq, *, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, step=1, grad_div=1.0):
    for t in (param, grad, exp_avg, exp_avg_sq):
        _req(t, F32, "adamw")
        assert t.is_contiguous() and t.numel() == param.numel()
    _l.check(_l.load().slam_adamw(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), param.numel(), lr, beta1,
                                  beta2, eps, weight_decay, step, grad_div, _stream()), "slam_adamw")
    return param


# ------------------------------------------------------------------------------------------- utilities
def cast_bf16(x, out=None, scale: float = 1.0):
    _req(x, F32, "cast_bf16")
    assert x.is_contiguous()
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=BF16)
    _l.check(_l.load().slam_cast_f32_to_bf16(x.data_ptr(), out.data_ptr(), x.numel(), scale, _stream()), "slam_cast_f32_to_bf16")
    return out


def cast_f32(x, out=None):
    _req(x, BF16, "cast_f32")
    assert x.is_contiguous()
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=F32)
    _l.check(_l.load().slam_cast_bf16_to_f32(x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "slam_cast_bf16_to_f32")
    return out


def transpose(x, out=None):
    _req(x, BF16, "transpose")
    R, Cc = x.shape
    if out is None:
        out = torch.empty((Cc, R), device=x.device, dtype=BF16)
    _l.check(_l.load().slam_transpose_bf16(x.data_ptr(), _row_major_2d(x, "transpose.x"), out.data_ptr(), _row_major_2d(out, "transpose.out"),
                                           R, Cc, _stream()), "slam_transpose_bf16")
    return out


def transpose_f32_batched(src, dst, batch: int, rows: int, cols: int):
    """dst[b][j][i] = src[b][i][j] on contiguous f32 buffers."""
    _req(src, F32, "transpose_f32_batched.src"); _req(dst, F32, "transpose_f32_batched.dst")
    assert src.is_contiguous() and dst.is_contiguous() and src.numel() == dst.numel() == batch * rows * cols
    _l.check(_l.load().slam_transpose_f32_batched(src.data_ptr(), dst.data_ptr(), batch, rows, cols, _stream()), "slam_transpose_f32_batched")
    return dst


def gather_rows(x, idx, out=None):
    _req(x, BF16, "gather_rows")
    assert idx.dtype == torch.int32 and x.is_contiguous()
    d = x.shape[-1]
    if out is None:
        out = torch.empty((idx.numel(), d), device=x.device, dtype=BF16)
    _l.check(_l.load().slam_gather_rows(x.data_ptr(), idx.data_ptr(), out.data_ptr(), idx.numel(), d, _stream()), "slam_gather_rows")
    return out


def scatter_rows(x, idx, out):
    _req(x, BF16, "scatter_rows")
    assert idx.dtype == torch.int32 and x.is_contiguous() and out.is_contiguous()
    _l.check(_l.load().slam_scatter_rows(x.data_ptr(), idx.data_ptr(), out.data_ptr(), idx.numel(), x.shape[-1], _stream()), "slam_scatter_rows")
    return out


def relu_bwd(dy, y, out=None):
    _req(dy, BF16, "relu_bwd")
    assert dy.is_contiguous() and y.is_contiguous()
    if out is None:
        out = torch.empty_like(dy)
    _l.check(_l.load().slam_relu_bwd(dy.data_ptr(), y.data_ptr(), out.data_ptr(), dy.numel(), _stream()), "slam_relu_bwd")
    return out


def colsum(x, out):
    _req(x, BF16, "colsum"); _req(out, F32, "colsum.out")
    R, Cc = x.shape
    _l.check(_l.load().slam_colsum(x.data_ptr(), _row_major_2d(x, "colsum.x"), R, Cc, out.data_ptr(), _stream()), "slam_colsum")
    return out


def colsum_add(x, out):
    """out[c] += sum_r x[r,c] (no zeroing: accumulates into a gradient buffer that the caller cleared)."""
    tmp = torch.empty_like(out)
    colsum(x, tmp)
    out.add_(tmp)
    return out


def rmsnorm_wgrad(dy, x, rstd, dw):
    """dw[c] += sum_r dy[r,c] * x[r,c] * rstd[r] (f32, accumulated)."""
    _req(dy, BF16, "rmsnorm_wgrad.dy"); _req(x, BF16, "rmsnorm_wgrad.x"); _req(rstd, F32, "rmsnorm_wgrad.rstd"); _req(dw, F32, "rmsnorm_wgrad.dw")
    assert dy.is_contiguous() and x.is_contiguous() and dw.is_contiguous()
    d = x.shape[-1]
    rows = x.numel() // d
    assert dw.numel() == d and rstd.numel() == rows
    _l.check(_l.load().slam_rmsnorm_wgrad(dy.data_ptr(), x.data_ptr(), rstd.data_ptr(), rows, d, dw.data_ptr(), _stream()), "slam_rmsnorm_wgrad")
    return dw


def embed_grad(ids, modality_mask, dx, de):
    """dE[ids[r]] += dx[r] for rows with modality_mask == 0 (f32 atomics)."""
    assert ids.dtype == torch.int64 and modality_mask.dtype == torch.uint8 and ids.is_contiguous() and modality_mask.is_contiguous()
    _req(dx, BF16, "embed_grad.dx"); _req(de, F32, "embed_grad.de")
    assert dx.is_contiguous() and de.is_contiguous()
    d = dx.shape[-1]
    rows = dx.numel() // d
    assert ids.numel() == rows and de.shape[1] == d
    _l.check(_l.load().slam_embed_grad(ids.data_ptr(), modality_mask.data_ptr(), dx.data_ptr(), de.data_ptr(), rows, d, de.shape[0], _stream()),
             "slam_embed_grad")
    return de


def pack2d(src, dst, *, batch, rows, cols, src_bs, src_ld, dst_bs, dst_ld, scale=1.0, transpose=False, src_off=0, dst_off=0):
    """Batched strided f32->bf16 cast with optional transpose; offsets in elements."""
    _req(src, F32, "pack2d.src"); _req(dst, BF16, "pack2d.dst")
    _l.check(_l.load().slam_pack2d(src.data_ptr() + 4 * src_off, src_bs, src_ld, dst.data_ptr() + 2 * dst_off, dst_bs, dst_ld, batch, rows, cols,
                                   scale, 1 if transpose else 0, _stream()), "slam_pack2d")
    return dst


def add(a, b, out=None):
    _req(a, BF16, "add")
    assert a.is_contiguous() and b.is_contiguous()
    if out is None:
        out = torch.empty_like(a)
    _l.check(_l.load().slam_add_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _stream()), "slam_add_bf16")
    return out
