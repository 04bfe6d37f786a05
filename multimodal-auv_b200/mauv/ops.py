"""Tensor-level wrappers over the C-ABI: torch supplies device memory and the stream,
every computation happens in libmauv_b200.so. No wrapper has a PyTorch fallback."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib

F16, F32, I64 = torch.float16, torch.float32, torch.int64


def _ptr(t: Optional[torch.Tensor], dtype=None) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.MauvError("mauv ops take CUDA tensors only (no CPU path)")
    if not t.is_contiguous():
        raise _lib.MauvError("mauv ops take contiguous tensors")
    if dtype is not None and t.dtype != dtype:
        raise _lib.MauvError(f"expected {dtype}, got {t.dtype}")
    return t.data_ptr()


_stream_override = None


def _stream() -> int:
    if _stream_override is not None:
        return _stream_override
    return torch.cuda.current_stream().cuda_stream


class on_current_stream:
    """Resolve torch's current stream ONCE for a whole engine walk (torch.cuda.current_stream() costs ~10 us per call,
    a walk makes ~2 000 calls). Do not switch streams inside the block."""

    def __enter__(self):
        global _stream_override
        self._prev = _stream_override
        _stream_override = torch.cuda.current_stream().cuda_stream
        return self

    def __exit__(self, *exc):
        global _stream_override
        _stream_override = self._prev
        return False



class sample_base:
    """While active, the forward sampling kernels launched by this thread add the uint32 word `word` (a 1-element int32 /
    uint32 CUDA tensor) to their Philox sample ids when they run (mauv_set_sample_base): a CUDA graph captured inside the
    block draws fresh eps on every replay after `word` is bumped."""

    def __init__(self, word: Optional[torch.Tensor]):
        if word is not None and (not word.is_cuda or word.numel() != 1 or word.element_size() != 4):
            raise _lib.MauvError("sample_base: a 1-element 32-bit CUDA tensor is required")
        self.word = word

    def __enter__(self):
        _lib.check(_lib.require_device().mauv_set_sample_base(self.word.data_ptr() if self.word is not None else None))
        return self

    def __exit__(self, *exc):
        _lib.check(_lib.require_device().mauv_set_sample_base(None))
        return False


# ------------------------------------------------------------------ launch accounting / profiling
# kernels launched per C-ABI call (bench.py reports the sum as gpu_launches)
KERNELS_PER_CALL = {"mauv_kl_fwd_bwd": 2, "mauv_bn_finalize": 2, "mauv_sampled_linear_bwd_f32": 2}
launch_count = 0
_prof = None   # list of (name, start_event, end_event) while profiling


def start_profile():
    global _prof
    _prof = []


def stop_profile():
    """-> {name: (calls, total_ms)} measured with CUDA events on the launching stream."""
    global _prof
    rec, _prof = _prof, None
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1 in rec or []:
        c, t = out.get(name, (0, 0.0))
        out[name] = (c + 1, t + e0.elapsed_time(e1))
    return out


def _run(name, fn, *args, tag=None):
    global launch_count
    launch_count += KERNELS_PER_CALL.get(name, 1)
    if _prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        _prof.append((name if tag is None else f"{name}|{tag}", e0, e1))
    else:
        rc = fn(*args)
    _lib.check(rc)


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# ------------------------------------------------------------------ sampling
def sample_weights_f16(mu: torch.Tensor, rho: torch.Tensor, G: int, *, eps: Optional[torch.Tensor] = None,
                       seed: int = 0, layer_id: int = 0, sample0: int = 0,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """mu/rho: [cout, cin, kh, kw] or [out, in] fp32 -> [G, cout, k_pad] fp16, K order (kh, kw, cin)."""
    lib = _lib.require_device()
    if mu.dim() == 2:
        cout, cin = mu.shape
        kh = kw = 1
    else:
        cout, cin, kh, kw = mu.shape
    K = cin * kh * kw
    k_pad = round_up(K, 8)
    if out is None:
        out = torch.empty((G, cout, k_pad), dtype=F16, device=mu.device)
    if eps is not None:
        assert eps.numel() == G * mu.numel(), "eps must be [G, *mu.shape]"
    _run("mauv_sample_weights_f16", lib.mauv_sample_weights_f16, _ptr(mu, F32), _ptr(rho, F32), _ptr(eps, F32), seed, layer_id, sample0,
                                           G, cout, cin, kh, kw, k_pad, _ptr(out, F16), _stream())
    return out


def sample_vector_f32(mu: torch.Tensor, rho: torch.Tensor, G: int, *, eps: Optional[torch.Tensor] = None,
                      seed: int = 0, layer_id: int = 0, sample0: int = 0) -> torch.Tensor:
    lib = _lib.require_device()
    out = torch.empty((G, mu.numel()), dtype=F32, device=mu.device)
    _run("mauv_sample_vector_f32", lib.mauv_sample_vector_f32, _ptr(mu, F32), _ptr(rho, F32), _ptr(eps, F32), seed, layer_id, sample0,
                                          G, mu.numel(), _ptr(out), _stream())
    return out


def philox_normal(n: int, *, seed: int, layer_id: int, sample_id: int, device="cuda") -> torch.Tensor:
    lib = _lib.require_device()
    out = torch.empty(n, dtype=F32, device=device)
    _run("mauv_philox_normal_f32", lib.mauv_philox_normal_f32, seed, layer_id, sample_id, n, _ptr(out), _stream())
    return out


# ------------------------------------------------------------------ contraction
def gemm_m_tiles(M: int) -> int:
    return (M + 127) // 128


def gemm_f16(a: torch.Tensor, w: torch.Tensor, *, bias: Optional[torch.Tensor] = None, stats: bool = False,
             shared_a: bool = False, out: Optional[torch.Tensor] = None,
             stats_out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """a: [G, M, K] (or [M, K] with shared_a) fp16, w: [G, N, K] fp16 -> y [G, M, N] fp16,
    stats partial [G, m_tiles, N, 2] fp32."""
    lib = _lib.require_device()
    G, N, K = w.shape
    if shared_a:
        M = a.shape[-2]
        stride = 0
    else:
        assert a.shape[0] == G
        M = a.shape[1]
        stride = M * K
    assert a.shape[-1] == K
    if out is None:
        out = torch.empty((G, M, N), dtype=F16, device=a.device)
    if stats and stats_out is None:
        stats_out = torch.empty((G, gemm_m_tiles(M), N, 2), dtype=F32, device=a.device)
    _run("mauv_gemm_f16", lib.mauv_gemm_f16, _ptr(a, F16), stride, _ptr(w, F16), _ptr(bias, F32), _ptr(out, F16),
                                 _ptr(stats_out, F32) if stats else None, G, M, N, K, _stream(),
         tag=f"G{G} M{M} N{N} K{K}" if _prof is not None else None)
    return out, (stats_out if stats else None)


def gemm_wmod_f16(a: torch.Tensor, w: torch.Tensor, out_f32: bool = False) -> torch.Tensor:
    """a [G, M, K], w [Gw, N, K] with G % Gw == 0 -> y [G, M, N] = a[g] @ w[g % Gw]^T (fp16, or the fp32 accumulator)"""
    lib = _lib.require_device()
    G, M, K = a.shape
    Gw, N, _ = w.shape
    assert w.shape[2] == K and G % Gw == 0
    out = torch.empty((G, M, N), dtype=F32 if out_f32 else F16, device=a.device)
    _run("mauv_gemm_wmod_f16", lib.mauv_gemm_wmod_f16, _ptr(a, F16), _ptr(w, F16), Gw, _ptr(out), int(out_f32), G, M, N, K, _stream(),
         tag=f"G{G} M{M} N{N} K{K} wmod{Gw}" if _prof is not None else None)
    return out


def gemm_stats_f16(a: torch.Tensor, w: torch.Tensor, stats_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """First pass of the recompute scheme: BN partial statistics of A*W^T without storing it. a [G,M,K], w [G,N,K]."""
    lib = _lib.require_device()
    G, N, K = w.shape
    M = a.shape[1]
    if stats_out is None:
        stats_out = torch.empty((G, (M + 255) // 256, N, 2), dtype=F32, device=a.device)
    _run("mauv_gemm_bn_f16", lib.mauv_gemm_bn_f16, _ptr(a, F16), _ptr(w, F16), None, _ptr(stats_out, F32), None, None, 0, 1,
         G, M, N, K, _stream(), tag=f"stats G{G} M{M} N{N} K{K}" if _prof is not None else None)
    return stats_out


def gemm_bn_act_f16(a: torch.Tensor, w: torch.Tensor, ss: torch.Tensor, *, residual: Optional[torch.Tensor] = None,
                    relu: bool = True, out: Optional[torch.Tensor] = None, a_ss: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Second pass: out = relu?((A W^T) * scale + shift [+ residual]) -> [G, M, N] fp16. a_ss [G, K, 2]: `a` is the RAW output
    of the previous conv and A = relu(a * a_ss.scale + a_ss.shift) is formed on the operand tiles in shared memory."""
    lib = _lib.require_device()
    G, N, K = w.shape
    M = a.shape[1]
    if out is None:
        out = torch.empty((G, M, N), dtype=F16, device=a.device)
    if a_ss is not None:
        _run("mauv_gemm_bn_xf_f16", lib.mauv_gemm_bn_xf_f16, _ptr(a, F16), _ptr(a_ss, F32), _ptr(w, F16), _ptr(out, F16), _ptr(ss, F32),
             _ptr(residual, F16), int(relu), G, M, N, K, _stream(),
             tag=f"fused G{G} M{M} N{N} K{K} res{int(residual is not None)} xf" if _prof is not None else None)
        return out
    _run("mauv_gemm_bn_f16", lib.mauv_gemm_bn_f16, _ptr(a, F16), _ptr(w, F16), _ptr(out, F16), None, _ptr(ss, F32),
         _ptr(residual, F16), int(relu), 2, G, M, N, K, _stream(),
         tag=f"fused G{G} M{M} N{N} K{K} res{int(residual is not None)}" if _prof is not None else None)
    return out


def conv2d_im2col_f16(x: torch.Tensor, w: torch.Tensor, G: int, kh: int, kw: int, stride: int, pad: int, *,
                      stats: bool = False, out: Optional[torch.Tensor] = None,
                      stats_out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """x: [G*B, H, W, Cin] NHWC fp16, w: [G, Cout, kh*kw*Cin] fp16 -> y [G*B, Ho, Wo, Cout] fp16."""
    lib = _lib.require_device()
    NB, H, W, Cin = x.shape
    assert NB % G == 0
    B = NB // G
    Cout = w.shape[1]
    assert w.shape[0] == G and w.shape[2] == kh * kw * Cin
    if (STREAM_CONV and Cin == 64 and Cout == 64 and kh == 3 and kw == 3 and stride == 1 and pad == 1 and W <= 254
            and out is None and stats_out is None):
        return conv3x3_c64_f16(x, w, G, stats=stats)          # ResNet layer1 conv2: padded-stream kernel
    Ho = (H + 2 * pad - kh) // stride + 1
    Wo = (W + 2 * pad - kw) // stride + 1
    if out is None:
        out = torch.empty((NB, Ho, Wo, Cout), dtype=F16, device=x.device)
    if stats and stats_out is None:
        stats_out = torch.empty((G, gemm_m_tiles(B * Ho * Wo), Cout, 2), dtype=F32, device=x.device)
    _run("mauv_conv2d_im2col_f16", lib.mauv_conv2d_im2col_f16, _ptr(x, F16), _ptr(w, F16), _ptr(out, F16),
                                          _ptr(stats_out, F32) if stats else None, G, B, H, W, Cin, Cout,
                                          kh, kw, stride, pad, _stream(),
         tag=f"G{G} M{B * Ho * Wo} N{Cout} K{kh * kw * Cin} {kh}x{kw}/{stride}" if _prof is not None else None)
    return out, (stats_out if stats else None)


def sample_weights_scaled_f16(mu, rho, G, ss, out, col0, *, eps=None, seed=0, layer_id=0, sample0=0) -> None:
    """1x1 weights [cout, cin(,1,1)] sampled into columns [col0, col0+cin) of out [G, cout, Ktot], rows scaled by ss[g, co, 0]"""
    lib = _lib.require_device()
    cout, cin = mu.shape[0], mu.shape[1]
    assert mu.numel() == cout * cin and out.shape[0] == G and out.shape[1] == cout
    _run("mauv_sample_weights_scaled_f16", lib.mauv_sample_weights_scaled_f16, _ptr(mu, F32), _ptr(rho, F32), _ptr(eps, F32), seed,
         layer_id, sample0, G, cout, cin, _ptr(ss, F32), out.shape[2], col0, _ptr(out, F16), _stream())


def bn_shift_sum(ss_a: torch.Tensor, ss_b: torch.Tensor) -> torch.Tensor:
    lib = _lib.require_device()
    out = torch.empty_like(ss_a)
    _run("mauv_bn_shift_sum", lib.mauv_bn_shift_sum, _ptr(ss_a, F32), _ptr(ss_b, F32), ss_a.numel() // 2, _ptr(out), _stream())
    return out


def subsample_f16(x: torch.Tensor, stride: int) -> torch.Tensor:
    lib = _lib.require_device()
    N, H, W, Cc = x.shape
    out = torch.empty((N, (H - 1) // stride + 1, (W - 1) // stride + 1, Cc), dtype=F16, device=x.device)
    _run("mauv_subsample_f16", lib.mauv_subsample_f16, _ptr(x, F16), N, H, W, Cc, stride, _ptr(out), _stream())
    return out


def gemm_bn_cat_f16(a1: torch.Tensor, a2: torch.Tensor, w_cat: torch.Tensor, shift: torch.Tensor, *, relu: bool = True,
                    a1_ss: Optional[torch.Tensor] = None) -> torch.Tensor:
    """relu?([a1 | a2] * w_cat^T + shift): a1 [G, M, K1], a2 [G, M, K2], w_cat [G, N, K1+K2], shift [G, N, 2] -> [G, M, N].
    a1_ss [G, K1, 2]: a1 is raw and relu(a1 * scale + shift) is formed on its operand tiles in shared memory."""
    lib = _lib.require_device()
    G, M, K1 = a1.shape
    K2 = a2.shape[2]
    N = w_cat.shape[1]
    assert w_cat.shape[2] == K1 + K2 and a2.shape[:2] == a1.shape[:2]
    out = torch.empty((G, M, N), dtype=F16, device=a1.device)
    if a1_ss is not None:
        _run("mauv_gemm_bn_cat_xf_f16", lib.mauv_gemm_bn_cat_xf_f16, _ptr(a1, F16), _ptr(a1_ss, F32), K1, _ptr(a2, F16), K2,
             _ptr(w_cat, F16), _ptr(out), _ptr(shift, F32), int(relu), G, M, N, _stream(),
             tag=f"fused G{G} M{M} N{N} K{K1 + K2} cat xf" if _prof is not None else None)
        return out
    _run("mauv_gemm_bn_cat_f16", lib.mauv_gemm_bn_cat_f16, _ptr(a1, F16), K1, _ptr(a2, F16), K2, _ptr(w_cat, F16), _ptr(out),
         _ptr(shift, F32), int(relu), G, M, N, _stream(),
         tag=f"fused G{G} M{M} N{N} K{K1 + K2} cat" if _prof is not None else None)
    return out


# C-ABI entry points whose kernel is a tcgen05 contraction (bench.py's roofline family: every launch of these is counted)
TCGEN05_ENTRY_POINTS = ("mauv_gemm_f16", "mauv_conv2d_im2col_f16", "mauv_gemm_bn_f16", "mauv_conv3x3_c64_f16", "mauv_gemm_bn_cat_f16",
                        "mauv_wgrad_f16", "mauv_stem_conv_pool_f16", "mauv_gemm_bn_xf_f16", "mauv_gemm_bn_cat_xf_f16",
                        "mauv_gram_bn_f16", "mauv_gemm_wmod_f16", "mauv_gemm_x3_f16", "mauv_conv2d_im2col_x3_f16")

STREAM_CONV = __import__("os").environ.get("MAUV_STREAM_CONV", "1") != "0"


def conv3x3_c64_f16(x: torch.Tensor, w: torch.Tensor, G: int, *, stats: bool = False, in_ss: Optional[torch.Tensor] = None):
    """3x3 / stride 1 / pad 1, 64 -> 64 channels (padded-stream kernel). x [G*B, H, W, 64], w [G, 64, 576].
    in_ss [G, 64, 2]: x is a raw conv output; relu(x * scale + shift) is applied to the tiles in shared memory."""
    lib = _lib.require_device()
    NB, H, W, Cin = x.shape
    assert Cin == 64 and tuple(w.shape) == (G, 64, 576) and NB % G == 0
    B = NB // G
    out = torch.empty((NB, H, W, 64), dtype=F16, device=x.device)
    st = torch.empty((G, lib.mauv_conv3x3_c64_tiles(B, H, W), 64, 2), dtype=F32, device=x.device) if stats else None
    _run("mauv_conv3x3_c64_f16", lib.mauv_conv3x3_c64_f16, _ptr(x, F16), _ptr(w, F16), _ptr(out), _ptr(st), _ptr(in_ss, F32), G, B, H,
         W, _stream(), tag=f"G{G} M{B * H * W} N64 K576 stream{'+bn' if in_ss is not None else ''}" if _prof is not None else None)
    return out, st


# ------------------------------------------------------------------ BN / pooling
def stem_im2col_f16(x_nchw: torch.Tensor, kh: int, kw: int, stride: int, pad: int,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.require_device()
    B, C, H, W = x_nchw.shape
    k_pad = round_up(kh * kw * C, 8)
    Ho = (H + 2 * pad - kh) // stride + 1
    Wo = (W + 2 * pad - kw) // stride + 1
    if out is None:
        out = torch.empty((B * Ho * Wo, k_pad), dtype=F16, device=x_nchw.device)
    _run("mauv_stem_im2col_f16", lib.mauv_stem_im2col_f16, _ptr(x_nchw, F32), B, C, H, W, kh, kw, stride, pad, k_pad,
                                        _ptr(out, F16), _stream())
    return out


def bn_finalize(stats_partial: torch.Tensor, count: int, gamma: Optional[torch.Tensor],
                beta: Optional[torch.Tensor], eps: float = 1e-5, momentum: float = 0.1,
                running_mean: Optional[torch.Tensor] = None, running_var: Optional[torch.Tensor] = None,
                want_batch_stats: bool = False, out: Optional[torch.Tensor] = None,
                num_batches_tracked: Optional[torch.Tensor] = None):
    lib = _lib.require_device()
    G, m_tiles, Cc, _ = stats_partial.shape
    if out is None:
        out = torch.empty((G, Cc, 2), dtype=F32, device=stats_partial.device)
    bs = torch.empty((G, Cc, 2), dtype=F32, device=stats_partial.device) if want_batch_stats else None
    ws_bytes = lib.mauv_bn_finalize_ws_bytes(G, m_tiles, Cc)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=stats_partial.device)
    _run("mauv_bn_finalize", lib.mauv_bn_finalize, _ptr(stats_partial, F32), G, m_tiles, Cc, count, _ptr(gamma, F32),
                                    _ptr(beta, F32), eps, momentum, _ptr(running_mean, F32),
                                    _ptr(running_var, F32), _ptr(num_batches_tracked, I64), _ptr(out), _ptr(bs), _ptr(ws), _stream())
    return (out, bs) if want_batch_stats else out


def bn_act_f16(y: torch.Tensor, ss: torch.Tensor, G: int, C: int, *, residual: Optional[torch.Tensor] = None,
               y2: Optional[torch.Tensor] = None, ss2: Optional[torch.Tensor] = None, relu: bool = True,
               out: Optional[torch.Tensor] = None, colsum: bool = False):
    """out = relu?(y*ss [+ residual] [+ y2*ss2]). colsum=True -> (out, per-block column sums [G, nblk, C] fp32 of out)."""
    lib = _lib.require_device()
    M = y.numel() // (G * C)
    if out is None:
        out = torch.empty_like(y)
    cs = torch.empty((G, lib.mauv_bn_act_blocks(G, M, C), C), dtype=F32, device=y.device) if colsum else None
    _run("mauv_bn_act_f16", lib.mauv_bn_act_f16, _ptr(y, F16), _ptr(ss, F32), _ptr(residual, F16), _ptr(y2, F16), _ptr(ss2, F32),
                                   int(relu), G, M, C, _ptr(out, F16), _ptr(cs), _stream(),
         tag=f"G{G} M{M} C{C} res{int(residual is not None)} dual{int(y2 is not None)}" if _prof is not None else None)
    return (out, cs) if colsum else out


def colsum_f16(x: torch.Tensor, G: int, C: int) -> torch.Tensor:
    """per-block column sums [G, nblk, C] fp32 of x [G, M, C] fp16"""
    lib = _lib.require_device()
    M = x.numel() // (G * C)
    cs = torch.empty((G, lib.mauv_bn_act_blocks(G, M, C), C), dtype=F32, device=x.device)
    _run("mauv_colsum_f16", lib.mauv_colsum_f16, _ptr(x, F16), G, M, C, _ptr(cs), _stream())
    return cs


KERNELS_PER_CALL["mauv_bn_stats_from_gram"] = 3


def gram_splits(M: int, G: int, K: int) -> int:
    """Pixel chunks per sample for the second-moment contraction a^T a (0 = shape not eligible). Chunks are a multiple of 64
    pixels (the weight-gradient kernel's k-block) and at most 4096 pixels long: the tensor core accumulates in fp32 with
    truncation, and a sum of same-sign products (the diagonal of a^T a) picks up a relative bias of ~n_adds * 2^-25, so
    the long sums are split and the partials are combined in double precision (4096 px = 256 MMA accumulations: < 1e-5);
    shorter chunks (down to 2048) only when that is needed to fill the GPU."""
    if K % 64 != 0 or K > 256 or M % 64 != 0:
        return 0
    tiles = G * ((K + 127) // 128) * ((K + 255) // 256)
    s = 1
    while M % (2 * s * 64) == 0 and (M // (2 * s) >= 4096 or (tiles * s < 2 * 148 and M // (2 * s) >= 2048)):
        s *= 2
    return s


def bn_stats_from_gram(a: torch.Tensor, colsum: Optional[torch.Tensor], w: torch.Tensor, count: int, gamma, beta, eps: float,
                       momentum: float, running_mean=None, running_var=None, num_batches_tracked=None,
                       want_batch_stats: bool = False, splits: Optional[int] = None, a_ss: Optional[torch.Tensor] = None):
    """BatchNorm(train) scale/shift [G, N, 2] of y = a w^T without computing y: a [G, M, K] fp16 (contiguous), colsum
    [G, nblk, K] (bn_act_f16(colsum=True) / colsum_f16), w [G, N, K] fp16. With a_ss [G, K, 2] `a` is the RAW output of the
    previous conv: its BatchNorm + ReLU is applied inside the second-moment kernel, which then also produces the column sums."""
    lib = _lib.require_device()
    G, M, K = a.shape
    N = w.shape[1]
    assert w.shape[0] == G and w.shape[2] == K
    splits = splits or gram_splits(M, G, K)
    if splits == 0:
        raise _lib.MauvError(f"bn_stats_from_gram: shape M={M} K={K} is not eligible")
    if a_ss is not None:
        gram = torch.empty((G * splits, K, K), dtype=F32, device=a.device)
        colsum = torch.empty((G, splits, K), dtype=F32, device=a.device)
        _run("mauv_gram_bn_f16", lib.mauv_gram_bn_f16, _ptr(a, F16), _ptr(a_ss, F32), _ptr(gram), _ptr(colsum), G, splits, M, K, _stream(),
             tag=f"G{G}x{splits} Cout{K} K{K} px{M // splits} xf" if _prof is not None else None)
    else:
        assert colsum.shape[0] == G and colsum.shape[2] == K
        x4 = a.view(G, M, 1, K)                     # NHWC view [G*B', H, W, C] with B'=1, H=M, W=1
        gram = wgrad_f16(x4, x4, G, splits, 1, 1, 1, 0)            # [G*splits, K, K] fp32
    out = torch.empty((G, N, 2), dtype=F32, device=a.device)
    bs = torch.empty((G, N, 2), dtype=F32, device=a.device) if want_batch_stats else None
    ws = torch.empty(lib.mauv_bn_stats_from_gram_ws_bytes(G, N, K), dtype=torch.uint8, device=a.device)
    _run("mauv_bn_stats_from_gram", lib.mauv_bn_stats_from_gram, _ptr(gram, F32), splits, _ptr(colsum, F32), colsum.shape[1],
         _ptr(w, F16), G, N, K, count, _ptr(gamma, F32), _ptr(beta, F32), eps, momentum, _ptr(running_mean, F32),
         _ptr(running_var, F32), _ptr(num_batches_tracked, I64), _ptr(out), _ptr(bs), _ptr(ws), _stream())
    return (out, bs) if want_batch_stats else out


def bn_relu_maxpool_f16(y: torch.Tensor, ss: torch.Tensor, G: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.require_device()
    NB, H, W, Cc = y.shape
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    if out is None:
        out = torch.empty((NB, Ho, Wo, Cc), dtype=F16, device=y.device)
    _run("mauv_bn_relu_maxpool_f16", lib.mauv_bn_relu_maxpool_f16, _ptr(y, F16), _ptr(ss, F32), G, NB // G, H, W, Cc, _ptr(out, F16), _stream())
    return out


def stem_conv_pool_f16(a0: torch.Tensor, w: torch.Tensor, imgs: int, Ho: int, gamma: Optional[torch.Tensor] = None):
    """Inference stem (256 x 256 inputs: Wo = 128): a0 [imgs*Ho*128, Kp] im2col matrix shared by all samples, w [G, 64, Kp]
    -> (3x3/2 max-pool of the RAW conv output [G*imgs, Ho/2, 64, 64] fp16 - window min where gamma < 0 -,
        BN partial statistics [G, imgs*Ho, 64, 2] of the full-resolution output)."""
    lib = _lib.require_device()
    G, N, Kp = w.shape
    assert N == 64 and a0.shape == (imgs * Ho * 128, Kp) and Ho % 2 == 0
    pooled = torch.empty((G * imgs, Ho // 2, 64, 64), dtype=F16, device=a0.device)
    stats = torch.empty((G, imgs * Ho, 64, 2), dtype=F32, device=a0.device)
    _run("mauv_stem_conv_pool_f16", lib.mauv_stem_conv_pool_f16, _ptr(a0, F16), _ptr(w, F16), _ptr(pooled, F16), _ptr(stats, F32),
         _ptr(gamma, F32), G, imgs, Ho, Kp, _stream(), tag=f"G{G} M{imgs * Ho * 128} N64 K{Kp} pool" if _prof is not None else None)
    return pooled, stats


def avgpool_f16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: [N, H, W, C] fp16 -> [N, C] fp32"""
    lib = _lib.require_device()
    N, H, W, Cc = x.shape
    if out is None:
        out = torch.empty((N, Cc), dtype=F32, device=x.device)
    _run("mauv_avgpool_f16", lib.mauv_avgpool_f16, _ptr(x, F16), N, H * W, Cc, _ptr(out, F32), _stream())
    return out


def nchw_f32_to_nhwc_f16(x: torch.Tensor, c_pad: Optional[int] = None) -> torch.Tensor:
    lib = _lib.require_device()
    N, Cc, H, W = x.shape
    c_pad = c_pad or Cc
    out = torch.empty((N, H, W, c_pad), dtype=F16, device=x.device)
    _run("mauv_nchw_f32_to_nhwc_f16", lib.mauv_nchw_f32_to_nhwc_f16, _ptr(x, F32), N, Cc, H * W, c_pad, _ptr(out), _stream())
    return out


def nhwc_f16_to_nchw_f32(x: torch.Tensor) -> torch.Tensor:
    lib = _lib.require_device()
    N, H, W, Cc = x.shape
    out = torch.empty((N, Cc, H, W), dtype=F32, device=x.device)
    _run("mauv_nhwc_f16_to_nchw_f32", lib.mauv_nhwc_f16_to_nchw_f32, _ptr(x, F16), N, Cc, H * W, _ptr(out), _stream())
    return out


# ------------------------------------------------------------------ head
def sampled_linear_f32(x: torch.Tensor, mu_w, rho_w, mu_b, rho_b, *, eps_w=None, eps_b=None, seed: int = 0,
                       layer_id: int = 0, sample0: int = 0, out: Optional[torch.Tensor] = None,
                       out_col: int = 0) -> torch.Tensor:
    """x: [G, B, in] fp32 (last-dim stride 1; may be a column slice of a wider buffer) -> y [G, B, out]."""
    lib = _lib.require_device()
    G, B, fin = x.shape
    fout = mu_w.shape[0]
    assert x.stride(2) == 1
    if out is None:
        out = torch.empty((G, B, fout), dtype=F32, device=x.device)
        out_col = 0
    y_view = out[:, :, out_col:out_col + fout]
    _run("mauv_sampled_linear_f32", lib.mauv_sampled_linear_f32, 
        x.data_ptr(), x.stride(0), x.stride(1), _ptr(mu_w, F32), _ptr(rho_w, F32), _ptr(eps_w, F32),
        _ptr(mu_b, F32), _ptr(rho_b, F32), _ptr(eps_b, F32), seed, layer_id, sample0, G, B, fin, fout,
        y_view.data_ptr(), out.stride(0), out.stride(1), _stream())
    return out


def tanh_add_f32(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    lib = _lib.require_device()
    out = torch.empty_like(a)
    _run("mauv_tanh_add_f32", lib.mauv_tanh_add_f32, _ptr(a, F32), _ptr(b, F32), a.numel(), _ptr(out), _stream())
    return out


def softmax_gate_f32(score: torch.Tensor, v: torch.Tensor, out: torch.Tensor, out_col: int = 0) -> torch.Tensor:
    """out[..., out_col:out_col+n] = v * softmax(score, -1); out is [rows..., ld]."""
    lib = _lib.require_device()
    n = score.shape[-1]
    rows = score.numel() // n
    assert out.is_contiguous() and out.dtype == F32
    ld = out.shape[-1]
    _run("mauv_softmax_gate_f32", lib.mauv_softmax_gate_f32, _ptr(score, F32), _ptr(v, F32), rows, n,
                                         out.data_ptr() + 4 * out_col, ld, _stream())
    return out


# ------------------------------------------------------------------ fp16x3 validation mode ((hi | lo) pairs)
X3_SCALE = 256.0   # weights are scaled by 2^8 so that their lo parts stay clear of fp16 subnormals; BN absorbs it


def sample_weights_x3_f16(mu, rho, G, *, per_tap: bool, eps=None, seed=0, layer_id=0, sample0=0) -> torch.Tensor:
    lib = _lib.require_device()
    if mu.dim() == 2:
        cout, cin = mu.shape
        kh = kw = 1
    else:
        cout, cin, kh, kw = mu.shape
    kp = round_up(cin * kh * kw, 64)
    row = kh * kw * 3 * cin if per_tap else 3 * kp
    out = torch.empty((G, cout, row), dtype=F16, device=mu.device)
    _run("mauv_sample_weights_x3_f16", lib.mauv_sample_weights_x3_f16, _ptr(mu, F32), _ptr(rho, F32), _ptr(eps, F32), seed,
         layer_id, sample0, G, cout, cin, kh, kw, kp, int(per_tap), X3_SCALE, _ptr(out), _stream())
    return out


def stem_im2col_x3_f16(x_nchw: torch.Tensor, kh, kw, stride, pad) -> torch.Tensor:
    lib = _lib.require_device()
    B, C, H, W = x_nchw.shape
    kp = round_up(kh * kw * C, 64)
    Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    out = torch.empty((B * Ho * Wo, 2 * kp), dtype=F16, device=x_nchw.device)
    _run("mauv_stem_im2col_x3_f16", lib.mauv_stem_im2col_x3_f16, _ptr(x_nchw, F32), B, C, H, W, kh, kw, stride, pad, kp,
         _ptr(out), _stream())
    return out


def gemm_x3_f16(a2: torch.Tensor, w3: torch.Tensor, *, shared_a: bool = False):
    """a2 [G, M, 2K] (or [M, 2K] shared), w3 [G, N, 3K] -> y2 [G, M, 2N] (hi | lo), stats [G, m_tiles, N, 2]."""
    lib = _lib.require_device()
    G, N, K3 = w3.shape
    K = K3 // 3
    M = a2.shape[-2]
    assert a2.shape[-1] == 2 * K
    y2 = torch.empty((G, M, 2 * N), dtype=F16, device=a2.device)
    st = torch.empty((G, gemm_m_tiles(M), N, 2), dtype=F32, device=a2.device)
    _run("mauv_gemm_x3_f16", lib.mauv_gemm_x3_f16, _ptr(a2, F16), 0 if shared_a else M * 2 * K, _ptr(w3, F16), _ptr(y2),
         _ptr(st), G, M, N, K, _stream())
    return y2, st


def conv2d_im2col_x3_f16(x2: torch.Tensor, w3: torch.Tensor, G: int, kh, kw, stride, pad):
    """x2 [G*B, H, W, 2Cin], w3 [G, Cout, kh*kw*3*Cin] -> y2 [G*B, Ho, Wo, 2Cout], stats."""
    lib = _lib.require_device()
    NB, H, W, C2 = x2.shape
    Cin, B, Cout = C2 // 2, NB // G, w3.shape[1]
    Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    y2 = torch.empty((NB, Ho, Wo, 2 * Cout), dtype=F16, device=x2.device)
    st = torch.empty((G, gemm_m_tiles(B * Ho * Wo), Cout, 2), dtype=F32, device=x2.device)
    _run("mauv_conv2d_im2col_x3_f16", lib.mauv_conv2d_im2col_x3_f16, _ptr(x2, F16), _ptr(w3, F16), _ptr(y2), _ptr(st), G, B,
         H, W, Cin, Cout, kh, kw, stride, pad, _stream())
    return y2, st


def bn_act_x3_f16(y2, ss, G, C, *, residual=None, y2b=None, ss2=None, relu=True) -> torch.Tensor:
    lib = _lib.require_device()
    M = y2.numel() // (G * 2 * C)
    out = torch.empty_like(y2)
    _run("mauv_bn_act_x3_f16", lib.mauv_bn_act_x3_f16, _ptr(y2, F16), _ptr(ss, F32), _ptr(residual, F16), _ptr(y2b, F16),
         _ptr(ss2, F32), int(relu), G, M, C, _ptr(out), _stream())
    return out


def bn_relu_maxpool_x3_f16(y2, ss, G) -> torch.Tensor:
    lib = _lib.require_device()
    NB, H, W, C2 = y2.shape
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    out = torch.empty((NB, Ho, Wo, C2), dtype=F16, device=y2.device)
    _run("mauv_bn_relu_maxpool_x3_f16", lib.mauv_bn_relu_maxpool_x3_f16, _ptr(y2, F16), _ptr(ss, F32), G, NB // G, H, W, C2 // 2,
         _ptr(out), _stream())
    return out


def avgpool_x3_f16(x2) -> torch.Tensor:
    lib = _lib.require_device()
    N, H, W, C2 = x2.shape
    out = torch.empty((N, C2 // 2), dtype=F32, device=x2.device)
    _run("mauv_avgpool_x3_f16", lib.mauv_avgpool_x3_f16, _ptr(x2, F16), N, H * W, C2 // 2, _ptr(out), _stream())
    return out


# ------------------------------------------------------------------ backward helpers
def sample_weights_dgrad_f16(mu, rho, G, *, eps=None, seed=0, layer_id=0, sample0=0) -> torch.Tensor:
    """-> [G, cin, kh*kw*cout] fp16: transposed + tap-flipped sample for the data-gradient conv."""
    lib = _lib.require_device()
    if mu.dim() == 2:
        cout, cin = mu.shape
        kh = kw = 1
    else:
        cout, cin, kh, kw = mu.shape
    out = torch.empty((G, cin, kh * kw * cout), dtype=F16, device=mu.device)
    _run("mauv_sample_weights_dgrad_f16", lib.mauv_sample_weights_dgrad_f16, _ptr(mu, F32), _ptr(rho, F32),
         _ptr(eps, F32), seed, layer_id, sample0, G, cout, cin, kh, kw, _ptr(out), _stream())
    return out


def weights_to_dgrad_f16(w: torch.Tensor, cin: int, kh: int, kw: int) -> torch.Tensor:
    """forward sample [G, cout, kh*kw*cin] -> [G, cin, kh*kw*cout] (taps flipped) for the data-gradient conv"""
    lib = _lib.require_device()
    G, cout, K = w.shape
    assert K == kh * kw * cin
    out = torch.empty((G, cin, kh * kw * cout), dtype=F16, device=w.device)
    _run("mauv_weights_to_dgrad_f16", lib.mauv_weights_to_dgrad_f16, _ptr(w, F16), G, cout, cin, kh, kw, _ptr(out), _stream())
    return out


def dilate_f16(x: torch.Tensor, Hd: int, Wd: int, stride: int) -> torch.Tensor:
    lib = _lib.require_device()
    N, Ho, Wo, Cc = x.shape
    out = torch.empty((N, Hd, Wd, Cc), dtype=F16, device=x.device)
    _run("mauv_dilate_f16", lib.mauv_dilate_f16, _ptr(x, F16), N, Ho, Wo, Cc, Hd, Wd, stride, _ptr(out), _stream())
    return out


def transpose_chunks_f16(src: torch.Tensor, splits: int, scale: float = 1.0) -> torch.Tensor:
    """src [M, C] fp16 -> [splits, C, M/splits]"""
    lib = _lib.require_device()
    M, Cc = src.shape
    out = torch.empty((splits, Cc, M // splits), dtype=F16, device=src.device)
    _run("mauv_transpose_chunks_f16", lib.mauv_transpose_chunks_f16, _ptr(src, F16), M, Cc, splits, scale, _ptr(out), _stream())
    return out


def im2col_t_f16(x: torch.Tensor, kh: int, kw: int, stride: int, pad: int, splits: int) -> torch.Tensor:
    """x [N, H, W, Cin] fp16 -> [splits, k_pad, M/splits] transposed im2col, K order (kh, kw, cin)."""
    lib = _lib.require_device()
    N, H, W, Cin = x.shape
    Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    M = N * Ho * Wo
    k_pad = round_up(kh * kw * Cin, 8)
    out = torch.empty((splits, k_pad, M // splits), dtype=F16, device=x.device)
    _run("mauv_im2col_t_f16", lib.mauv_im2col_t_f16, _ptr(x, F16), N, H, W, Cin, kh, kw, stride, pad, k_pad, splits,
         _ptr(out), _stream())
    return out


def wgrad_finalize(dw_partial: torch.Tensor, mu_shape, inv_scale: float, rho, grad_mu, grad_rho, *, eps=None, seed=0,
                   layer_id=0, sample_id=0) -> None:
    lib = _lib.require_device()
    splits, cout, k_pad = dw_partial.shape
    if len(mu_shape) == 2:
        cin, kh, kw = mu_shape[1], 1, 1
    else:
        _, cin, kh, kw = mu_shape
    _run("mauv_wgrad_finalize", lib.mauv_wgrad_finalize, _ptr(dw_partial, F16), splits, cout, cin, kh, kw, k_pad, inv_scale,
         _ptr(rho, F32), _ptr(eps, F32), seed, layer_id, sample_id, _ptr(grad_mu, F32), _ptr(grad_rho, F32), _stream())


def sampled_linear_bwd_f32(x, gy, mu_w, rho_w, rho_b, grad_mu_w, grad_rho_w, grad_mu_b, grad_rho_b, *, eps_w=None,
                           eps_b=None, seed=0, layer_id=0, sample_id=0, need_gx=True):
    """x [B, in], gy [B, out] fp32 -> gx [B, in] (or None); parameter grads accumulate in place."""
    lib = _lib.require_device()
    B, fin = x.shape
    fout = gy.shape[1]
    gx = torch.empty((B, fin), dtype=F32, device=x.device) if need_gx else None
    _run("mauv_sampled_linear_bwd_f32", lib.mauv_sampled_linear_bwd_f32, _ptr(x, F32), _ptr(gy, F32), _ptr(mu_w, F32),
         _ptr(rho_w, F32), _ptr(eps_w, F32), _ptr(rho_b, F32), _ptr(eps_b, F32), seed, layer_id, sample_id, B, fin, fout,
         _ptr(gx), _ptr(grad_mu_w, F32), _ptr(grad_rho_w, F32), _ptr(grad_mu_b, F32), _ptr(grad_rho_b, F32), _stream())
    return gx



# ------------------------------------------------------------------ S-batched training backward
GRAD_TARGET = 16.0   # amax the fp16 gradient tensors are renormalised to at every BatchNorm site
KERNELS_PER_CALL.update({"mauv_avgpool_bwd_f16": 2, "mauv_sampled_linear_bwd_group_f32": 2, "mauv_bn_bwd_coeffs": 2, "mauv_maxpool_bwd_f16": 2})


class GradScratch:
    """Device scalars (scales, amax / kmax bits) for one backward walk, carved from one small buffer."""

    def __init__(self, device, n: int = 4096):
        self.f = torch.zeros(n, dtype=F32, device=device)
        self.u = self.f.view(torch.int32)
        self.used = 0

    def slot(self) -> int:
        """-> byte address of a fresh 4-byte slot"""
        if self.used >= self.f.numel():
            raise _lib.MauvError("GradScratch exhausted")
        self.used += 1
        return self.f.data_ptr() + 4 * (self.used - 1)

    def value(self, addr: int) -> float:
        return float(self.f[(addr - self.f.data_ptr()) // 4])


def bn_bwd_site(gs: GradScratch, d1, s1, y, batch_stats, gamma, bn_eps, grad_gamma, grad_beta, G, C, *, d2=None, s2=None,
                relu_out=None, y2=None, batch_stats2=None, gamma2=None, bn_eps2=0.0, grad_gamma2=None, grad_beta2=None,
                want_dz=False):
    """Train-mode BatchNorm backward at one site (three launches + one per extra BN).
    -> (dy, s_dy, dy2 | None, s_dy2 | None, dz | None); scales are device addresses."""
    lib = _lib.require_device()
    M = y.numel() // (G * C)
    dev = y.device
    nblk = lib.mauv_bn_bwd_blocks(G, M, C)
    partial = torch.empty((G, nblk, 3, C), dtype=F32, device=dev)
    amax = gs.slot()
    st = _stream()
    _run("mauv_bn_bwd_reduce", lib.mauv_bn_bwd_reduce, _ptr(d1, F16), _ptr(d2, F16), s1, s2, _ptr(relu_out, F16), _ptr(y, F16),
         _ptr(y2, F16), G, M, C, _ptr(partial), amax, st,
         tag=f"G{G} M{M} C{C} d2{int(d2 is not None)} y2{int(y2 is not None)} nblk{nblk}" if _prof is not None else None)
    coef = torch.empty((G, C, 4), dtype=F32, device=dev)
    ws = torch.empty((G, C, 2), dtype=torch.float64, device=dev)
    kmax = gs.slot()
    _run("mauv_bn_bwd_coeffs", lib.mauv_bn_bwd_coeffs, _ptr(partial), G, M, C, 1, _ptr(batch_stats, F32), _ptr(gamma, F32), bn_eps,
         s1, _ptr(grad_gamma, F32), _ptr(grad_beta, F32), _ptr(coef), kmax, _ptr(ws), st)
    coef2 = kmax2 = dy2 = s_out2 = None
    if y2 is not None:
        coef2 = torch.empty((G, C, 4), dtype=F32, device=dev)
        kmax2 = gs.slot()
        _run("mauv_bn_bwd_coeffs", lib.mauv_bn_bwd_coeffs, _ptr(partial), G, M, C, 2, _ptr(batch_stats2, F32), _ptr(gamma2, F32),
             bn_eps2, s1, _ptr(grad_gamma2, F32), _ptr(grad_beta2, F32), _ptr(coef2), kmax2, _ptr(ws), st)
        dy2 = torch.empty_like(y2)
        s_out2 = gs.slot()
    dy = torch.empty_like(y)
    dz = torch.empty_like(y) if want_dz else None
    s_out = gs.slot()
    _run("mauv_bn_bwd_apply", lib.mauv_bn_bwd_apply, _ptr(d1, F16), _ptr(d2, F16), s1, s2, _ptr(relu_out, F16), _ptr(y, F16),
         _ptr(y2, F16), _ptr(coef), _ptr(coef2), amax, kmax, kmax2, GRAD_TARGET, G, M, C, _ptr(dy), _ptr(dy2), _ptr(dz), s_out,
         s_out2, st,
         tag=f"G{G} M{M} C{C} d2{int(d2 is not None)} y2{int(y2 is not None)} dz{int(want_dz)}" if _prof is not None else None)
    return dy, s_out, dy2, s_out2, dz


def maxpool_bwd_f16(y, ss, d1, s1, G, *, d2=None, s2=None) -> torch.Tensor:
    lib = _lib.require_device()
    NB, H, W, Cc = y.shape
    dz = torch.empty_like(y)
    idx = torch.empty(d1.shape, dtype=torch.uint8, device=y.device)
    _run("mauv_maxpool_bwd_f16", lib.mauv_maxpool_bwd_f16, _ptr(y, F16), _ptr(ss, F32), _ptr(d1, F16), _ptr(d2, F16), s1, s2, G,
         NB // G, H, W, Cc, _ptr(idx), _ptr(dz), _stream())
    return dz


def avgpool_bwd_f16(gs: GradScratch, dfeat: torch.Tensor, HW: int):
    """dfeat [N, C] fp32 -> (d [N, HW, C] fp16, scale address)"""
    lib = _lib.require_device()
    N, Cc = dfeat.shape
    out = torch.empty((N, HW, Cc), dtype=F16, device=dfeat.device)
    amax, s_out = gs.slot(), gs.slot()
    _run("mauv_avgpool_bwd_f16", lib.mauv_avgpool_bwd_f16, _ptr(dfeat, F32), N, HW, Cc, GRAD_TARGET, amax, _ptr(out), s_out, _stream())
    return out, s_out


def wgrad_f16(dy: torch.Tensor, x: torch.Tensor, G: int, splits: int, kh: int, kw: int, stride: int, pad: int) -> torch.Tensor:
    """dy [G*B, Ho, Wo, Cout], x [G*B, H, W, Cin] NHWC fp16 -> dw partials [G*splits, Cout, kh*kw*Cin] FP32"""
    lib = _lib.require_device()
    NB, H, W, Cin = x.shape
    Cout = dy.shape[-1]
    dw = torch.empty((G * splits, Cout, kh * kw * Cin), dtype=F32, device=x.device)
    _run("mauv_wgrad_f16", lib.mauv_wgrad_f16, _ptr(dy, F16), _ptr(x, F16), _ptr(dw), G, splits, NB // G, H, W, Cin, Cout,
         kh, kw, stride, pad, _stream(),
         tag=f"G{G}x{splits} Cout{Cout} K{kh * kw * Cin} px{dy.numel() // (G * splits * Cout)}" if _prof is not None else None)
    return dw


def wgrad_finalize_group(dw_partial, G, mu_shape, inv_alpha, scale_addr, rho, grad_mu, grad_rho, *, eps=None, seed=0,
                         layer_id=0, sample0=0, stale=False) -> None:
    lib = _lib.require_device()
    gsplits, cout, k_pad = dw_partial.shape
    if len(mu_shape) == 2:
        cin, kh, kw = mu_shape[1], 1, 1
    else:
        _, cin, kh, kw = mu_shape
    if dw_partial.dtype not in (F16, F32):
        raise _lib.MauvError("wgrad_finalize_group: partial sums must be fp16 or fp32")
    _run("mauv_wgrad_finalize_group", lib.mauv_wgrad_finalize_group, _ptr(dw_partial), int(dw_partial.dtype == F32), G, gsplits // G,
         cout, cin, kh, kw,
         k_pad, inv_alpha, scale_addr, _ptr(rho, F32), _ptr(eps, F32), seed, layer_id, sample0, int(stale), _ptr(grad_mu, F32),
         _ptr(grad_rho, F32), _stream(),
         tag=f"G{G} sp{gsplits // G} cout{cout} cin{cin} k{kh}" if _prof is not None else None)


def sampled_linear_bwd_group_f32(x, gy, mu_w, rho_w, rho_b, grad_mu_w, grad_rho_w, grad_mu_b, grad_rho_b, *, eps_w=None,
                                 eps_b=None, seed=0, layer_id=0, sample0=0, stale=False, gx=None, accumulate=False,
                                 need_gx=True):
    """x [G, B, in], gy [G, B, out] fp32 (last-dim stride 1) -> gx [G, B, in]; parameter grads accumulate in place."""
    lib = _lib.require_device()
    G, B, fin = x.shape
    fout = gy.shape[2]
    assert x.stride(2) == 1 and gy.stride(2) == 1
    if need_gx and gx is None:
        gx = torch.empty((G, B, fin), dtype=F32, device=x.device)
        accumulate = False
    _run("mauv_sampled_linear_bwd_group_f32", lib.mauv_sampled_linear_bwd_group_f32, x.data_ptr(), x.stride(0), x.stride(1),
         gy.data_ptr(), gy.stride(0), gy.stride(1), _ptr(mu_w, F32), _ptr(rho_w, F32), _ptr(eps_w, F32), _ptr(rho_b, F32),
         _ptr(eps_b, F32), seed, layer_id, sample0, G, B, fin, fout, int(stale),
         gx.data_ptr() if need_gx else None, gx.stride(0) if need_gx else 0, gx.stride(1) if need_gx else 0, int(accumulate),
         _ptr(grad_mu_w, F32), _ptr(grad_rho_w, F32), _ptr(grad_mu_b, F32), _ptr(grad_rho_b, F32), _stream())
    return gx if need_gx else None


def tanh_bwd_f32(t: torch.Tensor, dt: torch.Tensor) -> torch.Tensor:
    lib = _lib.require_device()
    out = torch.empty_like(t)
    _run("mauv_tanh_bwd_f32", lib.mauv_tanh_bwd_f32, _ptr(t, F32), _ptr(dt, F32), t.numel(), _ptr(out), _stream())
    return out


def softmax_gate_bwd_f32(score: torch.Tensor, v: torch.Tensor, dout: torch.Tensor):
    """dout: [..., n] view with last-dim stride 1 (row stride = dout.stride(-2)) -> (dscore, dv)"""
    lib = _lib.require_device()
    n = score.shape[-1]
    rows = score.numel() // n
    assert dout.stride(-1) == 1 and dout.stride(0) == dout.stride(1) * dout.shape[1]
    dscore, dv = torch.empty_like(score), torch.empty_like(v)
    _run("mauv_softmax_gate_bwd_f32", lib.mauv_softmax_gate_bwd_f32, _ptr(score, F32), _ptr(v, F32), dout.data_ptr(),
         dout.stride(-2), rows, n, _ptr(dscore), _ptr(dv), _stream())
    return dscore, dv


def ce_mean_fwd_bwd_f32(logits: torch.Tensor, labels: torch.Tensor):
    """logits [S, B, C] fp32, labels [B] int64 -> (loss 0-d, mean_logit [B, C], dlogits [S, B, C])"""
    lib = _lib.require_device()
    S, B, Cc = logits.shape
    mean_logit = torch.empty((B, Cc), dtype=F32, device=logits.device)
    dlogits = torch.empty_like(logits)
    loss = torch.empty((), dtype=F32, device=logits.device)
    _run("mauv_ce_mean_fwd_bwd_f32", lib.mauv_ce_mean_fwd_bwd_f32, _ptr(logits, F32), _ptr(labels, I64), S, B, Cc,
         _ptr(mean_logit), _ptr(dlogits), _ptr(loss), _stream())
    return loss, mean_logit, dlogits


# ------------------------------------------------------------------ optimizer
KERNELS_PER_CALL["mauv_adam_step_f32"] = 3


def adam_step_f32(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, lr: float, beta1: float, beta2: float,
                  eps: float, weight_decay: float, state: torch.Tensor) -> None:
    """Guarded Adam over flat fp32 buffers (see include/mauv_b200.h): p, m, v updated in place iff every g is finite."""
    lib = _lib.require_device()
    n = p.numel()
    assert g.numel() == n and m.numel() == n and v.numel() == n and state.numel() * 4 >= lib.mauv_adam_state_bytes()
    _run("mauv_adam_step_f32", lib.mauv_adam_step_f32, _ptr(p, F32), _ptr(g, F32), _ptr(m, F32), _ptr(v, F32), n, lr, beta1, beta2,
         eps, weight_decay, state.data_ptr(), _stream())


# ------------------------------------------------------------------ statistics
def mc_reduce(logits: torch.Tensor, eps_entropy: float = 1e-7) -> dict:
    """logits [S, B, C] fp32 -> dict of all MC statistics (see include/mauv_b200.h)."""
    lib = _lib.require_device()
    S, B, Cc = logits.shape
    dev = logits.device
    o = {
        "mean_prob": torch.empty((B, Cc), dtype=F32, device=dev),
        "mean_logit": torch.empty((B, Cc), dtype=F32, device=dev),
        "argmax_prob": torch.empty((B,), dtype=I64, device=dev),
        "argmax_logit": torch.empty((B,), dtype=I64, device=dev),
        "pred_entropy": torch.empty((B,), dtype=F32, device=dev),
        "aleatoric": torch.empty((B,), dtype=F32, device=dev),
        "mutual_info": torch.empty((B,), dtype=F32, device=dev),
        "var_mean": torch.empty((B,), dtype=F32, device=dev),
    }
    _run("mauv_mc_reduce", lib.mauv_mc_reduce, _ptr(logits, F32), S, B, Cc, 0, eps_entropy, _ptr(o["mean_prob"]),
                                  _ptr(o["mean_logit"]), _ptr(o["argmax_prob"]), _ptr(o["argmax_logit"]),
                                  _ptr(o["pred_entropy"]), _ptr(o["aleatoric"]), _ptr(o["mutual_info"]),
                                  _ptr(o["var_mean"]), _stream())
    return o


class KlPlan:
    """Device-side table of (mu, rho, grad_mu, grad_rho, n) records for one model."""

    def __init__(self, pairs, device, grads=None):
        lib = _lib.require_device()
        self.pairs = list(pairs)  # [(mu, rho)]
        self.grads = grads        # optional [(grad_mu, grad_rho)] accumulation buffers (else .grad)
        chunk = lib.mauv_kl_chunk_elems()
        prefix, tot = [], 0
        for mu, _ in self.pairs:
            prefix.append(tot)
            tot += (mu.numel() + chunk - 1) // chunk
        self.total_chunks = tot
        self.prefix = torch.tensor(prefix, dtype=I64, device=device)
        self.ws = torch.empty(lib.mauv_kl_ws_bytes(), dtype=torch.uint8, device=device)
        self.device = device
        self._table = None
        self._table_key = None

    def _build_table(self, with_grad: bool):
        rows, key = [], []
        for i, (mu, rho) in enumerate(self.pairs):
            if self.grads is not None:
                g_mu, g_rho = self.grads[i]
            else:
                g_mu, g_rho = mu.grad, rho.grad
            gm = _ptr(g_mu, F32) if (with_grad and g_mu is not None) else 0
            gr = _ptr(g_rho, F32) if (with_grad and g_rho is not None) else 0
            if with_grad and (gm == 0 or gr == 0):
                raise _lib.MauvError("KlPlan: grad buffers must be allocated before the fused KL backward")
            rows.append([mu.data_ptr(), rho.data_ptr(), gm, gr, mu.numel()])
            key.append((mu.data_ptr(), rho.data_ptr(), gm, gr))
        key = tuple(key)
        if key != self._table_key:
            self._table = torch.tensor(rows, dtype=I64, device=self.device)
            self._table_key = key
        return self._table

    def run(self, prior_mu: float, prior_sigma: float, grad_scale: Optional[float] = None) -> torch.Tensor:
        lib = _lib.require_device()
        table = self._build_table(grad_scale is not None)
        out = torch.empty((), dtype=F32, device=self.device)
        _run("mauv_kl_fwd_bwd", lib.mauv_kl_fwd_bwd, table.data_ptr(), self.prefix.data_ptr(), len(self.pairs), self.total_chunks,
                                       prior_mu, prior_sigma, grad_scale if grad_scale is not None else 0.0,
                                       out.data_ptr(), self.ws.data_ptr(), _stream())
        return out
