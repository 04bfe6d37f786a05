"""Backward of the sampled conv / linear layers (the autograd of bayesian-torch's forward that the reference's
`loss.backward()` runs: train/multimodal.py:138, train/unimodal.py:145):

    dX   = conv_transpose(dY, W_s)      tcgen05 conv over the flipped / transposed sample (zero-stuffed dY for stride 2)
    dW_s = X^T dY                       tcgen05 GEMM, reduction over pixels, split-K mapped on the kernel's batch axis
    dmu += dW_s ; drho += dW_s * eps_s * sigmoid(rho)          (eps replayed from Philox, or the injected eps)

Gradients travel in fp16 through the tensor cores with a per-tensor power-of-two loss scale (exactly undone),
accumulate in fp32 TMEM and land in fp32 parameter gradients.
"""
from __future__ import annotations

import math

import torch

from . import ops
from .bayesian import current_seed

F16, F32 = torch.float16, torch.float32


RECALIBRATE_EVERY = 64   # backward calls of a layer between two loss-scale calibrations


def _pow2_scale(t: torch.Tensor, target: float = 1024.0) -> float:
    """Power-of-two scale bringing max|t| near `target` (fp16 has 5 exponent bits; gradients are tiny)."""
    amax = float(t.abs().amax())
    if not math.isfinite(amax) or amax == 0.0:
        return 1.0
    return float(2.0 ** math.floor(math.log2(target / amax)))


def _layer_scale(layer, gy: torch.Tensor) -> float:
    """Per-layer loss scale, calibrated with one device->host read every RECALIBRATE_EVERY backward calls instead of
    one per call (a sync per layer per MC pass would drain the launch pipeline ~900 times per training step). The
    target leaves 32x headroom below the fp16 maximum for drift between calibrations; an overflow would surface as
    inf gradients, which the training drivers already guard against (reference train/multimodal.py:141-145)."""
    n = getattr(layer, "_gscale_age", RECALIBRATE_EVERY)
    if n >= RECALIBRATE_EVERY or not hasattr(layer, "_gscale"):
        layer._gscale = _pow2_scale(gy)
        n = 0
    layer._gscale_age = n + 1
    return layer._gscale


def _splits(M: int) -> int:
    """Split-K factor for the pixel reduction: chunks of >= 2048 pixels, chunk length a multiple of 8."""
    s = 1
    while s < 32 and M % (2 * s) == 0 and (M // (2 * s)) % 8 == 0 and M // (2 * s) >= 2048:
        s *= 2
    return s


def conv2d_backward(layer, x, gy, sample_id, eps_w, need_gx, seed=None):
    """x [N,Cin,H,W] fp32, gy [N,Cout,Ho,Wo] fp32 -> (gx | None, grad_mu_kernel, grad_rho_kernel)."""
    seed = current_seed() if seed is None else seed
    kh, kw, stride, pad = layer.geometry()
    N, Cin, H, W = x.shape
    Cout = layer.out_channels
    Ho, Wo = gy.shape[2], gy.shape[3]
    mu, rho = layer.mu_kernel.detach(), layer.rho_kernel.detach()
    scale = _layer_scale(layer, gy)
    gyh = ops.nchw_f32_to_nhwc_f16((gy * scale).contiguous())                 # [N, Ho, Wo, Cout] fp16, loss-scaled
    M = N * Ho * Wo

    # ---- dW = X^T dY (reduction over the M pixels), split-K on the batch axis of the grouped GEMM
    splits = _splits(M)
    if M % 8 != 0:
        raise RuntimeError(f"conv backward needs N*Ho*Wo to be a multiple of 8 (got {M})")
    a_t = ops.transpose_chunks_f16(gyh.view(M, Cout), splits)                   # [splits, Cout, Mc]
    xh = ops.nchw_f32_to_nhwc_f16(x.to(F32).contiguous())                     # [N, H, W, Cin]
    b_t = ops.im2col_t_f16(xh, kh, kw, stride, pad, splits)                     # [splits, Kp, Mc]
    dw, _ = ops.gemm_f16(a_t, b_t)                                              # [splits, Cout, Kp] fp16 (x scale)
    gmu = torch.zeros_like(mu)
    grho = torch.zeros_like(rho)
    ops.wgrad_finalize(dw, tuple(mu.shape), 1.0 / scale, rho, gmu, grho,
                       eps=None if eps_w is None else eps_w.reshape(-1).contiguous(), seed=seed,
                       layer_id=layer.layer_uid, sample_id=sample_id)

    # ---- dX = conv(dY (zero-stuffed by the stride), flipped/transposed W_s, stride 1, pad k-1-p)
    gx = None
    if need_gx:
        if Cin % 8 != 0:
            raise RuntimeError("data gradient through a conv with Cin % 8 != 0 (the stems) is never needed")
        wd = ops.sample_weights_dgrad_f16(mu, rho, 1, eps=eps_w, seed=seed, layer_id=layer.layer_uid, sample0=sample_id)
        Hd, Wd = H + 2 * pad - kh + 1, W + 2 * pad - kw + 1
        gyd = gyh if stride == 1 else ops.dilate_f16(gyh, Hd, Wd, stride)
        if kh == 1 and kw == 1:
            y, _ = ops.gemm_f16(gyd.view(1, N * Hd * Wd, Cout), wd)             # 1x1: plain GEMM
            y = y.view(N, Hd, Wd, Cin)
        else:
            y, _ = ops.conv2d_im2col_f16(gyd, wd, 1, kh, kw, 1, kh - 1 - pad)
        gx = ops.nhwc_f16_to_nchw_f32(y) / scale
    return gx, gmu, grho


def linear_backward(layer, x, gy, sample_id, eps, need_gx, seed=None):
    seed = current_seed() if seed is None else seed
    ew, eb = eps
    lead = x.shape[:-1]
    x2 = x.to(F32).reshape(-1, x.shape[-1]).contiguous()
    gy2 = gy.to(F32).reshape(-1, gy.shape[-1]).contiguous()
    mu_w, rho_w = layer.mu_weight.detach(), layer.rho_weight.detach()
    gmw, grw = torch.zeros_like(mu_w), torch.zeros_like(rho_w)
    has_b = layer.mu_bias is not None
    gmb = torch.zeros_like(layer.mu_bias) if has_b else None
    grb = torch.zeros_like(layer.rho_bias) if has_b else None
    gx = ops.sampled_linear_bwd_f32(x2, gy2, mu_w, rho_w, layer.rho_bias.detach() if has_b else None, gmw, grw, gmb, grb,
                                    eps_w=None if ew is None else ew.reshape(-1).contiguous(),
                                    eps_b=None if eb is None else eb.reshape(-1).contiguous(), seed=seed,
                                    layer_id=layer.layer_uid, sample_id=sample_id, need_gx=need_gx)
    if gx is not None:
        gx = gx.view(*lead, x.shape[-1])
    return gx, gmw, grw, gmb, grb
