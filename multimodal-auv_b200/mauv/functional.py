"""Layer-level dispatch: what `Conv2dReparameterization.forward(x)` / `LinearReparameterization.forward(x)`
run when the model is called the way the reference calls it (one MC pass per `model(...)` call:
train/multimodal.py:112, train/unimodal.py:128, inference/predictors.py:61).

x (NCHW fp32, CUDA) -> NHWC fp16 -> sampled weights (Philox, sample id = per-layer call counter, or the
layer's `eps_override`) -> tcgen05 implicit GEMM -> NCHW fp32. The S-batched fast path is engine.MCEngine;
this module exists so that the drop-in modules behave like bayesian-torch's, one call = one fresh sample.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from .bayesian import current_seed, reference_stale_eps

F16, F32 = torch.float16, torch.float32


def _need_cuda(x: torch.Tensor, what: str) -> None:
    if not x.is_cuda:
        raise _lib.MauvError(f"{what}: input is on {x.device}; mauv_b200 layers run on sm_100a CUDA only "
                             "(no CPU fallback)")


def _layer_eps(layer):
    ov = layer.eps_override
    if ov is None:
        return None, None
    ew, eb = ov
    return (None if ew is None else ew.to(F32).contiguous().unsqueeze(0),
            None if eb is None else eb.to(F32).contiguous().unsqueeze(0))


def conv2d_forward(layer, x: torch.Tensor, sample_id: int, eps_w=None, seed=None) -> torch.Tensor:
    """One sampled conv: x [N, Cin, H, W] fp32 -> [N, Cout, Ho, Wo] fp32."""
    kh, kw, stride, pad = layer.geometry()
    N, Cin, H, W = x.shape
    seed = current_seed() if seed is None else seed
    w = ops.sample_weights_f16(layer.mu_kernel.detach(), layer.rho_kernel.detach(), 1, eps=eps_w, seed=seed,
                               layer_id=layer.layer_uid, sample0=sample_id)
    Ho = (H + 2 * pad - kh) // stride + 1
    Wo = (W + 2 * pad - kw) // stride + 1
    x = x.to(F32).contiguous()
    if Cin % 64 != 0:
        a = ops.stem_im2col_f16(x, kh, kw, stride, pad)                   # explicit im2col (tiny Cin)
        y, _ = ops.gemm_f16(a, w, shared_a=True)
        y = y.view(N, Ho, Wo, layer.out_channels)
    else:
        xh = ops.nchw_f32_to_nhwc_f16(x)
        if kh == 1 and kw == 1 and stride == 1 and pad == 0:
            y, _ = ops.gemm_f16(xh.view(1, N * H * W, Cin), w)
            y = y.view(N, H, W, layer.out_channels)
        else:
            y, _ = ops.conv2d_im2col_f16(xh, w, 1, kh, kw, stride, pad)
    out = ops.nhwc_f16_to_nchw_f32(y)
    if layer.mu_bias is not None:
        raise _lib.MauvError("Bayesian conv with bias is not used by the reference and not implemented")
    return out


def linear_forward(layer, x: torch.Tensor, sample_id: int, eps_w=None, eps_b=None, seed=None) -> torch.Tensor:
    seed = current_seed() if seed is None else seed
    lead = x.shape[:-1]
    x2 = x.to(F32).reshape(1, -1, x.shape[-1]).contiguous()
    y = ops.sampled_linear_f32(x2, layer.mu_weight.detach(), layer.rho_weight.detach(),
                               None if layer.mu_bias is None else layer.mu_bias.detach(),
                               None if layer.rho_bias is None else layer.rho_bias.detach(),
                               eps_w=eps_w, eps_b=eps_b, seed=seed, layer_id=layer.layer_uid, sample0=sample_id)
    return y.view(*lead, layer.out_features)


class _SampledConv2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mu, rho, layer):
        sid = layer._calls
        layer._calls += 1
        ew, _ = _layer_eps(layer)
        ctx.layer, ctx.sid, ctx.eps_w = layer, sid, ew
        layer._last_draw = (sid, ew, None)
        ctx.save_for_backward(x)
        return conv2d_forward(layer, x, sid, ew)

    @staticmethod
    def backward(ctx, gy):
        from .backward import conv2d_backward
        (x,) = ctx.saved_tensors
        sid, ew = ctx.sid, ctx.eps_w
        if reference_stale_eps():            # the reference's saved eps buffer holds the LAST pass's draw
            # dX and dW must still use this pass's weights; only the eps factor of d(rho) is stale
            gx, gmu, _ = conv2d_backward(ctx.layer, x, gy, sid, ew, ctx.needs_input_grad[0])
            lsid, lew, _ = ctx.layer._last_draw
            grho = rho_grad_from_dw(ctx.layer, gmu, lsid, lew)
            return gx, gmu, grho, None
        gx, gmu, grho = conv2d_backward(ctx.layer, x, gy, sid, ew, ctx.needs_input_grad[0])
        return gx, gmu, grho, None


class _SampledLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mu_w, rho_w, mu_b, rho_b, layer):
        sid = layer._calls
        layer._calls += 1
        ew, eb = _layer_eps(layer)
        ctx.layer, ctx.sid, ctx.eps = layer, sid, (ew, eb)
        layer._last_draw = (sid, ew, eb)
        ctx.save_for_backward(x)
        return linear_forward(layer, x, sid, ew, eb)

    @staticmethod
    def backward(ctx, gy):
        from .backward import linear_backward
        (x,) = ctx.saved_tensors
        gx, gmw, grw, gmb, grb = linear_backward(ctx.layer, x, gy, ctx.sid, ctx.eps, ctx.needs_input_grad[0])
        if reference_stale_eps():
            lsid, lew, leb = ctx.layer._last_draw
            grw = rho_grad_from_dw(ctx.layer, gmw, lsid, lew)
            if gmb is not None:
                grb = rho_grad_from_dw(ctx.layer, gmb, lsid, leb, bias=True)
        return gx, gmw, grw, gmb, grb, None


def rho_grad_from_dw(layer, dw: torch.Tensor, sample_id: int, eps, bias: bool = False) -> torch.Tensor:
    """d(rho) = dW * eps * sigmoid(rho) for a given eps draw (compat path only; the normal path fuses this)."""
    from .bayesian import current_seed as _seed
    if bias:
        rho, lid = layer.rho_bias.detach(), layer.layer_uid | 0x80000000
    else:
        rho, lid = layer._weight_params()[1].detach(), layer.layer_uid
    if eps is None:
        eps = ops.philox_normal(dw.numel(), seed=_seed(), layer_id=lid, sample_id=sample_id, device=dw.device)
    return dw * eps.reshape(dw.shape) * torch.sigmoid(rho)


def sampled_conv2d(layer, x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x, "Conv2dReparameterization.forward")
    return _SampledConv2d.apply(x, layer.mu_kernel, layer.rho_kernel, layer)


def sampled_linear(layer, x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x, "LinearReparameterization.forward")
    return _SampledLinear.apply(x, layer.mu_weight, layer.rho_weight, layer.mu_bias, layer.rho_bias, layer)
