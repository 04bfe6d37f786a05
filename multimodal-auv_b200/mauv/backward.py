"""Backward of the sampled conv / linear layers (autograd of bayesian-torch's forward:
dX, dmu = dW, drho = dW * eps * sigmoid(rho))."""
from __future__ import annotations

from . import _lib


def conv2d_backward(layer, x, gy, sample_id, eps_w, need_gx):
    raise _lib.MauvError("sampled conv backward: CUDA kernels not built yet (no PyTorch fallback by design)")


def linear_backward(layer, x, gy, sample_id, eps, need_gx):
    raise _lib.MauvError("sampled linear backward: CUDA kernels not built yet (no PyTorch fallback by design)")
