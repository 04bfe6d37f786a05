"""Drop-in replacement for the part of `bayesian-torch` (==0.5.0) the reference uses:

    from bayesian_torch.models.dnn_to_bnn import dnn_to_bnn, get_kl_loss
    (reference models/model_utils.py:6,26-35; train/multimodal.py:9,114,284; train/unimodal.py:9,130,262)

Same class names, constructor keywords, parameter names (mu_kernel/rho_kernel, mu_weight/rho_weight,
mu_bias/rho_bias — so checkpoints written by reference train/checkpointing.py:40 load unchanged),
`dnn_to_bnn_flag`, `kl_loss()`, `forward(x) -> Tensor`. The arithmetic runs in libmauv_b200.so:
weight sampling + tensor-core contraction for forward, one fused kernel for the KL sum.
Construction/conversion (host-side, one-off) is ordinary PyTorch; forward on a non-CUDA tensor raises.
"""
from __future__ import annotations

import itertools
from typing import Optional

import torch
import torch.nn as nn

from . import _lib, ops

_layer_uid = itertools.count()
_global_seed = 0x5EED_B200


def manual_seed(seed: int) -> None:
    """Seed of the in-kernel Philox eps stream (the reference relies on torch's global RNG)."""
    global _global_seed
    _global_seed = int(seed) & 0xFFFFFFFFFFFFFFFF


def current_seed() -> int:
    return _global_seed


# bayesian-torch draws eps in place into a buffer (`eps = self.eps_kernel.data.normal_()`), and the reference runs S
# forward passes BEFORE one backward (train/multimodal.py:107-138). Autograd saved a view of that buffer for the
# sigma*eps product, so by the time backward runs every pass's saved eps has been overwritten by the LAST pass's
# draw: the reference's grad_rho is sum_s dW_s * eps_S * sigmoid(rho), not sum_s dW_s * eps_s * sigmoid(rho).
# Default here is the mathematically intended gradient (each pass its own eps); set_reference_stale_eps(True)
# reproduces the reference bit for bit in structure (used by the parity tests against the oracle).
_stale_eps = False


def set_reference_stale_eps(flag: bool) -> None:
    global _stale_eps
    _stale_eps = bool(flag)


def reference_stale_eps() -> bool:
    return _stale_eps


def get_rho(sigma: torch.Tensor, delta: float) -> torch.Tensor:
    """MOPED: rho with softplus(rho) ~= delta*|w| (bayesian_torch/utils/util.py)."""
    return torch.log(torch.expm1(delta * torch.abs(sigma)) + 1e-20)


class _BayesBase(nn.Module):
    dnn_to_bnn_flag = False

    def _init_common(self, prior_mean, prior_variance, posterior_mu_init, posterior_rho_init, bias):
        self.prior_mean, self.prior_variance = prior_mean, prior_variance
        self.posterior_mu_init, self.posterior_rho_init = posterior_mu_init, posterior_rho_init
        self.bias = bias
        self.layer_uid = next(_layer_uid)   # Philox layer id when used outside an engine plan
        self._calls = 0                     # Philox sample id for layer-level forward calls
        self.eps_override = None            # tests: (eps_w, eps_b) tensors to use instead of Philox

    def _weight_params(self):
        raise NotImplementedError

    def kl_loss(self) -> torch.Tensor:
        """KL(q||p) of this layer: per-tensor mean for the weight (+ the bias), as kl_div().mean()."""
        pairs = [self._weight_params()]
        if self.mu_bias is not None:
            pairs.append((self.mu_bias, self.rho_bias))
        return _KlFunction.apply(float(self.prior_mean), float(self.prior_variance), *itertools.chain(*pairs))


class _KlFunction(torch.autograd.Function):
    """sum over (mu, rho) pairs of mean(kl_div) with the fused CUDA forward/backward (K4)."""

    @staticmethod
    def forward(ctx, prior_mu, prior_sigma, *params):
        pairs = [(params[i], params[i + 1]) for i in range(0, len(params), 2)]
        plan = ops.KlPlan([(m.detach(), r.detach()) for m, r in pairs], params[0].device)
        ctx.pairs = pairs
        ctx.prior = (prior_mu, prior_sigma)
        return plan.run(prior_mu, prior_sigma)

    @staticmethod
    def backward(ctx, grad_out):
        # gradients via the same kernel: accumulate d KL into fresh buffers, then scale by grad_out
        prior_mu, prior_sigma = ctx.prior
        bufs = [(torch.zeros_like(m), torch.zeros_like(r)) for m, r in ctx.pairs]
        plan = ops.KlPlan([(m.detach(), r.detach()) for m, r in ctx.pairs], ctx.pairs[0][0].device, grads=bufs)
        plan.run(prior_mu, prior_sigma, grad_scale=1.0)
        grads = []
        for gm, gr in bufs:
            grads += [gm * grad_out, gr * grad_out]
        return (None, None, *grads)


class Conv2dReparameterization(_BayesBase):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 prior_mean=0, prior_variance=1, posterior_mu_init=0, posterior_rho_init=-3.0, bias=True):
        super().__init__()
        if in_channels % groups != 0 or out_channels % groups != 0:
            raise ValueError("invalid in_channels size")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = tuple(kernel_size) if isinstance(kernel_size, (tuple, list)) else (kernel_size, kernel_size)
        self.stride, self.padding, self.dilation, self.groups = stride, padding, dilation, groups
        self._init_common(prior_mean, prior_variance, posterior_mu_init, posterior_rho_init, bias)
        shape = (out_channels, in_channels // groups, *self.kernel_size)
        self.mu_kernel = nn.Parameter(torch.empty(shape))
        self.rho_kernel = nn.Parameter(torch.empty(shape))
        if bias:
            self.mu_bias = nn.Parameter(torch.empty(out_channels))
            self.rho_bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("mu_bias", None)
            self.register_parameter("rho_bias", None)
        self.init_parameters()

    def init_parameters(self):
        self.mu_kernel.data.normal_(mean=self.posterior_mu_init, std=0.1)
        self.rho_kernel.data.normal_(mean=self.posterior_rho_init, std=0.1)
        if self.bias:
            self.mu_bias.data.normal_(mean=self.posterior_mu_init, std=0.1)
            self.rho_bias.data.normal_(mean=self.posterior_rho_init, std=0.1)

    def _weight_params(self):
        return self.mu_kernel, self.rho_kernel

    def geometry(self):
        def one(v):
            return v[0] if isinstance(v, (tuple, list)) else v
        if self.groups != 1 or one(self.dilation) != 1:
            raise _lib.MauvError("mauv_b200 conv supports groups=1, dilation=1 (all the reference uses)")
        s, p = self.stride, self.padding
        if isinstance(s, (tuple, list)) and s[0] != s[1] or isinstance(p, (tuple, list)) and p[0] != p[1]:
            raise _lib.MauvError("mauv_b200 conv supports square stride/padding")
        return self.kernel_size[0], self.kernel_size[1], one(s), one(p)

    def forward(self, input, return_kl=True):
        if self.dnn_to_bnn_flag:
            return_kl = False
        from .functional import sampled_conv2d
        out = sampled_conv2d(self, input)
        if return_kl:
            return out, self.kl_loss()
        return out


class LinearReparameterization(_BayesBase):
    def __init__(self, in_features, out_features, prior_mean=0, prior_variance=1, posterior_mu_init=0,
                 posterior_rho_init=-3.0, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self._init_common(prior_mean, prior_variance, posterior_mu_init, posterior_rho_init, bias)
        self.mu_weight = nn.Parameter(torch.empty(out_features, in_features))
        self.rho_weight = nn.Parameter(torch.empty(out_features, in_features))
        if bias:
            self.mu_bias = nn.Parameter(torch.empty(out_features))
            self.rho_bias = nn.Parameter(torch.empty(out_features))
        else:
            self.register_parameter("mu_bias", None)
            self.register_parameter("rho_bias", None)
        self.init_parameters()

    def init_parameters(self):
        self.mu_weight.data.normal_(mean=self.posterior_mu_init, std=0.1)
        self.rho_weight.data.normal_(mean=self.posterior_rho_init, std=0.1)
        if self.bias:
            self.mu_bias.data.normal_(mean=self.posterior_mu_init, std=0.1)
            self.rho_bias.data.normal_(mean=self.posterior_rho_init, std=0.1)

    def _weight_params(self):
        return self.mu_weight, self.rho_weight

    def forward(self, input, return_kl=True):
        if self.dnn_to_bnn_flag:
            return_kl = False
        from .functional import sampled_linear
        out = sampled_linear(self, input)
        if return_kl:
            return out, self.kl_loss()
        return out


_LAYERS = {"Conv2dReparameterization": Conv2dReparameterization,
           "LinearReparameterization": LinearReparameterization}


def bnn_conv_layer(params, d):
    layer_fn = _LAYERS.get(d.__class__.__name__ + params["type"])
    if layer_fn is None:
        raise _lib.MauvError(f"mauv_b200 has no Bayesian layer for {d.__class__.__name__}{params['type']}")
    layer = layer_fn(in_channels=d.in_channels, out_channels=d.out_channels, kernel_size=d.kernel_size,
                     stride=d.stride, padding=d.padding, dilation=d.dilation, groups=d.groups,
                     prior_mean=params["prior_mu"], prior_variance=params["prior_sigma"],
                     posterior_mu_init=params["posterior_mu_init"],
                     posterior_rho_init=params["posterior_rho_init"], bias=d.bias is not None)
    if params["moped_enable"]:
        delta = params["moped_delta"]
        layer.mu_kernel.data.copy_(d.weight.data)
        layer.rho_kernel.data.copy_(get_rho(d.weight.data, delta))
        if layer.mu_bias is not None:
            layer.mu_bias.data.copy_(d.bias.data)
            layer.rho_bias.data.copy_(get_rho(d.bias.data, delta))
    layer.dnn_to_bnn_flag = True
    return layer.to(d.weight.device)


def bnn_linear_layer(params, d):
    layer_fn = _LAYERS.get(d.__class__.__name__ + params["type"])
    if layer_fn is None:
        raise _lib.MauvError(f"mauv_b200 has no Bayesian layer for {d.__class__.__name__}{params['type']}")
    layer = layer_fn(in_features=d.in_features, out_features=d.out_features, prior_mean=params["prior_mu"],
                     prior_variance=params["prior_sigma"], posterior_mu_init=params["posterior_mu_init"],
                     posterior_rho_init=params["posterior_rho_init"], bias=d.bias is not None)
    if params["moped_enable"]:
        delta = params["moped_delta"]
        layer.mu_weight.data.copy_(d.weight.data)
        layer.rho_weight.data.copy_(get_rho(d.weight.data, delta))
        if layer.mu_bias is not None:
            layer.mu_bias.data.copy_(d.bias.data)
            layer.rho_bias.data.copy_(get_rho(d.bias.data, delta))
    layer.dnn_to_bnn_flag = True
    return layer.to(d.weight.device)


def dnn_to_bnn(m: nn.Module, bnn_prior_parameters: dict) -> None:
    """In-place: every nn.Conv*/nn.Linear leaf becomes its Bayesian twin (MOPED init optional)."""
    for name, value in list(m._modules.items()):
        if m._modules[name]._modules:
            dnn_to_bnn(m._modules[name], bnn_prior_parameters)
        elif "Conv" in m._modules[name].__class__.__name__:
            setattr(m, name, bnn_conv_layer(bnn_prior_parameters, m._modules[name]))
        elif "Linear" in m._modules[name].__class__.__name__:
            setattr(m, name, bnn_linear_layer(bnn_prior_parameters, m._modules[name]))
    return


def bayesian_layers(m: nn.Module):
    return [(n, l) for n, l in m.named_modules() if hasattr(l, "kl_loss")]


def get_kl_loss(m: nn.Module) -> Optional[torch.Tensor]:
    """Sum of every Bayesian layer's kl_loss() — one fused launch over all 174 layers' parameters
    instead of the reference's ~12 elementwise/reduce launches per layer."""
    layers = [l for _, l in bayesian_layers(m)]
    if not layers:
        return None
    params = []
    priors = set()
    for l in layers:
        params += list(l._weight_params())
        if l.mu_bias is not None:
            params += [l.mu_bias, l.rho_bias]
        priors.add((float(l.prior_mean), float(l.prior_variance)))
    if len(priors) != 1:
        # heterogeneous priors: fall back to per-layer launches (still the CUDA kernel)
        out = None
        for l in layers:
            out = l.kl_loss() if out is None else out + l.kl_loss()
        return out
    (pm, ps), = priors
    return _KlFunction.apply(pm, ps, *params)
