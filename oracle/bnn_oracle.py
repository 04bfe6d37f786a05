"""CPU ORACLE — test infrastructure only (tests/, __graft_entry__.smoke(), bench.py cpu_baseline /
--impl reference). Never imported by the product package `mauv`.

Plain-PyTorch CPU restatement of the reference's Monte-Carlo Bayesian hot path:

  * the third-party layer library the reference calls but does not vendor —
    `bayesian-torch` pinned ==0.5.0 (reference reqirements.txt:4, pyproject.toml:45) —
    restated from its published algorithm:
      Conv2dReparameterization / LinearReparameterization .forward, .kl_loss
      BaseVariationalLayer_.kl_div, dnn_to_bnn, get_kl_loss, get_rho (MOPED);
    call sites in the reference: models/model_utils.py:6,26-35; train/multimodal.py:9,114,284;
    train/unimodal.py:9,130,262.
  * the network topology: models/base_models.py:7-90 (ResNet50Custom, Identity,
    AdditiveAttention, MultiModalModel) over torchvision resnet50;
  * the MC statistics / losses of inference/predictors.py:65-84, train/multimodal.py:104-130,
    287-310 and train/unimodal.py:125-142, 282-308.

PARITY UNPINNED (bayesian-torch part): the reference's tests mock `get_kl_loss`/`dnn_to_bnn`
(unittests/test_train.py:227, unittests/test_model.py:94-97) and ship no golden vector for a
Bayesian layer, a KL value or an uncertainty value, and the package itself cannot be installed
offline. The model/driver part IS pinned: oracle/make_golden.py runs the reference's own
models/base_models.py and inference/predictors.py (loaded by file path) on top of these layers
and tests/test_oracle.py checks this file against those outputs.

Two additions over the reference semantics, both opt-in: eps injection/capture (so the CUDA
path and the oracle can share identical noise) and a float64 mode.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# bayesian-torch 0.5.0 restatement
# --------------------------------------------------------------------------------------


def get_rho(sigma: torch.Tensor, delta: float) -> torch.Tensor:
    """bayesian_torch/utils/util.py get_rho: rho such that softplus(rho) ~= delta*|w| (MOPED)."""
    return torch.log(torch.expm1(delta * torch.abs(sigma)) + 1e-20)


def kl_div(mu_q, sigma_q, mu_p, sigma_p) -> torch.Tensor:
    """BaseVariationalLayer_.kl_div: elementwise KL(N(mu_q,sigma_q)||N(mu_p,sigma_p)), then .mean()."""
    kl = torch.log(sigma_p) - torch.log(sigma_q) + (sigma_q ** 2 + (mu_q - mu_p) ** 2) / (2 * (sigma_p ** 2)) - 0.5
    return kl.mean()


# bayesian-torch draws eps IN PLACE into a module buffer and multiplies by that buffer's `.data`
# (`eps_kernel = self.eps_kernel.data.normal_()`), so autograd saves a view of the buffer for d(sigma*eps)/d(sigma).
# When several forward passes precede one backward - exactly the reference's S-pass loops (train/multimodal.py:107-138,
# train/unimodal.py:127-145) - every pass's saved eps is overwritten by the LAST draw: grad_rho = sum_s dW_s*eps_S*sig'.
# True (default) restates that behaviour faithfully; False clones eps per pass (the mathematically intended gradient).
STALE_EPS_QUIRK = True


class _ReparamBase(nn.Module):
    dnn_to_bnn_flag = False
    # test aid (not reference behaviour): round the sampled weight to fp16, the operand precision of the CUDA
    # tensor-core path, so that tests can separate "engine logic" from "operand precision" (emulate_fp16_pipeline)
    round_weight_fp16 = False

    def _maybe_round(self, w: torch.Tensor) -> torch.Tensor:
        return w.half().to(w.dtype) if self.round_weight_fp16 else w

    def _draw(self, buf: torch.Tensor, injected: Optional[torch.Tensor]) -> torch.Tensor:
        if injected is not None:
            buf.data.copy_(injected.to(buf.dtype))
        else:
            buf.data.normal_()
        return buf.data if STALE_EPS_QUIRK else buf.data.clone()


class Conv2dReparameterization(_ReparamBase):
    """layers/variational_layers/conv_variational.py Conv2dReparameterization (0.5.0)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 prior_mean=0, prior_variance=1, posterior_mu_init=0, posterior_rho_init=-3.0, bias=True):
        super().__init__()
        if in_channels % groups != 0 or out_channels % groups != 0:
            raise ValueError("invalid in_channels size")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = kernel_size if isinstance(kernel_size, tuple) else (kernel_size, kernel_size)
        self.stride, self.padding, self.dilation, self.groups = stride, padding, dilation, groups
        self.prior_mean, self.prior_variance = prior_mean, prior_variance
        self.posterior_mu_init, self.posterior_rho_init = posterior_mu_init, posterior_rho_init
        self.bias = bias
        shape = (out_channels, in_channels // groups, self.kernel_size[0], self.kernel_size[1])
        self.mu_kernel = nn.Parameter(torch.empty(shape))
        self.rho_kernel = nn.Parameter(torch.empty(shape))
        self.register_buffer("eps_kernel", torch.empty(shape), persistent=False)
        self.register_buffer("prior_weight_mu", torch.empty(shape), persistent=False)
        self.register_buffer("prior_weight_sigma", torch.empty(shape), persistent=False)
        if bias:
            self.mu_bias = nn.Parameter(torch.empty(out_channels))
            self.rho_bias = nn.Parameter(torch.empty(out_channels))
            self.register_buffer("eps_bias", torch.empty(out_channels), persistent=False)
            self.register_buffer("prior_bias_mu", torch.empty(out_channels), persistent=False)
            self.register_buffer("prior_bias_sigma", torch.empty(out_channels), persistent=False)
        else:
            self.register_parameter("mu_bias", None)
            self.register_parameter("rho_bias", None)
            self.register_buffer("eps_bias", None, persistent=False)
            self.register_buffer("prior_bias_mu", None, persistent=False)
            self.register_buffer("prior_bias_sigma", None, persistent=False)
        self.injected_eps_kernel: Optional[torch.Tensor] = None
        self.injected_eps_bias: Optional[torch.Tensor] = None
        self.init_parameters()

    def init_parameters(self):
        self.prior_weight_mu.fill_(self.prior_mean)
        self.prior_weight_sigma.fill_(self.prior_variance)
        self.mu_kernel.data.normal_(mean=self.posterior_mu_init, std=0.1)
        self.rho_kernel.data.normal_(mean=self.posterior_rho_init, std=0.1)
        if self.bias:
            self.prior_bias_mu.fill_(self.prior_mean)
            self.prior_bias_sigma.fill_(self.prior_variance)
            self.mu_bias.data.normal_(mean=self.posterior_mu_init, std=0.1)
            self.rho_bias.data.normal_(mean=self.posterior_rho_init, std=0.1)

    def kl_loss(self):
        sigma_weight = torch.log1p(torch.exp(self.rho_kernel))
        kl = kl_div(self.mu_kernel, sigma_weight, self.prior_weight_mu, self.prior_weight_sigma)
        if self.bias:
            sigma_bias = torch.log1p(torch.exp(self.rho_bias))
            kl = kl + kl_div(self.mu_bias, sigma_bias, self.prior_bias_mu, self.prior_bias_sigma)
        return kl

    def forward(self, input, return_kl=True):
        if self.dnn_to_bnn_flag:
            return_kl = False
        sigma_weight = torch.log1p(torch.exp(self.rho_kernel))
        eps_kernel = self._draw(self.eps_kernel, self.injected_eps_kernel)
        weight = self._maybe_round(self.mu_kernel + (sigma_weight * eps_kernel))
        if return_kl:
            kl_weight = kl_div(self.mu_kernel, sigma_weight, self.prior_weight_mu, self.prior_weight_sigma)
        bias = None
        if self.bias:
            sigma_bias = torch.log1p(torch.exp(self.rho_bias))
            eps_bias = self._draw(self.eps_bias, self.injected_eps_bias)
            bias = self.mu_bias + (sigma_bias * eps_bias)
            if return_kl:
                kl_bias = kl_div(self.mu_bias, sigma_bias, self.prior_bias_mu, self.prior_bias_sigma)
        out = F.conv2d(input, weight, bias, self.stride, self.padding, self.dilation, self.groups)
        if return_kl:
            kl = kl_weight + kl_bias if self.bias else kl_weight
            return out, kl
        return out


class LinearReparameterization(_ReparamBase):
    """layers/variational_layers/linear_variational.py LinearReparameterization (0.5.0)."""

    def __init__(self, in_features, out_features, prior_mean=0, prior_variance=1, posterior_mu_init=0,
                 posterior_rho_init=-3.0, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.prior_mean, self.prior_variance = prior_mean, prior_variance
        self.posterior_mu_init, self.posterior_rho_init = posterior_mu_init, posterior_rho_init
        self.bias = bias
        self.mu_weight = nn.Parameter(torch.empty(out_features, in_features))
        self.rho_weight = nn.Parameter(torch.empty(out_features, in_features))
        self.register_buffer("eps_weight", torch.empty(out_features, in_features), persistent=False)
        self.register_buffer("prior_weight_mu", torch.empty(out_features, in_features), persistent=False)
        self.register_buffer("prior_weight_sigma", torch.empty(out_features, in_features), persistent=False)
        if bias:
            self.mu_bias = nn.Parameter(torch.empty(out_features))
            self.rho_bias = nn.Parameter(torch.empty(out_features))
            self.register_buffer("eps_bias", torch.empty(out_features), persistent=False)
            self.register_buffer("prior_bias_mu", torch.empty(out_features), persistent=False)
            self.register_buffer("prior_bias_sigma", torch.empty(out_features), persistent=False)
        else:
            self.register_parameter("mu_bias", None)
            self.register_parameter("rho_bias", None)
            self.register_buffer("eps_bias", None, persistent=False)
            self.register_buffer("prior_bias_mu", None, persistent=False)
            self.register_buffer("prior_bias_sigma", None, persistent=False)
        self.injected_eps_weight: Optional[torch.Tensor] = None
        self.injected_eps_bias: Optional[torch.Tensor] = None
        self.init_parameters()

    def init_parameters(self):
        self.prior_weight_mu.fill_(self.prior_mean)
        self.prior_weight_sigma.fill_(self.prior_variance)
        self.mu_weight.data.normal_(mean=self.posterior_mu_init, std=0.1)
        self.rho_weight.data.normal_(mean=self.posterior_rho_init, std=0.1)
        if self.bias:
            self.prior_bias_mu.fill_(self.prior_mean)
            self.prior_bias_sigma.fill_(self.prior_variance)
            self.mu_bias.data.normal_(mean=self.posterior_mu_init, std=0.1)
            self.rho_bias.data.normal_(mean=self.posterior_rho_init, std=0.1)

    def kl_loss(self):
        sigma_weight = torch.log1p(torch.exp(self.rho_weight))
        kl = kl_div(self.mu_weight, sigma_weight, self.prior_weight_mu, self.prior_weight_sigma)
        if self.bias:
            sigma_bias = torch.log1p(torch.exp(self.rho_bias))
            kl = kl + kl_div(self.mu_bias, sigma_bias, self.prior_bias_mu, self.prior_bias_sigma)
        return kl

    def forward(self, input, return_kl=True):
        if self.dnn_to_bnn_flag:
            return_kl = False
        sigma_weight = torch.log1p(torch.exp(self.rho_weight))
        eps_weight = self._draw(self.eps_weight, self.injected_eps_weight)
        weight = self.mu_weight + (sigma_weight * eps_weight)
        if return_kl:
            kl_weight = kl_div(self.mu_weight, sigma_weight, self.prior_weight_mu, self.prior_weight_sigma)
        bias = None
        if self.bias:
            sigma_bias = torch.log1p(torch.exp(self.rho_bias))
            eps_bias = self._draw(self.eps_bias, self.injected_eps_bias)
            bias = self.mu_bias + (sigma_bias * eps_bias)
            if return_kl:
                kl_bias = kl_div(self.mu_bias, sigma_bias, self.prior_bias_mu, self.prior_bias_sigma)
        out = F.linear(input, weight, bias)
        if return_kl:
            kl = kl_weight + kl_bias if self.bias else kl_weight
            return out, kl
        return out


_LAYERS = {"Conv2dReparameterization": Conv2dReparameterization,
           "LinearReparameterization": LinearReparameterization}


def bnn_conv_layer(params, d):
    layer_fn = _LAYERS[d.__class__.__name__ + params["type"]]
    bnn_layer = layer_fn(in_channels=d.in_channels, out_channels=d.out_channels, kernel_size=d.kernel_size,
                         stride=d.stride, padding=d.padding, dilation=d.dilation, groups=d.groups,
                         prior_mean=params["prior_mu"], prior_variance=params["prior_sigma"],
                         posterior_mu_init=params["posterior_mu_init"],
                         posterior_rho_init=params["posterior_rho_init"], bias=d.bias is not None)
    if params["moped_enable"]:
        delta = params["moped_delta"]
        bnn_layer.mu_kernel.data.copy_(d.weight.data)
        bnn_layer.rho_kernel.data.copy_(get_rho(d.weight.data, delta))
        if bnn_layer.mu_bias is not None:
            bnn_layer.mu_bias.data.copy_(d.bias.data)
            bnn_layer.rho_bias.data.copy_(get_rho(d.bias.data, delta))
    bnn_layer.dnn_to_bnn_flag = True
    return bnn_layer


def bnn_linear_layer(params, d):
    layer_fn = _LAYERS[d.__class__.__name__ + params["type"]]
    bnn_layer = layer_fn(in_features=d.in_features, out_features=d.out_features,
                         prior_mean=params["prior_mu"], prior_variance=params["prior_sigma"],
                         posterior_mu_init=params["posterior_mu_init"],
                         posterior_rho_init=params["posterior_rho_init"], bias=d.bias is not None)
    if params["moped_enable"]:
        delta = params["moped_delta"]
        bnn_layer.mu_weight.data.copy_(d.weight.data)
        bnn_layer.rho_weight.data.copy_(get_rho(d.weight.data, delta))
        if bnn_layer.mu_bias is not None:
            bnn_layer.mu_bias.data.copy_(d.bias.data)
            bnn_layer.rho_bias.data.copy_(get_rho(d.bias.data, delta))
    bnn_layer.dnn_to_bnn_flag = True
    return bnn_layer


def dnn_to_bnn(m: nn.Module, bnn_prior_parameters: dict) -> None:
    """bayesian_torch/models/dnn_to_bnn.py dnn_to_bnn: in-place recursive replacement."""
    for name, value in list(m._modules.items()):
        if m._modules[name]._modules:
            dnn_to_bnn(m._modules[name], bnn_prior_parameters)
        elif "Conv" in m._modules[name].__class__.__name__:
            setattr(m, name, bnn_conv_layer(bnn_prior_parameters, m._modules[name]))
        elif "Linear" in m._modules[name].__class__.__name__:
            setattr(m, name, bnn_linear_layer(bnn_prior_parameters, m._modules[name]))
    return


def get_kl_loss(m: nn.Module):
    """bayesian_torch/models/dnn_to_bnn.py get_kl_loss: sum of per-layer kl_loss()."""
    kl_loss = None
    for layer in m.modules():
        if hasattr(layer, "kl_loss"):
            if kl_loss is None:
                kl_loss = layer.kl_loss()
            else:
                kl_loss = kl_loss + layer.kl_loss()
    return kl_loss


# The canonical prior dict: reference Examples/Example_Inference_model.py:51-59, cli.py:120-128
DEFAULT_PRIOR = {
    "prior_mu": 0.0, "prior_sigma": 1.0, "posterior_mu_init": 0.0, "posterior_rho_init": -3.0,
    "type": "Reparameterization", "moped_enable": True, "moped_delta": 0.1,
}

# --------------------------------------------------------------------------------------
# models/base_models.py restatement (random-init: no ImageNet download offline)
# --------------------------------------------------------------------------------------


class ResNet50Custom(nn.Module):
    """models/base_models.py:7-29 with weights=None."""

    def __init__(self, input_channels, num_classes):
        super().__init__()
        from torchvision.models import resnet50
        self.input_channels = input_channels
        self.model = resnet50(weights=None)
        self.model.conv1 = nn.Conv2d(input_channels, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.model.fc = nn.Linear(self.model.fc.in_features, num_classes)

    def forward(self, x):
        return self.model(x)

    def get_feature_size(self):
        return self.model.fc.in_features


class Identity(nn.Module):
    """models/base_models.py:31-33"""

    def forward(self, x):
        return x


class AdditiveAttention(nn.Module):
    """models/base_models.py:35-52"""

    def __init__(self, d_model, hidden_dim=128):
        super().__init__()
        self.query_projection = nn.Linear(d_model, hidden_dim)
        self.key_projection = nn.Linear(d_model, hidden_dim)
        self.value_projection = nn.Linear(d_model, hidden_dim)
        self.attention_mechanism = nn.Linear(hidden_dim, hidden_dim)

    def forward(self, query):
        keys = self.key_projection(query)
        values = self.value_projection(query)
        queries = self.query_projection(query)
        attention_scores = torch.tanh(queries + keys)
        attention_weights = F.softmax(self.attention_mechanism(attention_scores), dim=1)
        return values * attention_weights


class MultiModalModel(nn.Module):
    """models/base_models.py:54-90"""

    def __init__(self, image_model_feat, bathy_model_feat, sss_model_feat, num_classes,
                 attention_type="scaled_dot_product"):
        super().__init__()
        self.image_model_feat = image_model_feat
        self.bathy_model_feat = bathy_model_feat
        self.sss_model_feat = sss_model_feat
        self.fc = nn.Linear(384, 1284)
        self.fc1 = nn.Linear(1284, 32)
        self.fc2 = nn.Linear(32, int(num_classes))
        self.attention_type = attention_type
        self.attention_image = AdditiveAttention(2048)
        self.attention_bathy = AdditiveAttention(2048)
        self.attention_sss = AdditiveAttention(2048)

    def forward(self, inputs, bathy_tensor, sss_image):
        image_features = self.image_model_feat(inputs)
        bathy_features = self.bathy_model_feat(bathy_tensor)
        sss_features = self.sss_model_feat(sss_image)
        a = self.attention_image(image_features)
        b = self.attention_bathy(bathy_features)
        c = self.attention_sss(sss_features)
        combined = torch.cat([a, b, c], dim=1)
        return self.fc2(self.fc1(self.fc(combined)))


def feature_extractor(input_channels: int = 3) -> nn.Module:
    """models/model_utils.py:52-64 with weights=None."""
    from torchvision.models import resnet50
    model = resnet50(weights=None)
    if input_channels == 1:
        model.conv1 = nn.Conv2d(1, 64, kernel_size=(7, 7), stride=(2, 2), padding=(3, 3), bias=False)
    model.fc = Identity()
    return model


def define_models(num_classes: int, prior: dict = DEFAULT_PRIOR, seed: Optional[int] = 1234,
                  unimodal: bool = True) -> Dict[str, nn.Module]:
    """models/model_utils.py:10-45 (random init + MOPED, as BASELINE.md §4 specifies)."""
    if seed is not None:
        torch.manual_seed(seed)
    out: Dict[str, nn.Module] = {}
    if unimodal:
        out["image_model"] = ResNet50Custom(3, num_classes)
        out["bathy_model"] = ResNet50Custom(3, num_classes)
        out["sss_model"] = ResNet50Custom(1, num_classes)
        for k in ("image_model", "bathy_model", "sss_model"):
            dnn_to_bnn(out[k], prior)
    f_img, f_bat, f_sss = feature_extractor(), feature_extractor(), feature_extractor(1)
    mm = MultiModalModel(f_img, f_bat, f_sss, num_classes)
    dnn_to_bnn(mm, prior)
    out.update(multimodal_model=mm, image_model_feat=f_img, bathy_model_feat=f_bat, sss_model_feat=f_sss)
    return out


# --------------------------------------------------------------------------------------
# eps capture / injection
# --------------------------------------------------------------------------------------


def emulate_fp16_pipeline(model: nn.Module, unrounded_convs=()):
    """Test aid: make the oracle round exactly where the CUDA engine stores fp16 - conv operands (inputs and sampled
    weights), raw conv outputs (except `unrounded_convs`: the convs whose BatchNorm is fused into the GEMM epilogue),
    max-pool and bottleneck outputs - while BatchNorm, the head and all statistics stay fp32 as in the engine.
    Differences that remain against the engine are accumulation order only. Returns the hook handles."""
    from torchvision.models.resnet import Bottleneck
    hs = []

    def r16(t):
        return t.half().float()

    for name, m in model.named_modules():
        if hasattr(m, "mu_kernel"):
            m.round_weight_fp16 = True
            hs.append(m.register_forward_pre_hook(lambda mod, inp: tuple(r16(i) for i in inp)))
            if name not in unrounded_convs:
                hs.append(m.register_forward_hook(lambda mod, inp, out: r16(out)))
        elif isinstance(m, (Bottleneck, nn.MaxPool2d)):
            hs.append(m.register_forward_hook(lambda mod, inp, out: r16(out)))
    return hs


def stop_emulation(model: nn.Module, hooks) -> None:
    for h in hooks:
        h.remove()
    for m in model.modules():
        if hasattr(m, "mu_kernel"):
            m.round_weight_fp16 = False


def bayesian_layers(model: nn.Module) -> List[tuple]:
    """[(qualified name, layer)] in module registration order (= get_kl_loss order)."""
    return [(n, m) for n, m in model.named_modules() if hasattr(m, "kl_loss")]


def draw_eps(model: nn.Module, S: int, seed: int) -> Dict[str, Dict[str, torch.Tensor]]:
    """eps[name] = {"w": [S, *weight.shape], "b": [S, out] | None} from a seeded CPU generator."""
    gen = torch.Generator().manual_seed(seed)
    eps = {}
    for name, layer in bayesian_layers(model):
        w = layer.mu_kernel if hasattr(layer, "mu_kernel") else layer.mu_weight
        e = {"w": torch.randn((S, *w.shape), generator=gen), "b": None}
        if layer.mu_bias is not None:
            e["b"] = torch.randn((S, layer.mu_bias.numel()), generator=gen)
        eps[name] = e
    return eps


def inject_eps(model: nn.Module, eps: Optional[dict], s: int) -> None:
    for name, layer in bayesian_layers(model):
        e = None if eps is None else eps[name]
        if hasattr(layer, "mu_kernel"):
            layer.injected_eps_kernel = None if e is None else e["w"][s]
        else:
            layer.injected_eps_weight = None if e is None else e["w"][s]
        layer.injected_eps_bias = None if (e is None or e["b"] is None) else e["b"][s]


def mc_logits(model: nn.Module, inputs: Iterable[torch.Tensor], S: int, eps: Optional[dict] = None,
              grad: bool = False) -> torch.Tensor:
    """The S-pass loop of predictors.py:54-66 / multimodal.py:107-118 -> logits [S, B, C] (train mode)."""
    model.train()
    outs = []
    ctx = torch.enable_grad() if grad else torch.no_grad()
    with ctx:
        for s in range(S):
            inject_eps(model, eps, s)
            outs.append(model(*inputs))
    inject_eps(model, None, 0)
    return torch.stack(outs, dim=0)


# --------------------------------------------------------------------------------------
# MC statistics and losses (L3 math)
# --------------------------------------------------------------------------------------


def predictor_stats(logits: torch.Tensor) -> Dict[str, torch.Tensor]:
    """inference/predictors.py:65-84 (given the stacked per-pass logits)."""
    prob = F.softmax(logits, dim=2)
    predictive_uncertainty = torch.var(prob, dim=0).mean(dim=1)
    epsilon = 1e-7
    entropy_per_mc = -torch.sum(prob * torch.log(prob + epsilon), dim=-1)
    aleatoric = torch.mean(entropy_per_mc, dim=0)
    mean_prob = torch.mean(prob, dim=0)
    return {"predicted_class": torch.argmax(mean_prob, dim=1), "predictive_uncertainty": predictive_uncertainty,
            "aleatoric_uncertainty": aleatoric, "mean_prob": mean_prob}


def multimodal_eval_stats(logits: torch.Tensor) -> Dict[str, torch.Tensor]:
    """train/multimodal.py:287-310."""
    epsilon = 1e-8
    softmax_stack = F.softmax(logits, dim=2)
    output_mean = torch.mean(logits, dim=0)
    _, predicted = torch.max(output_mean, 1)
    mean_softmax = torch.mean(softmax_stack, dim=0)
    predictive = -torch.sum(mean_softmax * torch.log(mean_softmax + epsilon), dim=1)
    entropy_per = -torch.sum(softmax_stack * torch.log(softmax_stack + epsilon), dim=2)
    aleatoric = torch.mean(entropy_per, dim=0)
    return {"output_mean": output_mean, "predicted": predicted, "predictive_uncertainty": predictive,
            "aleatoric_uncertainty": aleatoric, "model_uncertainty": predictive - aleatoric}


def unimodal_eval_stats(logits: torch.Tensor) -> Dict[str, torch.Tensor]:
    """train/unimodal.py:268-308."""
    output_mean = torch.mean(logits, dim=0)
    probs_mean = F.softmax(output_mean, dim=1)
    _, predicted = torch.max(probs_mean, 1)
    probs = F.softmax(logits, dim=2)
    predictive_uncertainty = torch.var(probs, dim=0).mean(dim=1)
    epsilon = 1e-7
    entropy = -torch.sum(probs * torch.log(probs + epsilon), dim=2)
    return {"output_mean": output_mean, "predicted": predicted, "predictive_uncertainty": predictive_uncertainty,
            "aleatoric_uncertainty": entropy.mean(dim=0)}


def kl_weight(epoch: int, total_num_epochs: int) -> float:
    """train/multimodal.py:80"""
    return (2 ** (epoch + 1)) / (2 ** total_num_epochs)


def elbo_loss_multimodal(logits: torch.Tensor, labels: torch.Tensor, kl: torch.Tensor, batch_size: int,
                         epoch: int, total_num_epochs: int):
    """train/multimodal.py:121-130: CE(mean_s logits) + mean_s(KL)/batch_size*kl_weight."""
    output = torch.mean(logits, dim=0)
    scaled_kl = kl / batch_size * kl_weight(epoch, total_num_epochs)
    ce = F.cross_entropy(output, labels)
    return ce + scaled_kl, ce, scaled_kl


def elbo_loss_unimodal(logits: torch.Tensor, labels: torch.Tensor, kl: torch.Tensor, batch_size: int,
                       epoch: int, total_num_epochs: int):
    """train/unimodal.py:133-142: CE(mean_s logits) + kl_weight * mean_s(KL)/batch_size."""
    output = torch.mean(logits, dim=0)
    scaled_kl = kl / batch_size
    ce = F.cross_entropy(output, labels)
    return ce + kl_weight(epoch, total_num_epochs) * scaled_kl, ce, scaled_kl


def synthetic_batch(B: int, seed: int = 1234, size: int = 256):
    """BASELINE.md §4 / SURVEY §8d synthetic inputs: image N(0,1), bathy U[0,1), sss U[0,1), labels."""
    g = torch.Generator().manual_seed(seed)
    img = torch.randn((B, 3, size, size), generator=g)
    bathy = torch.rand((B, 3, size, size), generator=g)
    sss = torch.rand((B, 1, size, size), generator=g)
    labels = torch.randint(0, 7, (B,), generator=g)
    return img, bathy, sss, labels
