"""S-batched Monte-Carlo forward engine: the B200 execution plan behind the reference's S-pass loops

    for s in range(S): model(img, bathy, sss)      (inference/predictors.py:54-66,
                                                    train/multimodal.py:107-118, 280-285,
                                                    train/unimodal.py:127-131, 259-264)

Instead of S sequential full-network passes of ~1000 small ATen launches each, the engine walks the
network ONCE per group of G samples: for every Bayesian layer it samples G weight copies
(w = mu + softplus(rho)*eps, Philox in-kernel or injected eps) into a small fp16 operand buffer and runs
one grouped tcgen05 implicit-GEMM over all G samples; BatchNorm uses per-(sample, channel) batch
statistics accumulated in the GEMM epilogue (the reference keeps BN in train mode for every MC pass).
Activations are NHWC fp16, statistics and the fusion head fp32. All compute is libmauv_b200.so.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib, ops
from .bayesian import bayesian_layers, current_seed

F16, F32 = torch.float16, torch.float32

# Test hooks (validation only): engines built by the drivers pick these up. DEFAULT_PRECISION selects the production
# ("fp16") or validation ("x3") arithmetic; DEBUG_EPS, when set to {layer name: {"w": [S, ...], "b": [S, out] | None}},
# replaces the in-kernel Philox eps so that a driver run can be compared with the oracle on identical noise.
DEFAULT_PRECISION = "fp16"
DEBUG_EPS = None


def eval_precision(kind: str) -> str:
    """Arithmetic of the reference's fp32 evaluation drivers (train/multimodal.py:280-310 and train/unimodal.py:259-308 run
    WITHOUT autocast). Measured against the fp32 oracle at BASELINE sizes (tests/test_gpu_golden.py): the multimodal logits
    of the fp16-operand path are within 6.4e-4 of scale (north_star: 1e-3) because the attention / fusion head damps the
    trunks' rounding noise, so `evaluate_multimodal_model` keeps the fast path; a unimodal ResNet50Custom exposes that noise
    directly (5.7e-2 of scale), so `evaluate_unimodal_model` runs the fp32-class "x3" arithmetic (8.6e-5 of scale, argmax
    exact, ~3x the tensor work). MAUV_EVAL_PRECISION=fp16|x3 overrides both; DEFAULT_PRECISION (test hook) wins if changed."""
    env = os.environ.get("MAUV_EVAL_PRECISION", "")
    if env in ("fp16", "x3"):
        return env
    if DEFAULT_PRECISION != "fp16":
        return DEFAULT_PRECISION
    return "x3" if kind == "unimodal" else "fp16"


def _one(v):
    return v[0] if isinstance(v, (tuple, list)) else v


@dataclass
class _Conv:
    name: str
    layer: nn.Module
    layer_id: int
    cin: int
    cout: int
    k: int
    stride: int
    pad: int


@dataclass
class _Block:
    conv1: _Conv
    bn1: nn.BatchNorm2d
    conv2: _Conv
    bn2: nn.BatchNorm2d
    conv3: _Conv
    bn3: nn.BatchNorm2d
    down: Optional[_Conv] = None
    down_bn: Optional[nn.BatchNorm2d] = None


@dataclass
class _Trunk:
    stem: _Conv
    stem_bn: nn.BatchNorm2d
    blocks: List[_Block] = field(default_factory=list)
    fc: Optional[tuple] = None   # (name, layer, layer_id) Bayesian Linear head of a unimodal ResNet50Custom


class MCEngine:
    """Execution plan for a (Bayesian) MultiModalModel or ResNet50Custom living on a CUDA device."""

    def __init__(self, model: nn.Module, max_group: Optional[int] = None, precision: Optional[str] = None):
        _lib.require_device()
        if isinstance(model, (nn.DataParallel, nn.parallel.DistributedDataParallel)):
            model = model.module
        self.model = model
        # samples walked together. None: as many as fit in half of the free device memory (auto_group) - every grouping gives
        # bit-identical logits, larger groups mean fewer, larger launches (cfg2: 461 / 470 / 475 triplets/s at 10 / 15 / 30)
        self.max_group = max_group
        self._auto_groups: Dict[tuple, int] = {}
        self.layer_ids: Dict[str, int] = {n: i for i, (n, _) in enumerate(bayesian_layers(model))}
        self._by_module = {id(l): n for n, l in bayesian_layers(model)}
        if hasattr(model, "image_model_feat"):
            self.kind = "multimodal"
            self.trunks = [self._plan_trunk(model.image_model_feat, "image_model_feat"),
                           self._plan_trunk(model.bathy_model_feat, "bathy_model_feat"),
                           self._plan_trunk(model.sss_model_feat, "sss_model_feat")]
            self.attn = [model.attention_image, model.attention_bathy, model.attention_sss]
        elif hasattr(model, "model") and hasattr(model.model, "layer1"):
            self.kind = "unimodal"
            self.trunks = [self._plan_trunk(model.model, "model")]
        else:
            raise _lib.MauvError("MCEngine supports the reference's MultiModalModel and ResNet50Custom topologies")
        p = next(model.parameters())
        if not p.is_cuda:
            raise _lib.MauvError("MCEngine: move the model to a CUDA device first (there is no CPU path)")
        self.device = p.device
        self.launches = 0
        # "fp16": production path (fp16 operands, fp32 accumulate). "x3": validation path - every value is an fp16
        # (hi | lo) pair and every product a 3-term split on the same tensor-core kernel (~fp32 accuracy, ~3x cost)
        precision = precision or DEFAULT_PRECISION
        if precision not in ("fp16", "x3"):
            raise _lib.MauvError("precision must be 'fp16' or 'x3'")
        self.precision = precision
        # recompute-fusion of conv3 + bn3 + residual + ReLU (removes the y3 write and re-read) for bottlenecks whose
        # conv3 has K <= fuse_conv3_max_k (layer1/layer2: HBM-write bound; deeper ones are tensor bound)
        self.fuse_conv3 = True
        self.fuse_input_bn = os.environ.get("MAUV_FUSE_INPUT_BN", "0") == "1"
        self.fuse_conv3_max_k = int(os.environ.get('MAUV_FUSE_CONV3_MAX_K', '256'))
        # statistics of the recompute scheme in closed form: sum y = w . colsum(a), sum y^2 = w^T (a^T a) w - one K x K
        # second-moment contraction over the pixels instead of the N x K statistics pass (N = 4K)
        self.gram_stats = os.environ.get("MAUV_GRAM_STATS", "1") != "0"
        # stem: conv1 + bn1 statistics + max-pool of the raw output in one kernel (ops.stem_conv_pool_f16): the full-resolution
        # conv1 output never reaches HBM. 256 x 256 inputs (Wo = 128) only; other sizes take the three-kernel path.
        self.stem_pool = os.environ.get("MAUV_STEM_POOL", "1") != "0"
        self.trunk_streams = None    # None: automatic (see _trunks_in_parallel); True / False force it
        self.stem_colsum = True      # the fused stem's bn_act pass also emits the column sums layer1.0's downsample statistics need
        # bn2 + ReLU of the recompute tails' input applied to the operand tiles in shared memory (second-moment contraction and
        # fused conv3): a2 = relu(bn2(conv2(.))) is never written to / re-read from HBM
        self.fuse_a2 = os.environ.get("MAUV_FUSE_A2", "1") != "0"
        self.fuse_a2_max_k = int(os.environ.get("MAUV_FUSE_A2_MAX_K", "128"))     # K = 256 (layer3): the transform costs more than bn_act

    # Philox sample-id cursor, shared by every engine built on the same model (it lives on the model object): each
    # forward_mc / TrainEngine.step / predictor batch that is not given explicit sample ids takes the next S ids, so
    # every batch, epoch and call sees fresh Monte-Carlo draws like the reference (eps.data.normal_() per pass) instead of
    # replaying ids [0, S).
    @property
    def _sample_cursor(self) -> int:
        return self.model.__dict__.get("_mauv_sample_cursor", 0)

    @_sample_cursor.setter
    def _sample_cursor(self, v: int) -> None:
        self.model.__dict__["_mauv_sample_cursor"] = int(v) & 0xFFFFFFFF

    def auto_group(self, inputs: Sequence[torch.Tensor], S: int) -> int:
        """Largest sample group whose live activations fit in half of the free device memory (measured: ~9.8 MB per
        (sample, triplet) at 256 x 256 in the fp16 plan - one trunk is live at a time -, 2.2x that in the x3 plan); cached per
        input geometry so that consecutive batches take the same plan."""
        B, _, H, W = inputs[0].shape
        key = (B, H, W, S, self.precision)
        g = self._auto_groups.get(key)
        if g is None:
            free, _total = torch.cuda.mem_get_info(self.device)
            free += torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device)   # the allocator's cache
            per = 9.8e6 * (H * W / 65536.0) * (2.2 if self.precision == "x3" else 1.0)
            g = max(1, min(S, int(0.5 * free / (B * per))))
            self._auto_groups[key] = g
        return g

    def _trunk_stream(self, i: int) -> torch.cuda.Stream:
        if not hasattr(self, "_trunk_streams"):
            self._trunk_streams = [torch.cuda.Stream(self.device) for _ in range(3)]
        return self._trunk_streams[i]

    def _trunks_in_parallel(self, xs, G: int) -> bool:
        """Run the three trunks on three streams? MAUV_TRUNK_STREAMS=0|1 forces it; by default when three trunks' live
        activations (3 x the sequential footprint, see auto_group) fit in half of the free memory AND the group is small
        enough for the latency-bound launches to matter (G <= 8)."""
        if self.trunk_streams is not None:
            return bool(self.trunk_streams)
        mode = os.environ.get("MAUV_TRUNK_STREAMS", "auto")
        if mode in ("0", "1"):
            return mode == "1"
        if self.precision != "fp16" or G > 8:
            return False
        B, _, H, W = xs[0].shape
        key = ("par", B, H, W, G)
        ok = self._auto_groups.get(key)
        if ok is None:
            free, _total = torch.cuda.mem_get_info(self.device)
            free += torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device)
            ok = int(3 * B * G * 9.8e6 * (H * W / 65536.0) <= 0.5 * free)
            self._auto_groups[key] = ok
        return bool(ok)

    def take_samples(self, S: int) -> int:
        """-> first id of a fresh block of S sample ids (advances the model's cursor)"""
        s0 = self._sample_cursor
        self._sample_cursor = s0 + S
        return s0

    # ------------------------------------------------------------------ planning
    def _conv(self, layer: nn.Module, name: str) -> _Conv:
        if not hasattr(layer, "mu_kernel"):
            raise _lib.MauvError(f"{name}: expected a Bayesian conv (run dnn_to_bnn first)")
        kh, kw = layer.kernel_size
        assert kh == kw
        if _one(layer.dilation) != 1 or layer.groups != 1 or layer.mu_bias is not None:
            raise _lib.MauvError(f"{name}: unsupported conv configuration")
        return _Conv(name, layer, self.layer_ids[name], layer.in_channels, layer.out_channels, kh,
                     _one(layer.stride), _one(layer.padding))

    def _plan_trunk(self, net: nn.Module, prefix: str) -> _Trunk:
        t = _Trunk(self._conv(net.conv1, f"{prefix}.conv1"), net.bn1)
        for li in range(1, 5):
            stage = getattr(net, f"layer{li}")
            for bi, b in enumerate(stage):
                pre = f"{prefix}.layer{li}.{bi}"
                blk = _Block(self._conv(b.conv1, pre + ".conv1"), b.bn1, self._conv(b.conv2, pre + ".conv2"), b.bn2,
                             self._conv(b.conv3, pre + ".conv3"), b.bn3)
                if b.downsample is not None:
                    blk.down = self._conv(b.downsample[0], pre + ".downsample.0")
                    blk.down_bn = b.downsample[1]
                t.blocks.append(blk)
        fc = getattr(net, "fc", None)
        if fc is not None and hasattr(fc, "mu_weight"):
            t.fc = (f"{prefix}.fc", fc, self.layer_ids[f"{prefix}.fc"])
        return t

    # ------------------------------------------------------------------ helpers
    def _eps_w(self, eps, name, s0, G):
        if eps is None:
            return None
        return eps[name]["w"][s0:s0 + G].to(self.device, F32, non_blocking=True).contiguous()

    def _eps_b(self, eps, name, s0, G):
        if eps is None or eps[name]["b"] is None:
            return None
        return eps[name]["b"][s0:s0 + G].to(self.device, F32, non_blocking=True).contiguous()

    def _sample(self, c: _Conv, G, s0, eps, seed):
        self.launches += 1
        return ops.sample_weights_f16(c.layer.mu_kernel.detach(), c.layer.rho_kernel.detach(), G,
                                      eps=self._eps_w(eps, c.name, s0, G), seed=seed, layer_id=c.layer_id, sample0=s0)

    def _bn(self, stats, count, bn: nn.BatchNorm2d, G):
        self.launches += 2 if stats.shape[1] > 64 else 1
        mom = 0.1 if bn.momentum is None else bn.momentum
        track = bn.track_running_stats and bn.running_mean is not None
        ss = ops.bn_finalize(stats, count, bn.weight.detach() if bn.weight is not None else None,
                             bn.bias.detach() if bn.bias is not None else None, bn.eps, mom,
                             bn.running_mean if track else None, bn.running_var if track else None,
                             num_batches_tracked=bn.num_batches_tracked if track else None)
        return ss

    def _bn_gram(self, a, colsum, w, count, bn: nn.BatchNorm2d, G, a_ss=None):
        """BN scale/shift of the 1x1 conv a @ w^T from the input's moments (ops.bn_stats_from_gram); same running-statistics
        bookkeeping as _bn. a_ss: `a` is raw and relu(a * a_ss) is what the conv reads (transform inside the kernels)."""
        self.launches += 4
        mom = 0.1 if bn.momentum is None else bn.momentum
        track = bn.track_running_stats and bn.running_mean is not None
        return ops.bn_stats_from_gram(a, colsum, w, count, bn.weight.detach() if bn.weight is not None else None,
                                      bn.bias.detach() if bn.bias is not None else None, bn.eps, mom,
                                      bn.running_mean if track else None, bn.running_var if track else None,
                                      bn.num_batches_tracked if track else None, a_ss=a_ss)

    def _conv_bn(self, c: _Conv, bn, x, G, B, s0, eps, seed):
        """x: [G*B, H, W, Cin] fp16 -> raw conv output [G*B, Ho, Wo, Cout] fp16 + BN scale/shift [G, Cout, 2]."""
        w = self._sample(c, G, s0, eps, seed)
        NB, H, W, Cin = x.shape
        self.launches += 1
        if c.k == 1 and c.stride == 1 and c.pad == 0:
            y, st = ops.gemm_f16(x.view(G, B * H * W, Cin), w, stats=True)
            y = y.view(NB, H, W, c.cout)
        else:
            y, st = ops.conv2d_im2col_f16(x, w, G, c.k, c.k, c.stride, c.pad, stats=True)
        count = y.numel() // (G * c.cout)
        return y, self._bn(st, count, bn, G)

    # ------------------------------------------------------------------ trunk, fp16x3 validation mode
    def _bn_x3(self, stats, count, bn: nn.BatchNorm2d, G):
        # the conv output carries the 2^8 weight scale: BN(256*y) with eps*2^16 == BN(y) with eps (gamma/beta unchanged);
        # running statistics are not replayed in this mode (they do not influence train-mode outputs)
        return ops.bn_finalize(stats, count, bn.weight.detach() if bn.weight is not None else None,
                               bn.bias.detach() if bn.bias is not None else None, bn.eps * ops.X3_SCALE ** 2, 0.0)

    def _conv_bn_x3(self, c: _Conv, bn, x2, G, B, s0, eps, seed):
        plain = c.k == 1 and c.stride == 1 and c.pad == 0
        w3 = ops.sample_weights_x3_f16(c.layer.mu_kernel.detach(), c.layer.rho_kernel.detach(), G, per_tap=not plain,
                                       eps=self._eps_w(eps, c.name, s0, G), seed=seed, layer_id=c.layer_id, sample0=s0)
        NB, H, W, C2 = x2.shape
        if plain:
            y2, st = ops.gemm_x3_f16(x2.view(G, B * H * W, C2), w3)
            y2 = y2.view(NB, H, W, 2 * c.cout)
        else:
            y2, st = ops.conv2d_im2col_x3_f16(x2, w3, G, c.k, c.k, c.stride, c.pad)
        count = y2.numel() // (G * 2 * c.cout)
        return y2, self._bn_x3(st, count, bn, G)

    def _run_trunk_x3(self, t: _Trunk, x_nchw: torch.Tensor, G: int, s0: int, eps, seed) -> torch.Tensor:
        B = x_nchw.shape[0]
        st = t.stem
        a0 = ops.stem_im2col_x3_f16(x_nchw, st.k, st.k, st.stride, st.pad)
        w3 = ops.sample_weights_x3_f16(st.layer.mu_kernel.detach(), st.layer.rho_kernel.detach(), G, per_tap=False,
                                       eps=self._eps_w(eps, st.name, s0, G), seed=seed, layer_id=st.layer_id, sample0=s0)
        y2, stats = ops.gemm_x3_f16(a0, w3, shared_a=True)
        Ho = (x_nchw.shape[2] + 2 * st.pad - st.k) // st.stride + 1
        Wo = (x_nchw.shape[3] + 2 * st.pad - st.k) // st.stride + 1
        ss = self._bn_x3(stats, B * Ho * Wo, t.stem_bn, G)
        x = ops.bn_relu_maxpool_x3_f16(y2.view(G * B, Ho, Wo, 2 * st.cout), ss, G)
        for blk in t.blocks:
            y1, ss1 = self._conv_bn_x3(blk.conv1, blk.bn1, x, G, B, s0, eps, seed)
            a1 = ops.bn_act_x3_f16(y1, ss1, G, blk.conv1.cout)
            y2, ss2 = self._conv_bn_x3(blk.conv2, blk.bn2, a1, G, B, s0, eps, seed)
            a2 = ops.bn_act_x3_f16(y2, ss2, G, blk.conv2.cout)
            y3, ss3 = self._conv_bn_x3(blk.conv3, blk.bn3, a2, G, B, s0, eps, seed)
            if blk.down is not None:
                yd, ssd = self._conv_bn_x3(blk.down, blk.down_bn, x, G, B, s0, eps, seed)
                x = ops.bn_act_x3_f16(y3, ss3, G, blk.conv3.cout, y2b=yd, ss2=ssd)
            else:
                x = ops.bn_act_x3_f16(y3, ss3, G, blk.conv3.cout, residual=x)
        return ops.avgpool_x3_f16(x).view(G, B, -1)

    # ------------------------------------------------------------------ trunk
    def _run_trunk(self, t: _Trunk, x_nchw: torch.Tensor, G: int, s0: int, eps, seed, a0=None) -> torch.Tensor:
        if self.precision == "x3":
            return self._run_trunk_x3(t, x_nchw, G, s0, eps, seed)
        B = x_nchw.shape[0]
        st = t.stem
        x_colsum = None          # column sums of the current block input, when its producer emitted them
        if a0 is None:
            a0 = ops.stem_im2col_f16(x_nchw, st.k, st.k, st.stride, st.pad)      # shared by all samples
        w = self._sample(st, G, s0, eps, seed)
        Ho = (x_nchw.shape[2] + 2 * st.pad - st.k) // st.stride + 1
        Wo = (x_nchw.shape[3] + 2 * st.pad - st.k) // st.stride + 1
        if self.stem_pool and Wo == 128 and Ho % 2 == 0 and st.cout == 64 and a0.shape[1] <= 192:
            # max-pool of the RAW conv output in the conv's epilogue (window min where gamma < 0), then bn1 + ReLU on the pooled
            # tensor: relu(bn(.)) is monotone per channel, so this equals maxpool(relu(bn1(conv1(x)))) bit for bit
            yp, stats = ops.stem_conv_pool_f16(a0, w, B, Ho, gamma=t.stem_bn.weight.detach() if t.stem_bn.weight is not None else None)
            ss = self._bn(stats, B * Ho * Wo, t.stem_bn, G)
            # (the column sums of the activated tensor come for free with this pass: first moment of the closed-form
            # statistics of layer1.0's downsample conv, which reads x as is)
            if self.stem_colsum:
                x, x_colsum = ops.bn_act_f16(yp, ss, G, st.cout, relu=True, out=yp, colsum=True)
            else:
                x = ops.bn_act_f16(yp, ss, G, st.cout, relu=True, out=yp)
            self.launches += 3
        else:
            y, stats = ops.gemm_f16(a0, w, stats=True, shared_a=True)            # [G, B*Ho*Wo, 64]
            self.launches += 2
            ss = self._bn(stats, B * Ho * Wo, t.stem_bn, G)
            x = ops.bn_relu_maxpool_f16(y.view(G * B, Ho, Wo, st.cout), ss, G)
            self.launches += 1
            del y
        for bi, blk in enumerate(t.blocks):
            xcs = x_colsum if bi == 0 else None
            y1, ss1 = self._conv_bn(blk.conv1, blk.bn1, x, G, B, s0, eps, seed)
            c2 = blk.conv2
            if (self.fuse_input_bn and ops.STREAM_CONV and c2.cin == 64 and c2.cout == 64 and c2.k == 3 and c2.stride == 1
                    and c2.pad == 1 and y1.shape[2] <= 254):
                # layer1: bn1 + ReLU applied to conv2's input tiles in shared memory (padded-stream kernel), a1 never reaches
                # HBM. Correct (tested) but OFF by default: with 4 transform warps the in-smem pass costs 1.75 ms per launch
                # at cfg2 against the 0.44 ms bn_act pass it replaces (measured: 79 vs 31.6 + 12 ms per step).
                w2 = self._sample(c2, G, s0, eps, seed)
                y2, st2 = ops.conv3x3_c64_f16(y1, w2, G, stats=True, in_ss=ss1)
                ss2 = self._bn(st2, y2.numel() // (G * c2.cout), blk.bn2, G)
                self.launches += 1
            else:
                a1 = ops.bn_act_f16(y1, ss1, G, blk.conv1.cout, relu=True)
                y2, ss2 = self._conv_bn(blk.conv2, blk.bn2, a1, G, B, s0, eps, seed)
            fuse_tail = self.fuse_conv3 and blk.conv3.cin <= self.fuse_conv3_max_k and (
                blk.down is None or (blk.down.cin <= self.fuse_conv3_max_k and blk.down.k == 1 and blk.down.pad == 0))
            M2 = y2.shape[0] // G * y2.shape[1] * y2.shape[2]
            gram = fuse_tail and self.gram_stats and ops.gram_splits(M2, G, blk.conv2.cout) > 0
            cs2 = None
            if (gram and self.fuse_a2 and blk.conv2.cout in (64, 128, 256) and blk.conv2.cout <= self.fuse_a2_max_k
                    and blk.conv3.cout >= 128):
                # a2 = relu(bn2(y2)) only ever exists as operand tiles in shared memory: the second-moment contraction (which
                # then also yields the column sums) and the fused conv3 both read the RAW y2 and transform it on the fly
                c3 = blk.conv3
                NB, H, W, Cm = y2.shape
                if blk.down is not None:
                    x = self._fused_downsample_tail(blk, y2, x, G, B, s0, eps, seed, None, a_ss=ss2, x_colsum=xcs)
                    continue
                w3 = self._sample(c3, G, s0, eps, seed)
                y2v = y2.view(G, B * H * W, Cm)
                ss3 = self._bn_gram(y2v, None, w3, B * H * W, blk.bn3, G, a_ss=ss2)
                x = ops.gemm_bn_act_f16(y2v, w3, ss3, residual=x.view(G, B * H * W, c3.cout), relu=True,
                                        a_ss=ss2).view(NB, H, W, c3.cout)
                self.launches += 1
                continue
            if gram:    # the activation pass also emits the column sums of a2 (first moment of conv3's closed-form statistics)
                a2, cs2 = ops.bn_act_f16(y2, ss2, G, blk.conv2.cout, relu=True, colsum=True)
            else:
                a2 = ops.bn_act_f16(y2, ss2, G, blk.conv2.cout, relu=True)
            if blk.down is None and fuse_tail:
                # HBM-write-bound tail of the bottleneck: recompute scheme. Pass 1 = statistics of conv3 only (closed form
                # from a2's second moments, or the statistics-only contraction), pass 2 = conv3 with BN-apply + residual +
                # ReLU in the epilogue; y3 never reaches HBM.
                c3 = blk.conv3
                w3 = self._sample(c3, G, s0, eps, seed)
                NB, H, W, Cm = a2.shape
                a2v = a2.view(G, B * H * W, Cm)
                if gram:
                    ss3 = self._bn_gram(a2v, cs2, w3, B * H * W, blk.bn3, G)
                else:
                    ss3 = self._bn(ops.gemm_stats_f16(a2v, w3), B * H * W, blk.bn3, G)
                x = ops.gemm_bn_act_f16(a2v, w3, ss3, residual=x.view(G, B * H * W, c3.cout), relu=True).view(NB, H, W, c3.cout)
                self.launches += 2
                continue
            if blk.down is not None and fuse_tail:
                x = self._fused_downsample_tail(blk, a2, x, G, B, s0, eps, seed, cs2, x_colsum=xcs)
                continue
            y3, ss3 = self._conv_bn(blk.conv3, blk.bn3, a2, G, B, s0, eps, seed)
            if blk.down is not None:
                yd, ssd = self._conv_bn(blk.down, blk.down_bn, x, G, B, s0, eps, seed)
                x = ops.bn_act_f16(y3, ss3, G, blk.conv3.cout, y2=yd, ss2=ssd, relu=True)
            else:
                x = ops.bn_act_f16(y3, ss3, G, blk.conv3.cout, residual=x, relu=True)
            self.launches += 3
        feat = ops.avgpool_f16(x)                                                 # [G*B, 2048] fp32
        self.launches += 1
        return feat.view(G, B, -1)

    def _fused_downsample_tail(self, blk: _Block, a2, x, G, B, s0, eps, seed, cs2=None, a_ss=None, x_colsum=None):
        """relu(bn3(conv3(a2)) + bn_d(conv_d(x))) without either raw conv output in HBM: two statistics passes (recompute
        scheme), then ONE contraction over K-concatenated operands [a2 | x'] * [s3*W3 | sd*Wd]^T + (t3 + td) - the BN scales
        folded into freshly sampled weights, the shifts into the epilogue. x' = x for stride 1, else x subsampled."""
        c3, cd = blk.conv3, blk.down
        NB, H, W, Cm = a2.shape
        M = B * H * W
        a2v = a2.view(G, M, Cm)
        xs = x if cd.stride == 1 else ops.subsample_f16(x, cd.stride)
        assert xs.shape[1] == H and xs.shape[2] == W
        xv = xs.view(G, M, cd.cin)
        w3 = self._sample(c3, G, s0, eps, seed)
        wd = self._sample(cd, G, s0, eps, seed)
        if a_ss is not None:         # a2 is the raw conv2 output, bn2 + ReLU happen on the operand tiles (see _run_trunk)
            ss3 = self._bn_gram(a2v, None, w3, M, blk.bn3, G, a_ss=a_ss)
        elif cs2 is not None:
            ss3 = self._bn_gram(a2v, cs2, w3, M, blk.bn3, G)
        else:
            ss3 = self._bn(ops.gemm_stats_f16(a2v, w3), M, blk.bn3, G)
        if self.gram_stats and ops.gram_splits(M, G, cd.cin) > 0:
            csx = x_colsum if (x_colsum is not None and cd.stride == 1) else ops.colsum_f16(xv, G, cd.cin)
            ssd = self._bn_gram(xv, csx, wd, M, blk.down_bn, G)
        else:
            ssd = self._bn(ops.gemm_stats_f16(xv, wd), M, blk.down_bn, G)
        wcat = torch.empty((G, c3.cout, Cm + cd.cin), dtype=F16, device=self.device)
        ops.sample_weights_scaled_f16(c3.layer.mu_kernel.detach(), c3.layer.rho_kernel.detach(), G, ss3, wcat, 0,
                                      eps=self._eps_w(eps, c3.name, s0, G), seed=seed, layer_id=c3.layer_id, sample0=s0)
        ops.sample_weights_scaled_f16(cd.layer.mu_kernel.detach(), cd.layer.rho_kernel.detach(), G, ssd, wcat, Cm,
                                      eps=self._eps_w(eps, cd.name, s0, G), seed=seed, layer_id=cd.layer_id, sample0=s0)
        out = ops.gemm_bn_cat_f16(a2v, xv, wcat, ops.bn_shift_sum(ss3, ssd), relu=True, a1_ss=a_ss)
        self.launches += 9 + (cd.stride != 1)
        return out.view(NB, H, W, c3.cout)

    # ------------------------------------------------------------------ head
    def _linear(self, layer, name, x, G, s0, eps, seed, out=None, out_col=0):
        self.launches += 1
        return ops.sampled_linear_f32(
            x, layer.mu_weight.detach(), layer.rho_weight.detach(),
            layer.mu_bias.detach() if layer.mu_bias is not None else None,
            layer.rho_bias.detach() if layer.rho_bias is not None else None,
            eps_w=self._eps_w(eps, name, s0, G), eps_b=self._eps_b(eps, name, s0, G), seed=seed,
            layer_id=self.layer_ids[name], sample0=s0, out=out, out_col=out_col)

    def _attention(self, attn, prefix, feat, G, s0, eps, seed, concat, col):
        k = self._linear(attn.key_projection, prefix + ".key_projection", feat, G, s0, eps, seed)
        v = self._linear(attn.value_projection, prefix + ".value_projection", feat, G, s0, eps, seed)
        q = self._linear(attn.query_projection, prefix + ".query_projection", feat, G, s0, eps, seed)
        t = ops.tanh_add_f32(q, k)
        sc = self._linear(attn.attention_mechanism, prefix + ".attention_mechanism", t, G, s0, eps, seed)
        ops.softmax_gate_f32(sc, v, concat, col)
        self.launches += 2

    # ------------------------------------------------------------------ public
    @torch.no_grad()
    def stem_matrices(self, inputs: Sequence[torch.Tensor]):
        """Explicit im2col matrices of the three 7x7/2 stems: a function of the input batch only, so one build
        serves every MC sample (and every sample group) of the batch."""
        xs = [x.to(self.device, F32).contiguous() for x in inputs]
        return [ops.stem_im2col_f16(x, t.stem.k, t.stem.k, t.stem.stride, t.stem.pad) for t, x in zip(self.trunks, xs)]

    @torch.no_grad()
    def forward_group(self, inputs: Sequence[torch.Tensor], G: int, sample0: int = 0, eps: Optional[dict] = None,
                      seed: Optional[int] = None, stems=None) -> torch.Tensor:
        """One walk of the network for MC samples [sample0, sample0+G) -> logits [G, B, C] fp32."""
        seed = current_seed() if seed is None else seed
        xs = [x.to(self.device, F32).contiguous() for x in inputs]
        B = xs[0].shape[0]
        if self.kind == "unimodal":
            t = self.trunks[0]
            feat = self._run_trunk(t, xs[0], G, sample0, eps, seed, None if stems is None else stems[0])
            name, fc, _ = t.fc
            return self._linear(fc, name, feat, G, sample0, eps, seed)
        m = self.model
        concat = torch.empty((G, B, 384), dtype=F32, device=self.device)
        names = ("attention_image", "attention_bathy", "attention_sss")
        if self._trunks_in_parallel(xs, G):
            # The three trunks are independent until the fusion head: each runs on its own stream (forked from / joined to the
            # caller's stream with events, also inside a CUDA-graph capture), so the ~300 latency-bound launches per trunk (BN
            # finalize, closed-form evaluation, sampling, head projections) and the tails of under-filled kernels overlap with
            # another trunk's work. Same kernels, same order per trunk: bit-identical results. Needs three trunks' activations
            # live at once, so it is taken for small sample groups (the 4- and 8-GPU shards) only.
            main = torch.cuda.current_stream(self.device)
            fork = torch.cuda.Event()
            fork.record(main)
            joins = []
            for i, (t, x, attn, pre) in enumerate(zip(self.trunks, xs, self.attn, names)):
                side = self._trunk_stream(i)
                side.wait_event(fork)
                with torch.cuda.stream(side), ops.on_current_stream():
                    feat = self._run_trunk(t, x, G, sample0, eps, seed, None if stems is None else stems[i])
                    self._attention(attn, pre, feat, G, sample0, eps, seed, concat, 128 * i)
                    del feat
                    done = torch.cuda.Event()
                    done.record(side)
                joins.append(done)
            for done in joins:
                main.wait_event(done)
        else:
            for i, (t, x, attn, pre) in enumerate(zip(self.trunks, xs, self.attn, names)):
                feat = self._run_trunk(t, x, G, sample0, eps, seed, None if stems is None else stems[i])
                self._attention(attn, pre, feat, G, sample0, eps, seed, concat, 128 * i)
        h = self._linear(m.fc, "fc", concat, G, sample0, eps, seed)
        h = self._linear(m.fc1, "fc1", h, G, sample0, eps, seed)
        return self._linear(m.fc2, "fc2", h, G, sample0, eps, seed)

    @torch.no_grad()
    def forward_mc(self, inputs: Sequence[torch.Tensor], S: int, sample0: Optional[int] = None, eps: Optional[dict] = None,
                   seed: Optional[int] = None, group: Optional[int] = None) -> torch.Tensor:
        """logits [S, B, C] for MC samples sample0 .. sample0+S-1. sample0=None: a fresh block of ids from the model's
        cursor (production); with injected eps (validation) the ids index the eps tensors and default to 0."""
        G = min(group or self.max_group or self.auto_group(inputs, S), S)
        if eps is None:
            eps = DEBUG_EPS
        if sample0 is None:
            sample0 = 0 if eps is not None else self.take_samples(S)
        with ops.on_current_stream():
            return self._forward_mc(inputs, S, G, sample0, eps, seed)

    def _forward_mc(self, inputs, S, G, sample0, eps, seed) -> torch.Tensor:
        outs = []
        stems = self.stem_matrices(inputs) if (S > G and self.precision == "fp16") else None
        for s in range(0, S, G):
            g = min(G, S - s)
            outs.append(self.forward_group(inputs, g, sample0 + s, eps, seed, stems))
        return torch.cat(outs, dim=0)
