"""S-batched ELBO training step: the B200 execution plan behind the reference's training inner loop

    outputs = [model(x...) for _ in range(num_mc)]; output = mean(outputs)
    loss = criterion(output, labels) + kl / batch_size * kl_weight; loss.backward()
                                            (train/multimodal.py:104-138, train/unimodal.py:125-145)

The layer path (functional.py) replays that loop literally: S network walks forward and S autograd walks backward of
~2 000 small launches each. Here the S passes are walked ONCE forward (engine.MCEngine's grouped tcgen05 kernels, with
every conv input / raw output and the per-(sample, channel) BatchNorm statistics kept) and ONCE backward:

    dlogits_s = d CE(mean_s logits) / d logits_s                        mauv_ce_mean_fwd_bwd_f32
    fusion head (fp32, sampling fused, parameter grads summed over s)    mauv_sampled_linear_bwd_group_f32 ...
    per bottleneck, in reverse: ReLU mask + residual fan-in + train-mode BN backward (3 launches per site),
        dW_s = X_s^T dY_s as one grouped tcgen05 GEMM over (sample, pixel-chunk) batches  -> dmu, drho (eps replayed)
        dX_s = dY_s * W_s^T as one grouped tcgen05 GEMM / implicit-GEMM conv over the re-sampled, flipped weights
    KL and its gradient once per step                                    mauv_kl_fwd_bwd

Gradient tensors are fp16 with device-resident power-of-two scales (train_bwd.cu): no host synchronisation anywhere
in the step. Parameter gradients land in fp32 `.grad` (accumulated, like autograd), so any torch optimizer follows.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib, ops
from .backward import _splits
from .bayesian import bayesian_layers, current_seed, reference_stale_eps
from .engine import MCEngine, _Block, _Conv, _Trunk
from .flatgrad import FlatGrads

F16, F32 = torch.float16, torch.float32


def _group_splits(M: int, G: int, cout: int, kp: int) -> int:
    """Split-K factor of the grouped weight-gradient GEMM: with G samples on the batch axis only as many pixel chunks as
    needed to give every SM a couple of output tiles (the layer path, G = 1, needs up to 32)."""
    tiles = G * ((cout + 127) // 128) * ((kp + 255) // 256)
    s, cap = 1, _splits(M)
    while s < cap and tiles * s < 2 * 148:
        s *= 2
    return s


@dataclass
class _ConvRec:
    c: _Conv
    bn: nn.BatchNorm2d
    x: torch.Tensor        # conv input, NHWC fp16 [G*B, H, W, Cin] (None for the stem: explicit im2col matrix instead)
    y: torch.Tensor        # raw conv output, NHWC fp16
    bs: torch.Tensor       # batch statistics (mean, biased var) [G, Cout, 2]
    w: Optional[torch.Tensor] = None   # the forward weight samples [G, Cout, K] fp16 (re-laid-out for the data gradient)


@dataclass
class _BlockRec:
    blk: _Block
    r1: _ConvRec
    a1: torch.Tensor
    r2: _ConvRec
    a2: torch.Tensor
    r3: _ConvRec
    rd: Optional[_ConvRec]
    out: torch.Tensor


@dataclass
class _TrunkRec:
    t: _Trunk
    a0: torch.Tensor       # stem im2col matrix [B*Ho*Wo, Kp]
    stem: _ConvRec
    ss: torch.Tensor       # stem BN scale/shift (the maxpool backward recomputes relu(bn(y)))
    blocks: List[_BlockRec]
    HW: int


class TrainEngine(MCEngine):
    """forward + backward of one ELBO minibatch over S Monte-Carlo passes; gradients accumulate into `.grad`."""

    def __init__(self, model: nn.Module, max_group: int = 32):
        super().__init__(model, max_group=max_group, precision="fp16")
        frozen = [n for n, p in self.model.named_parameters() if not p.requires_grad]
        if frozen:      # raised here, before any epoch loop starts (train_engine_for then selects the autograd layer path)
            raise _lib.MauvError(f"TrainEngine computes gradients for every parameter; {len(frozen)} are frozen "
                                 f"(requires_grad=False), e.g. {frozen[0]}")
        self.direct_wgrad = os.environ.get("MAUV_DIRECT_WGRAD", "1") != "0"
        # north_star (1): the S sampled weight copies are never materialised - the forward samples of a layer are dropped as
        # soon as its contraction has been enqueued and the data-gradient operand is RE-SAMPLED from the same Philox ids in the
        # backward walk (bit-identical fp16 values; one layer's copies alive at a time). MAUV_KEEP_WEIGHT_SAMPLES=1 keeps them
        # on the tape instead (4.4 GB at cfg3) and saves the re-sampling launches (~5 % of the step).
        self.keep_weight_samples = os.environ.get("MAUV_KEEP_WEIGHT_SAMPLES", "0") == "1"
        # One ELBO step is ~2 300 C-ABI calls; Python enqueues them at ~57 us each (~130 ms), which is longer than the GPU
        # needs for cfg3 (~115 ms). The step is therefore recorded once per (shapes, S, kl_scale) as a CUDA graph and
        # replayed: inputs / labels are copied into static buffers and the Philox sample ids come from a device-resident
        # base word (ops.sample_base), so every replay draws fresh eps. The first step with a new key runs eagerly.
        self.use_graph = os.environ.get("MAUV_TRAIN_GRAPH", "1") != "0"
        self._graphs = {}
        self._update_running = True
        self._flat: Optional[FlatGrads] = None
        self._fused = {}        # id(optimizer) -> FusedAdam | None (optimizer_step)
        # activations kept per (triplet, MC sample) for the backward walk, bytes at 256x256 (raw conv outputs + activations of
        # 53 convs per trunk in fp16, plus the transient gradient tensors); scaled with the input resolution
        self.tape_bytes_256 = (230 if self.kind == "multimodal" else 80) * 2 ** 20
        self.live_samples = int(os.environ.get("MAUV_TRAIN_LIVE_SAMPLES", "0")) or None
        self._kl_plan = None
        self._bayes = [l for _, l in bayesian_layers(self.model)]
        self._sample_cursor = max([l._calls for l in self._bayes] + [self._sample_cursor])

    # ------------------------------------------------------------------ forward with tape
    def _conv_bn_rec(self, c: _Conv, bn, x, G, B, s0, eps, seed) -> tuple:
        w = self._sample(c, G, s0, eps, seed)
        NB, H, W, Cin = x.shape
        if c.k == 1 and c.stride == 1 and c.pad == 0:
            y, st = ops.gemm_f16(x.view(G, B * H * W, Cin), w, stats=True)
            y = y.view(NB, H, W, c.cout)
        else:
            y, st = ops.conv2d_im2col_f16(x, w, G, c.k, c.k, c.stride, c.pad, stats=True)
        ss, bs = self._bn_stats(st, y.numel() // (G * c.cout), bn)
        return _ConvRec(c, bn, x, y, bs, w if self.keep_weight_samples else None), ss

    def _bn_stats(self, stats, count, bn: nn.BatchNorm2d):
        if bn.weight is None or bn.bias is None:
            raise _lib.MauvError("TrainEngine needs affine BatchNorm (torchvision's default)")
        mom = 0.1 if bn.momentum is None else bn.momentum
        # the recompute pass of the two-phase step replays a forward whose running-statistics update already happened
        track = bn.track_running_stats and bn.running_mean is not None and self._update_running
        return ops.bn_finalize(stats, count, bn.weight.detach(), bn.bias.detach(), bn.eps, mom,
                               bn.running_mean if track else None, bn.running_var if track else None,
                               want_batch_stats=True, num_batches_tracked=bn.num_batches_tracked if track else None)

    def _trunk_forward(self, t: _Trunk, x_nchw, G, s0, eps, seed):
        B = x_nchw.shape[0]
        st = t.stem
        a0 = ops.stem_im2col_f16(x_nchw, st.k, st.k, st.stride, st.pad)
        w = self._sample(st, G, s0, eps, seed)
        y, stats = ops.gemm_f16(a0, w, stats=True, shared_a=True)
        Ho = (x_nchw.shape[2] + 2 * st.pad - st.k) // st.stride + 1
        Wo = (x_nchw.shape[3] + 2 * st.pad - st.k) // st.stride + 1
        ss, bs = self._bn_stats(stats, B * Ho * Wo, t.stem_bn)
        y = y.view(G * B, Ho, Wo, st.cout)
        x = ops.bn_relu_maxpool_f16(y, ss, G)
        rec = _TrunkRec(t, a0, _ConvRec(st, t.stem_bn, None, y, bs), ss, [], 0)
        for blk in t.blocks:
            r1, ss1 = self._conv_bn_rec(blk.conv1, blk.bn1, x, G, B, s0, eps, seed)
            a1 = ops.bn_act_f16(r1.y, ss1, G, blk.conv1.cout, relu=True)
            r2, ss2 = self._conv_bn_rec(blk.conv2, blk.bn2, a1, G, B, s0, eps, seed)
            a2 = ops.bn_act_f16(r2.y, ss2, G, blk.conv2.cout, relu=True)
            r3, ss3 = self._conv_bn_rec(blk.conv3, blk.bn3, a2, G, B, s0, eps, seed)
            rd = None
            if blk.down is not None:
                rd, ssd = self._conv_bn_rec(blk.down, blk.down_bn, x, G, B, s0, eps, seed)
                out = ops.bn_act_f16(r3.y, ss3, G, blk.conv3.cout, y2=rd.y, ss2=ssd, relu=True)
            else:
                out = ops.bn_act_f16(r3.y, ss3, G, blk.conv3.cout, residual=x, relu=True)
            rec.blocks.append(_BlockRec(blk, r1, a1, r2, a2, r3, rd, out))
            x = out
        rec.HW = x.shape[1] * x.shape[2]
        return rec, ops.avgpool_f16(x).view(G, B, -1)

    def _train_trunks_parallel(self) -> bool:
        if self.kind != "multimodal":
            return False
        if self.trunk_streams is not None:
            return bool(self.trunk_streams)
        return os.environ.get("MAUV_TRUNK_STREAMS", "1") != "0"

    def _lin(self, layer, name, x, G, s0, eps, seed, out=None, out_col=0):
        return self._linear(layer, name, x, G, s0, eps, seed, out=out, out_col=out_col)

    def _forward_group(self, xs, G, s0, eps, seed):
        """-> (logits [G, B, C], tape)"""
        B = xs[0].shape[0]
        if self.kind == "unimodal":
            t = self.trunks[0]
            rec, feat = self._trunk_forward(t, xs[0], G, s0, eps, seed)
            name, fc, _ = t.fc
            return self._lin(fc, name, feat, G, s0, eps, seed), {"trunks": [rec], "feat": [feat]}
        m = self.model
        concat = torch.empty((G, B, 384), dtype=F32, device=self.device)
        tape = {"trunks": [], "feat": [], "attn": [], "concat": concat}
        # The three trunks are independent until the fusion head: each runs on its own stream (forked from / joined to the
        # caller's stream; also inside the CUDA-graph capture of the step). A training step is ~2 300 launches over only
        # B x S = 240 images - most of them latency-bound or under-filled - so another trunk's kernels fill the gaps; the tape
        # holds all three trunks' activations anyway, so this costs no memory. Same kernels in the same order per trunk:
        # bit-identical losses and gradients.
        par = self._train_trunks_parallel()
        main = torch.cuda.current_stream(self.device)
        if par:
            fork = torch.cuda.Event()
            fork.record(main)
        joins = []
        for i, (t, x, attn, pre) in enumerate(zip(self.trunks, xs, self.attn,
                                                  ("attention_image", "attention_bathy", "attention_sss"))):
            side = self._trunk_stream(i) if par else main
            if par:
                side.wait_event(fork)
            with torch.cuda.stream(side), ops.on_current_stream():
                rec, feat = self._trunk_forward(t, x, G, s0, eps, seed)
                k = self._lin(attn.key_projection, pre + ".key_projection", feat, G, s0, eps, seed)
                v = self._lin(attn.value_projection, pre + ".value_projection", feat, G, s0, eps, seed)
                q = self._lin(attn.query_projection, pre + ".query_projection", feat, G, s0, eps, seed)
                th = ops.tanh_add_f32(q, k)
                sc = self._lin(attn.attention_mechanism, pre + ".attention_mechanism", th, G, s0, eps, seed)
                ops.softmax_gate_f32(sc, v, concat, 128 * i)
                if par:
                    done = torch.cuda.Event()
                    done.record(side)
                    joins.append(done)
            tape["trunks"].append(rec)
            tape["feat"].append(feat)
            tape["attn"].append((th, sc, v))
        for done in joins:
            main.wait_event(done)
        h1 = self._lin(m.fc, "fc", concat, G, s0, eps, seed)
        h2 = self._lin(m.fc1, "fc1", h1, G, s0, eps, seed)
        tape["h1"], tape["h2"] = h1, h2
        return self._lin(m.fc2, "fc2", h2, G, s0, eps, seed), tape

    # ------------------------------------------------------------------ backward pieces
    def _conv_backward(self, r: _ConvRec, dy, s_dy, G, s0, eps, seed, stale, need_dx=True):
        """dy [G*B, Ho, Wo, Cout] fp16 at device scale s_dy -> dX (same scale) ; accumulates dmu / drho."""
        c = r.c
        layer = c.layer
        mu, rho = layer.mu_kernel.detach(), layer.rho_kernel.detach()
        NB, Ho, Wo, Cout = dy.shape
        _, H, W, Cin = r.x.shape
        M = (NB // G) * Ho * Wo
        if M % 8 != 0:
            raise _lib.MauvError(f"conv backward needs B*Ho*Wo to be a multiple of 8 (got {M}) at {c.name}")
        splits = _group_splits(M, G, Cout, c.k * c.k * Cin)
        if (M // splits) % 64 == 0 and Cin % 64 == 0 and self.direct_wgrad:
            # both operands MN-major straight from the NHWC tensors (TMA boxes of [64 pixels][64 channels])
            dw = ops.wgrad_f16(dy, r.x, G, splits, c.k, c.k, c.stride, c.pad)                 # [G*splits, Cout, K]
        else:
            a_t = ops.transpose_chunks_f16(dy.view(G * M, Cout), G * splits)                  # [G*splits, Cout, Mc]
            if c.k == 1 and c.stride == 1:
                b_t = ops.transpose_chunks_f16(r.x.view(G * M, Cin), G * splits)              # [G*splits, Cin, Mc]
            else:
                b_t = ops.im2col_t_f16(r.x, c.k, c.k, c.stride, c.pad, G * splits)           # [G*splits, Kp, Mc]
            dw, _ = ops.gemm_f16(a_t, b_t)                                                    # [G*splits, Cout, Kp]
            del a_t, b_t
        ew = self._eps_w(eps, c.name, s0, G)
        ops.wgrad_finalize_group(dw, G, tuple(mu.shape), 1.0, s_dy, rho, layer.mu_kernel.grad, layer.rho_kernel.grad,
                                 eps=ew, seed=seed, layer_id=c.layer_id, sample0=s0, stale=stale)
        if not need_dx:
            return None
        # forward-layout sample (kept on the tape, or re-sampled from the same Philox ids / injected eps: coalesced writes),
        # re-laid-out for the data-gradient conv; sampling straight into the transposed layout is 3x slower (scattered stores)
        w_fwd = r.w if r.w is not None else self._sample(c, G, s0, eps, seed)
        wd = ops.weights_to_dgrad_f16(w_fwd, Cin, c.k, c.k)
        del w_fwd
        Hd, Wd = H + 2 * c.pad - c.k + 1, W + 2 * c.pad - c.k + 1
        dyd = dy if c.stride == 1 else ops.dilate_f16(dy, Hd, Wd, c.stride)
        if c.k == 1:
            dx, _ = ops.gemm_f16(dyd.view(G, (NB // G) * Hd * Wd, Cout), wd)
            return dx.view(NB, Hd, Wd, Cin)
        dx, _ = ops.conv2d_im2col_f16(dyd, wd, G, c.k, c.k, 1, c.k - 1 - c.pad)
        return dx

    def _stem_backward(self, tr: _TrunkRec, dy, s_dy, G, s0, eps, seed, stale):
        c = tr.stem.c
        layer = c.layer
        NB, Ho, Wo, Cout = dy.shape
        M = (NB // G) * Ho * Wo
        Kp = tr.a0.shape[1]
        splits = _group_splits(M, G, Cout, Kp)
        b_t = ops.transpose_chunks_f16(tr.a0, splits)                                         # [splits, Kp, Mc] (all samples)
        a_t = ops.transpose_chunks_f16(dy.view(G * M, Cout), G * splits)                      # [G*splits, Cout, Mc]
        dw = ops.gemm_wmod_f16(a_t, b_t, out_f32=True)                                        # batch (g, sp) uses b_t[sp]; fp32 sums
        ops.wgrad_finalize_group(dw.view(G * splits, Cout, Kp), G, tuple(layer.mu_kernel.shape), 1.0, s_dy,
                                 layer.rho_kernel.detach(), layer.mu_kernel.grad, layer.rho_kernel.grad,
                                 eps=self._eps_w(eps, c.name, s0, G), seed=seed, layer_id=c.layer_id, sample0=s0, stale=stale)

    def _site(self, gs, r: _ConvRec, d, s, G, **kw):
        bn = r.bn
        return ops.bn_bwd_site(gs, d, s, r.y, r.bs, bn.weight.detach(), bn.eps, bn.weight.grad, bn.bias.grad, G, r.c.cout, **kw)

    def _trunk_backward(self, tr: _TrunkRec, dfeat, G, s0, eps, seed, stale, gs):
        """dfeat [G*B, 2048] fp32 = d loss / d pooled features."""
        d, s = ops.avgpool_bwd_f16(gs, dfeat, tr.HW)
        last = tr.blocks[-1].out
        d = d.view(last.shape)
        d2 = s2 = None
        for br in reversed(tr.blocks):
            if br.rd is not None:
                bnd = br.rd.bn
                dy3, s3, dyd, sd, _ = self._site(gs, br.r3, d, s, G, d2=d2, s2=s2, relu_out=br.out, y2=br.rd.y,
                                                 batch_stats2=br.rd.bs, gamma2=bnd.weight.detach(), bn_eps2=bnd.eps,
                                                 grad_gamma2=bnd.weight.grad, grad_beta2=bnd.bias.grad)
                dz = None
            else:
                dy3, s3, _, _, dz = self._site(gs, br.r3, d, s, G, d2=d2, s2=s2, relu_out=br.out, want_dz=True)
            s_dz = s
            da2 = self._conv_backward(br.r3, dy3, s3, G, s0, eps, seed, stale)
            del dy3
            dy2, s2_, _, _, _ = self._site(gs, br.r2, da2, s3, G, relu_out=br.a2)
            del da2
            da1 = self._conv_backward(br.r2, dy2, s2_, G, s0, eps, seed, stale)
            del dy2
            dy1, s1_, _, _, _ = self._site(gs, br.r1, da1, s2_, G, relu_out=br.a1)
            del da1
            dx1 = self._conv_backward(br.r1, dy1, s1_, G, s0, eps, seed, stale)
            del dy1
            if br.rd is not None:
                dxd = self._conv_backward(br.rd, dyd, sd, G, s0, eps, seed, stale)
                d, s, d2, s2 = dx1, s1_, dxd, sd
            else:
                d, s, d2, s2 = dx1, s1_, dz, s_dz
        y = tr.stem.y
        dz = ops.maxpool_bwd_f16(y, tr.ss, d, s, G, d2=d2, s2=s2)
        dy, s_dy, _, _, _ = self._site(gs, tr.stem, dz, s, G)
        del dz
        self._stem_backward(tr, dy, s_dy, G, s0, eps, seed, stale)

    def _lin_bwd(self, layer, name, x, gy, G, s0, eps, seed, stale, gx=None, accumulate=False, need_gx=True):
        has_b = layer.mu_bias is not None
        return ops.sampled_linear_bwd_group_f32(
            x, gy, layer.mu_weight.detach(), layer.rho_weight.detach(), layer.rho_bias.detach() if has_b else None,
            layer.mu_weight.grad, layer.rho_weight.grad, layer.mu_bias.grad if has_b else None,
            layer.rho_bias.grad if has_b else None, eps_w=self._eps_w(eps, name, s0, G), eps_b=self._eps_b(eps, name, s0, G),
            seed=seed, layer_id=self.layer_ids[name], sample0=s0, stale=stale, gx=gx, accumulate=accumulate, need_gx=need_gx)

    def _backward_group(self, tape, dlogits, G, s0, eps, seed, stale):
        gs = ops.GradScratch(self.device)
        B = dlogits.shape[1]
        if self.kind == "unimodal":
            name, fc, _ = self.trunks[0].fc
            dfeat = self._lin_bwd(fc, name, tape["feat"][0], dlogits, G, s0, eps, seed, stale)
            self._trunk_backward(tape["trunks"][0], dfeat.view(G * B, -1), G, s0, eps, seed, stale, gs)
            return
        m = self.model
        dh2 = self._lin_bwd(m.fc2, "fc2", tape["h2"], dlogits, G, s0, eps, seed, stale)
        dh1 = self._lin_bwd(m.fc1, "fc1", tape["h1"], dh2, G, s0, eps, seed, stale)
        dcat = self._lin_bwd(m.fc, "fc", tape["concat"], dh1, G, s0, eps, seed, stale)
        par = self._train_trunks_parallel()       # each trunk's backward on the stream its forward ran on (see _forward_group)
        main = torch.cuda.current_stream(self.device)
        if par:
            fork = torch.cuda.Event()
            fork.record(main)
        joins = []
        for i, (attn, pre) in enumerate(zip(self.attn, ("attention_image", "attention_bathy", "attention_sss"))):
            side = self._trunk_stream(i) if par else main
            if par:
                side.wait_event(fork)
            with torch.cuda.stream(side), ops.on_current_stream():
                th, sc, v = tape["attn"][i]
                feat = tape["feat"][i]
                dsc, dv = ops.softmax_gate_bwd_f32(sc, v, dcat[:, :, 128 * i:128 * (i + 1)])
                dt = self._lin_bwd(attn.attention_mechanism, pre + ".attention_mechanism", th, dsc, G, s0, eps, seed, stale)
                dqk = ops.tanh_bwd_f32(th, dt)
                dfeat = self._lin_bwd(attn.key_projection, pre + ".key_projection", feat, dqk, G, s0, eps, seed, stale)
                self._lin_bwd(attn.query_projection, pre + ".query_projection", feat, dqk, G, s0, eps, seed, stale,
                              gx=dfeat, accumulate=True)
                self._lin_bwd(attn.value_projection, pre + ".value_projection", feat, dv, G, s0, eps, seed, stale,
                              gx=dfeat, accumulate=True)
                self._trunk_backward(tape["trunks"][i], dfeat.view(G * B, -1), G, s0, eps, seed, stale, gs)
                tape["trunks"][i] = None          # free this trunk's activations
                del th, sc, v, feat, dsc, dv, dt, dqk, dfeat
                if par:
                    done = torch.cuda.Event()
                    done.record(side)
                    joins.append(done)
        for done in joins:
            main.wait_event(done)

    # ------------------------------------------------------------------ public
    def flatten_grads(self) -> torch.Tensor:
        """Re-home every `.grad` into ONE contiguous fp32 buffer (flatgrad.FlatGrads): one memset, one finite check and
        (N > 1) one NCCL all-reduce per step."""
        self._flat = FlatGrads(self.model.parameters(), self.device)
        return self._flat.flat

    def zero_grad(self) -> None:
        if self._flat is not None:
            self._flat.zero()
        else:
            for p in self.model.parameters():
                if p.grad is not None:
                    p.grad.zero_()

    def allreduce_grads(self, group=None) -> None:
        """Data-parallel exchange (minibatch split over ranks, all S samples on every rank): gradients are
        averaged over the ranks, one all-reduce over the flat buffer."""
        if self._flat is None:
            self.flatten_grads()
        self._flat.all_reduce_mean(group)

    def grads_finite(self) -> torch.Tensor:
        """0-d bool tensor (device): the reference's per-parameter NaN/Inf guard (train/multimodal.py:141-145) in one pass."""
        if self._flat is not None:
            return self._flat.finite()
        return torch.stack([torch.isfinite(p.grad).all() for p in self.model.parameters() if p.grad is not None]).all()

    def optimizer_step(self, optimizer) -> torch.Tensor:
        """Guarded optimizer step (reference train/multimodal.py:141-145): the update is applied iff every gradient is
        finite. A plain torch.optim.Adam over this model's parameters is adopted by mauv.optim.FusedAdam (guard + update
        in one pass over flat buffers, decided on the device); any other optimizer keeps its own step() behind the
        one-pass device guard. -> 0-d device tensor, non-zero iff the step was applied."""
        from .optim import FusedAdam
        if self._flat is None:
            self.flatten_grads()
        key = id(optimizer)
        if key not in self._fused:
            self._fused[key] = FusedAdam.adopt(optimizer, self._flat)
            self._kl_plan = None        # adoption moved the parameters into a flat buffer: drop cached device pointers
        fa = self._fused[key]
        if fa is not None:
            return fa.step()
        ok = self.grads_finite()
        if bool(ok):
            optimizer.step()
        return ok

    def _ensure_grads(self):
        if self._flat is not None:
            self._flat.ensure_attached()
            return
        for p in self.model.parameters():
            if p.requires_grad and p.grad is None:
                p.grad = torch.zeros_like(p)

    def _kl(self, grad_scale: Optional[float]) -> torch.Tensor:
        if self._kl_plan is None:
            pairs, priors = [], set()
            for l in self._bayes:
                mu, rho = l._weight_params()
                pairs.append((mu, rho))
                if l.mu_bias is not None:
                    pairs.append((l.mu_bias, l.rho_bias))
                priors.add((float(l.prior_mean), float(l.prior_variance)))
            if len(priors) != 1:
                raise _lib.MauvError("TrainEngine: all Bayesian layers must share one prior")
            self._prior = priors.pop()
            self._kl_plan = ops.KlPlan([(m.detach(), r.detach()) for m, r in pairs], self.device)
            self._kl_pairs = pairs
        self._kl_plan.grads = [(m.grad, r.grad) for m, r in self._kl_pairs] if grad_scale is not None else None
        return self._kl_plan.run(self._prior[0], self._prior[1], grad_scale=grad_scale)

    @torch.no_grad()
    def step(self, inputs: Sequence[torch.Tensor], labels: torch.Tensor, S: int, kl_scale: float, *,
             sample0: Optional[int] = None, eps: Optional[dict] = None, seed: Optional[int] = None) -> dict:
        """One ELBO minibatch: loss = CE(mean_s logits_s, labels) + kl_scale * KL; gradients ACCUMULATE into `.grad`.
        -> {"loss", "ce", "kl" (unscaled), "mean_logit" [B, C]} (device tensors; nothing is synchronised)."""
        seed = current_seed() if seed is None else seed
        stale = reference_stale_eps()
        if eps is None:
            from . import engine as _engine
            eps = _engine.DEBUG_EPS               # test hook: injected eps instead of Philox (see engine.py)
        if sample0 is None:                   # fresh Philox sample ids every step, in step with the layers' own counters
            if eps is not None:
                sample0 = 0                   # injected eps is indexed by sample id
            else:
                sample0 = self.take_samples(S)
                for l in self._bayes:
                    l._calls += S
        xs = [x.to(self.device, F32).contiguous() for x in inputs]
        labels = labels.to(self.device, torch.int64).contiguous()
        B = xs[0].shape[0]
        live = self.live_samples
        if live is None:        # how many (triplet, sample) tapes fit in 70 % of the memory that is free right now
            free = torch.cuda.mem_get_info(self.device)[0] + torch.cuda.memory_reserved(self.device) - \
                torch.cuda.memory_allocated(self.device)
            per = self.tape_bytes_256 * xs[0].shape[2] * xs[0].shape[3] / 65536.0
            live = max(1, int(0.7 * free / per))
        G = min(self.max_group, S)
        recompute = B * S > live
        if recompute:
            G = max(1, min(G, live // B))
        if stale and G < S:
            raise _lib.MauvError("reference stale-eps mode needs all S samples in one group (raise max_group / free memory)")
        self._ensure_grads()
        if self.use_graph and eps is None and not recompute:
            return self._step_graphed(xs, labels, S, G, kl_scale, sample0, seed, stale)
        with ops.on_current_stream():
            if recompute:
                return self._step_recompute(xs, labels, S, G, kl_scale, sample0, eps, seed, stale)
            return self._step(xs, labels, S, G, kl_scale, sample0, eps, seed, stale)

    def _step_graphed(self, xs, labels, S, G, kl_scale, sample0, seed, stale) -> dict:
        """Replay of the recorded step (see __init__). Returned tensors are the graph's static outputs: valid until the
        next step with the same key (the scalars are cloned)."""
        import logging
        p0 = next(self.model.parameters())
        key = (tuple(tuple(x.shape) for x in xs), S, G, float(kl_scale), seed, stale, p0.data_ptr(),
               p0.grad.data_ptr() if p0.grad is not None else 0)
        ent = self._graphs.get(key)
        if ent is None:                      # first step with this key: eager (also runs every lazy initialisation)
            self._graphs[key] = {}
            with ops.on_current_stream():
                return self._step(xs, labels, S, G, kl_scale, sample0, None, seed, stale)
        if "graph" not in ent:
            try:
                self._graphs = {key: ent}    # one recorded step at a time: a graph pins its tape (tens of GB) in a private pool
                torch.cuda.empty_cache()
                ent["x"] = [torch.empty_like(x) for x in xs]
                ent["labels"] = torch.empty_like(labels)
                ent["base"] = torch.zeros(1, dtype=torch.int32, device=self.device)
                torch.cuda.synchronize(self.device)
                graph = torch.cuda.CUDAGraph()
                n0 = ops.launch_count
                with torch.cuda.graph(graph), ops.sample_base(ent["base"]), ops.on_current_stream():
                    ent["out"] = self._step(ent["x"], ent["labels"], S, G, kl_scale, 0, None, seed, stale)
                ent["launches"] = ops.launch_count - n0
                ent["graph"] = graph
            except Exception as e:           # still the CUDA path: the step simply stays Python-driven
                logging.warning(f"TrainEngine: CUDA-graph capture of the training step failed ({e}); running eagerly")
                self.use_graph = False
                self._graphs = {}
                with ops.on_current_stream():
                    return self._step(xs, labels, S, G, kl_scale, sample0, None, seed, stale)
        for d, x in zip(ent["x"], xs):
            d.copy_(x, non_blocking=True)
        ent["labels"].copy_(labels, non_blocking=True)
        ent["base"].fill_(((sample0 & 0xFFFFFFFF) ^ 0x80000000) - 0x80000000)       # uint32 bit pattern in an int32 word
        ent["graph"].replay()
        ops.launch_count += ent["launches"]
        o = ent["out"]
        return {"loss": o["loss"].clone(), "ce": o["ce"].clone(), "kl": o["kl"].clone(), "mean_logit": o["mean_logit"],
                "logits": o["logits"]}

    def _step_recompute(self, xs, labels, S, G, kl_scale, sample0, eps, seed, stale) -> dict:
        """Memory-bounded variant: the tapes of all S samples do not fit, so phase 1 walks forward for the logits only and
        phase 2 replays the forward of one sample group at a time (same Philox ids -> identical tensors; no second update
        of the BN running statistics) right before that group's backward walk. Costs one extra forward (~+25 %)."""
        logits = []
        for s in range(0, S, G):
            g = min(G, S - s)
            lg, tape = self._forward_group(xs, g, sample0 + s, eps, seed)
            del tape
            logits.append(lg)
        logits = logits[0] if len(logits) == 1 else torch.cat(logits, dim=0)
        ce, mean_logit, dlogits = ops.ce_mean_fwd_bwd_f32(logits, labels)
        self._update_running = False
        try:
            for s in range(0, S, G):
                g = min(G, S - s)
                _, tape = self._forward_group(xs, g, sample0 + s, eps, seed)
                self._backward_group(tape, dlogits[s:s + g], g, sample0 + s, eps, seed, stale)
                del tape
        finally:
            self._update_running = True
        kl = self._kl(kl_scale)
        return {"loss": ce + kl * kl_scale, "ce": ce, "kl": kl, "mean_logit": mean_logit, "logits": logits}

    def _step(self, xs, labels, S, G, kl_scale, sample0, eps, seed, stale) -> dict:
        tapes, logits = [], []
        for s in range(0, S, G):
            g = min(G, S - s)
            lg, tape = self._forward_group(xs, g, sample0 + s, eps, seed)
            logits.append(lg)
            tapes.append((tape, g, s))
        logits = logits[0] if len(logits) == 1 else torch.cat(logits, dim=0)
        ce, mean_logit, dlogits = ops.ce_mean_fwd_bwd_f32(logits, labels)
        for tape, g, s in tapes:
            self._backward_group(tape, dlogits[s:s + g], g, sample0 + s, eps, seed, stale)
        tapes.clear()
        kl = self._kl(kl_scale)
        return {"loss": ce + kl * kl_scale, "ce": ce, "kl": kl, "mean_logit": mean_logit, "logits": logits}
