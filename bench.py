#!/usr/bin/env python
"""Benchmark of the Multimodal-AUV Monte-Carlo Bayesian inference hot path on B200.

Metric (BASELINE.json): MC-sampled patch-triplets/sec at S=30, cfg2 = multimodal BNN inference,
batch 256 synthetic (image, bathymetry, side-scan) triplets of 256x256, 7 classes, random-init + MOPED
weights, with the entropy / mutual-information uncertainty output. One "step" = one batch through
S=30 MC passes + the MC statistics. At N>1 the 30 samples are block-partitioned over the ranks
(disjoint Philox sample ids), logits all-gathered over NCCL; total work is fixed -> "scaling": "strong".

  python bench.py [--gpus N --steps K --warmup W]              our CUDA path
  python bench.py --impl reference [...]                        the reference's CPU path (oracle port) on host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "multimodal-auv_b200"))

import torch  # noqa: E402

B_FULL, S_FULL, C_CLASSES, SIZE = 256, 30, 7, 256
GFLOP_PER_TRIPLET_SAMPLE = 31.830       # SURVEY.md §8d / BASELINE.md §3 (159 conv + 15 linear, forward)
CONV_GFLOP_PER_TRIPLET_SAMPLE = 31.824  # the tcgen05 kernel's share (head = 0.0059)
METRIC = "MC-sampled patch-triplets/sec (S=30)"
UNIT = "triplets/s"


def load_traffic():
    """Average DRAM bytes per launch of the tcgen05 kernels from the committed ncu launch list (profiles/traffic.json)."""
    p = ROOT / "profiles" / "traffic.json"
    try:
        return float(json.loads(p.read_text())["tcgen05_dram_bytes_per_launch"])
    except Exception:
        return None


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured (MEASURED_PEAKS.json, sustained)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx = [], set(), None
        for r in rows:
            try:
                r = [x.strip() for x in r]
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if sm:
            sm.sort()
            # median over samples taken under load (upper half: idle samples before/after are dropped)
            load = sm[len(sm) // 2:]
            out.update(sm_mhz=load[len(load) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------- model
MODEL_KIND = "multimodal"      # --model: cfg4 benches one unimodal ResNet50Custom branch ("image" | "bathy" | "sss")
GFLOP_PER_SAMPLE = {"multimodal": CONV_GFLOP_PER_TRIPLET_SAMPLE, "image": 10.677, "bathy": 10.677, "sss": 10.471}


def build_model_cpu(seed: int = 1234):
    """Random-init (torchvision default init) + MOPED(delta=0.1) multimodal BNN, as BASELINE.md §4."""
    from mauv.bayesian import dnn_to_bnn
    from mauv.models.base_models import MultiModalModel, ResNet50Custom
    from mauv.models.model_utils import load_pretrained_resnet_as_feature_extractor as feat
    import logging
    logging.disable(logging.WARNING)
    torch.manual_seed(seed)
    if MODEL_KIND == "multimodal":
        m = MultiModalModel(feat(), feat(), feat(input_channels=1), C_CLASSES)
    else:
        m = ResNet50Custom(1 if MODEL_KIND == "sss" else 3, C_CLASSES)
    prior = {"prior_mu": 0.0, "prior_sigma": 1.0, "posterior_mu_init": 0.0, "posterior_rho_init": -3.0,
             "type": "Reparameterization", "moped_enable": True, "moped_delta": 0.1}
    dnn_to_bnn(m, prior)
    logging.disable(logging.NOTSET)
    return m


def synthetic_inputs(B: int, seed: int = 1234, pin: bool = False):
    g = torch.Generator().manual_seed(seed)
    xs = [torch.randn((B, 3, SIZE, SIZE), generator=g), torch.rand((B, 3, SIZE, SIZE), generator=g),
          torch.rand((B, 1, SIZE, SIZE), generator=g)]
    if MODEL_KIND != "multimodal":
        xs = [xs[("image", "bathy", "sss").index(MODEL_KIND)]]
    return [x.pin_memory() for x in xs] if pin else xs


# ----------------------------------------------------------------------------------------- CPU arms
def cpu_reference_pass(n_threads: int, B: int, passes: int, autocast: bool):
    """Seconds per MC pass of the oracle port (== the reference's math) on the host cores."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import bnn_oracle as O
    torch.set_num_threads(n_threads)
    model = O.define_models(C_CLASSES, unimodal=False)["multimodal_model"].train()
    img, bathy, sss, _ = O.synthetic_batch(B, size=SIZE)
    times = []
    with torch.no_grad():
        for _ in range(passes):
            t0 = time.perf_counter()
            if autocast:   # the shipped predictor: torch.amp.autocast('cpu') -> bf16 (inference/predictors.py:55)
                with torch.amp.autocast(device_type="cpu"):
                    torch.softmax(model(img, bathy, sss), dim=1)
            else:
                torch.softmax(model(img, bathy, sss), dim=1)
            times.append(time.perf_counter() - t0)
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    Bs = 4
    times = cpu_reference_pass(cores, Bs, args.warmup + args.steps, autocast=True)
    timed = times[args.warmup:]
    per_pass = sum(timed) / len(timed)
    value = Bs / (S_FULL * per_pass)     # triplets/s at S=30: each triplet needs 30 passes
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_pass * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16 (torch.amp.autocast cpu, as inference/predictors.py:55)", "data": "synthetic",
        "config": {"workload": f"cfg2 multimodal BNN inference B={B_FULL} S={S_FULL} 256x256 C=7 (bounded CPU sample)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"each step = 1 MC pass of B={Bs} triplets through the oracle port of the reference path "
                                   f"(autocast bf16, BN train mode); triplets/s = {Bs}/(30 * s_per_pass), linear in S and B"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the mauv_b200 path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dist = world > 1
    if dist:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from mauv import ops
    from mauv.inference.predictors import MCPredictor, predict_stream, shard_samples

    B, S = args.batch, args.samples
    model = build_model_cpu().cuda().train()
    lo, hi = shard_samples(S, world, rank)
    group = min(args.group, max(1, hi - lo))
    pred = MCPredictor(model, S, group=group, eps_entropy=1e-8)
    host_in = synthetic_inputs(B, pin=True)
    dev_in = [x.cuda() for x in host_in]
    torch.cuda.synchronize()

    def barrier():
        if dist:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if dist:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return t.item()

    step_dev = lambda: pred.predict_device(dev_in)      # inputs resident in HBM
    # pinned host in -> host results out through the public predictor loop (H2D of batch i+1 overlaps batch i)
    run_e2e = lambda: [r for r in predict_stream(pred, (host_in for _ in range(args.steps)))]
    for _ in range(args.warmup):
        step_dev()
    [r for r in predict_stream(pred, (host_in for _ in range(1)))]      # warm the end-to-end path too (allocator, staging)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = ops.launch_count
    ms = timed(step_dev, args.steps)
    launches = ops.launch_count - l0
    ms_e2e = timed(run_e2e, 1)
    clocks = sampler.stop() if sampler else None

    # per-kernel breakdown: one extra instrumented step, run eagerly (CUDA events around every C-ABI launch)
    pred.use_graph = False            # profile eagerly: the event pairs cannot be recorded inside a graph replay
    pred.engine.trunk_streams = False  # ... and one trunk at a time: per-launch times of overlapping streams would count waits
    ops.start_profile()
    step_dev()
    prof = ops.stop_profile()
    pred.engine.trunk_streams = None

    # per-rank view of the instrumented step (strong scaling: which terms shrink with N and which do not)
    fam_local = {}
    for k, (c, t) in prof.items():
        b = k.split("|")[0]
        fam_local[b] = fam_local.get(b, 0.0) + t
    per_rank = None
    if dist:
        names = sorted(fam_local)
        gathered = [None] * world
        torch.distributed.all_gather_object(gathered, {"samples": hi - lo, "group": group,
                                                       "kernel_ms": {k: round(fam_local[k], 3) for k in names}})
        per_rank = gathered

    # the same batch in the fp32-class "x3" arithmetic (every value an fp16 hi|lo pair, 3-term split products on the same
    # tcgen05 kernels): the mode that meets north_star's rtol 1e-3 on EVERY network incl. the unimodal ones and that
    # evaluate_unimodal_model uses by default; one timed step after one warm-up
    x3_rec = None
    if not args.no_x3 and MODEL_KIND == "multimodal":
        pred.engine.precision, g_keep, ug_keep = "x3", pred.engine.max_group, pred.use_graph
        pred.engine.max_group, pred.use_graph = min(5, g_keep), False
        step_dev()
        ms_x3 = timed(step_dev, 1)
        pred.engine.precision, pred.engine.max_group, pred.use_graph = "fp16", g_keep, ug_keep
        x3_rec = {"value": B / (ms_x3 / 1e3), "unit": UNIT, "ms_per_step": ms_x3, "steps": 1,
                  "note": "precision='x3' (fp32-class validation / unimodal-evaluation arithmetic), same batch, same S"}

    # free the inference engine's workspaces, then the cfg3 training leg (every rank takes part: DP all-reduce)
    train_rec = None
    if not args.no_train_leg and MODEL_KIND == "multimodal":
        del pred, dev_in
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        train_rec = train_leg(args, 8, S_FULL, max(3, min(args.steps, 5)), 2)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and MODEL_KIND == "multimodal":   # the CPU arm is the headline config
        cores = os.cpu_count() or 1
        t = cpu_reference_pass(cores, 4, 3, autocast=True)[1:]
        per_pass = sum(t) / len(t)
        cpu_base = {"value": 4 / (S * per_pass), "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": "2 timed MC passes (after 1 warm-up) of B=4 triplets through the oracle port (autocast bf16 "
                              "as inference/predictors.py:55); scaled linearly to S=30"}
    if rank == 0:
        hbm_peak, tf_peak, peak_src = load_peaks()
        ms_step = ms / args.steps
        value = B / (ms_step / 1e3)
        e2e_value = B / (ms_e2e / args.steps / 1e3)
        fam = {}
        for k, (c, t) in prof.items():
            b = k.split("|")[0]
            fc, ft = fam.get(b, (0, 0.0))
            fam[b] = (fc + c, ft + t)
        # all launches of the tcgen05 kernels (gemm_f16_tc_kernel and its padded-stream sibling for layer1's 3x3 conv)
        tc = ops.TCGEN05_ENTRY_POINTS     # (inference: incl. the K x K second-moment contractions of the closed-form BN statistics)
        conv_calls = sum(fam.get(k, (0, 0.0))[0] for k in tc)
        conv_ms = sum(fam.get(k, (0, 0.0))[1] for k in tc)
        if args.detail:
            rows = []
            for k, (c, t) in prof.items():
                if "|" not in k:
                    continue
                base, tag = k.split("|")
                toks = {x[0]: x[1:] for x in tag.split() if x[0] in "GMNKC" and x[1:].isdigit()}
                gf = gb = 0.0
                if all(q in toks for q in "GMNK"):
                    g_, m_, n_, k_ = (int(toks[q]) for q in "GMNK")
                    gf = 2.0 * g_ * m_ * n_ * k_ * c / 1e9
                    kin = k_ // 9 if ("3x3" in tag or "stream" in tag) else k_
                    gb = g_ * m_ * (kin + n_) * 2 * c / 1e9
                elif all(q in toks for q in "GMC"):
                    g_, m_, c_ = (int(toks[q]) for q in "GMC")
                    nbuf = 2 + int("res1" in tag) + int("dual1" in tag)
                    gb = g_ * m_ * c_ * 2 * nbuf * c / 1e9
                rows.append((t, base.replace("mauv_", ""), tag, c, gf / t if t else 0, gb / t if t else 0))
            rows.sort(reverse=True)
            sys.stderr.write("ms      kernel                 shape                                   calls TFLOP/s  TB/s(min traffic)\n")
            for t, b, tag, c, tf, tb in rows[:60]:
                sys.stderr.write(f"{t:7.2f} {b:22s} {tag:40s} {c:4d} {tf:7.1f} {tb:7.2f}\n")
        # per-launch roofline: ideal time of a launch = max(flops / tensor peak, minimum bytes / HBM peak)
        ideal_ms = hbm_bound_ms = 0.0
        for k, (c, t) in prof.items():
            if "|" not in k or k.split("|")[0] not in tc:
                continue
            mdl = tc_launch_model(*k.split("|"))
            if mdl is None:
                continue
            t_tc, t_hbm = mdl[0] / (tf_peak * 1e12) * 1e3, mdl[1] / (hbm_peak * 1e9) * 1e3
            ideal_ms += c * max(t_tc, t_hbm)
            hbm_bound_ms += t if t_hbm > t_tc else 0.0
        s_local = hi - lo
        conv_tflop = GFLOP_PER_SAMPLE[MODEL_KIND] * B * s_local / 1e3
        achieved = conv_tflop / (conv_ms / 1e3) if conv_ms > 0 else 0.0
        total_prof = sum(v[1] for v in fam.values()) or 1.0
        line = {
            "metric": METRIC if MODEL_KIND == "multimodal" else f"MC-sampled patches/sec (S={S}, unimodal {MODEL_KIND} branch)",
            "value": value, "unit": UNIT if MODEL_KIND == "multimodal" else "patches/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16 operands / f32 accumulate+statistics (reference predictor: autocast fp16 on CUDA)",
            "data": "synthetic",
            "config": {"workload": f"{'cfg2 multimodal' if MODEL_KIND == 'multimodal' else 'cfg4 unimodal ' + MODEL_KIND} BNN inference "
                                   f"B={B} S={S} 256x256 C={C_CLASSES}, MC samples sharded over "
                                   f"{world} GPU(s), group={group}",
                       "l2": "no flush: per-step inputs (470 MB) and activations (GBs) exceed the 126 MB L2",
                       "triplet_samples_per_s": value * S},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": sum(x.numel() * 4 for x in host_in),
                    "d2h_bytes_per_step": B * 5 * 4},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "gemm_f16_tc_kernel + conv3x3_c64_stream_kernel + stem_conv_pool_kernel (tcgen05 implicit-GEMM conv, all 159 convs)",
                         "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak,
                         "peak_source": peak_src,
                         # mean DRAM read+write bytes per launch of these kernels in one step, from the committed ncu
                         # capture of this command (profiles/traffic.json, written by tests/tools/ncu_summary.py from
                         # dram__bytes_read.sum + dram__bytes_write.sum); null when no capture of this tree exists
                         "traffic": load_traffic(),
                         "note": "algorithmic FLOPs (31.824 GFLOP per triplet-sample, recompute passes not counted) / "
                                 "summed launch durations; the kernel is tensor-bound on the K >= 1152 shapes (0.9 of the "
                                 "sustained peak at K >= 2304) and HBM-bound on the 1x1 layers of layer1/2 (see per_launch)",
                         "launches_per_step": conv_calls, "avg_launch_ms": conv_ms / max(conv_calls, 1),
                         "share_of_step": conv_ms / total_prof,
                         # the kernel is tensor-bound on some shapes and HBM-bound on others: per launch,
                         # ideal = max(flops / tensor peak, minimum bytes (inputs + weights + outputs once) / HBM peak)
                         "per_launch": {"ideal_ms_per_step": ideal_ms, "measured_ms_per_step": conv_ms,
                                        "frac": ideal_ms / conv_ms if conv_ms else None,
                                        "hbm_bound_share_of_time": hbm_bound_ms / conv_ms if conv_ms else None,
                                        "hbm_peak_gbs": hbm_peak}},
            "kernel_ms_per_step": {k: round(v[1], 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1])},
            "cpu_baseline": cpu_base,
        }
        if per_rank is not None:
            line["per_rank"] = per_rank
        if x3_rec is not None:
            line["x3"] = x3_rec
        if train_rec is not None:
            line["train"] = train_rec
        print(json.dumps(line))
    if dist:
        torch.distributed.destroy_process_group()
    return 0


def tc_launch_model(base: str, tag: str):
    """Algorithmic (flops, minimum HBM bytes) of ONE launch of gemm_f16_tc_kernel from its profiling tag."""
    if base in ("mauv_wgrad_f16", "mauv_gram_bn_f16"):     # "G10x32 Cout64 K64 px4096": [Cout x px] x [px x K] per (sample, chunk); reads a once
        t = tag.split()
        g, sp = (int(v) for v in t[0][1:].split("x"))
        cout, k, px = int(t[1][4:]), int(t[2][1:]), int(t[3][2:])
        return 2.0 * g * sp * cout * k * px, g * sp * px * max(cout, k) * 2 + g * sp * cout * k * 4
    toks = {x[0]: int(x[1:]) for x in tag.split() if x[0] in "GMNK" and x[1:].isdigit()}
    if not all(q in toks for q in "GMNK"):
        return None
    g, m, n, k = (toks[q] for q in "GMNK")
    flops = 2.0 * g * m * n * k
    w_bytes = g * n * k * 2
    if base == "mauv_stem_conv_pool_f16":                      # A shared by all samples; only the pooled (1/4) output is written
        a_bytes = m * k * 2
        y_bytes = g * (m // 4) * n * 2
    elif base == "mauv_conv3x3_c64_f16":
        a_bytes = g * m * 64 * 2
        y_bytes = g * m * n * 2
    elif base == "mauv_conv2d_im2col_f16":
        geo = tag.split()[-1]                                   # "3x3/1"
        kk, stride = geo.split("/")
        taps = int(kk.split("x")[0]) * int(kk.split("x")[1])
        a_bytes = g * m * int(stride) ** 2 * (k // taps) * 2    # the NHWC input is read once
        y_bytes = g * m * n * 2
    elif base in ("mauv_gemm_bn_f16", "mauv_gemm_bn_xf_f16"):
        a_bytes = g * m * k * 2
        y_bytes = 0 if tag.startswith("stats") else g * m * n * 2 * (2 if "res1" in tag else 1)
    else:
        a_bytes = (m if k % 64 else g * m) * k * 2              # K = 152 / 56: the stem's im2col matrix, shared by all samples
        y_bytes = g * m * n * 2
    return flops, a_bytes + w_bytes + y_bytes


def train_leg(args, B: int, S: int, steps: int, warmup: int, detail: bool = False) -> dict:
    """cfg3 ELBO training step (BASELINE configs[2]; reference train/multimodal.py:104-145): S stochastic passes of the
    multimodal BNN through the S-batched TrainEngine (one grouped forward + one grouped hand-written backward), loss =
    CE(mean logits) + KL/B * 2^(e+1)/2^E, NaN/Inf guard + Adam in one fused device pass. N > 1: the minibatch is split over
    the ranks (B triplets per GPU, all S samples on every rank - nn.DataParallel semantics) and the flat
    fp32 gradient buffer is averaged with one NCCL all-reduce per step. Must be called by every rank. Returns the record
    (max over ranks of the device-timed step)."""
    from mauv import ops
    from mauv.bayesian import get_kl_loss
    from mauv.train_engine import TrainEngine
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = world > 1
    use_engine = args.train_path == "engine"
    model = build_model_cpu().cuda().train()
    opt = torch.optim.Adam(model.parameters(), lr=5e-5)
    g = torch.Generator().manual_seed(4321 + rank)             # every rank trains on its own shard of the global minibatch
    xs = [x.cuda() for x in synthetic_inputs(B, seed=4321 + rank)]
    labels = torch.randint(0, C_CLASSES, (B,), generator=g).cuda()
    kl_w = 2.0 / 2 ** 20                                       # epoch 0 of 20
    ar_events = []
    if use_engine:
        eng = TrainEngine(model)
        eng.flatten_grads()

        def step(timed_ar=False):
            eng.zero_grad()
            res = eng.step(xs, labels, S, kl_w / (B * world))
            if dist:
                if timed_ar:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    eng.allreduce_grads()
                    e1.record()
                    ar_events.append((e0, e1))
                else:
                    eng.allreduce_grads()
            eng.optimizer_step(opt)                            # fused finite guard + Adam (mauv.optim.FusedAdam)
            return res["loss"]
    else:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], broadcast_buffers=False) if dist else model

        def step(timed_ar=False):
            opt.zero_grad(set_to_none=False)
            out = torch.mean(torch.stack([net(*xs) for _ in range(S)]), dim=0)
            loss = torch.nn.functional.cross_entropy(out, labels) + get_kl_loss(model) / (B * world) * kl_w
            loss.backward()
            opt.step()
            return loss

    # warm-up: step 1 runs eagerly, the fused optimizer re-homes the parameters, step 2 runs eagerly with the final pointers,
    # step 3 records the CUDA graph of the step (seconds of host time) - the timed region only sees replays
    warmup = max(4, warmup)
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if dist:
        torch.distributed.barrier()
    l0 = ops.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step(timed_ar=True)
    e1.record()
    torch.cuda.synchronize()
    if dist:
        torch.distributed.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps, sum(a.elapsed_time(b) for a, b in ar_events) / max(1, len(ar_events))],
                     device="cuda")
    if dist:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, ar_ms = t.tolist()
    launches = ops.launch_count - l0
    prof_detail = None
    if detail and use_engine and rank == 0:
        ops.start_profile()
        step()
        prof = ops.stop_profile()
        agg = {}
        for k, (c, tms) in prof.items():
            n = k.split("|")[0]
            a = agg.get(n, (0, 0.0))
            agg[n] = (a[0] + c, a[1] + tms)
        prof_detail = {k: {"calls": c, "ms": round(tms, 3)} for k, (c, tms) in sorted(agg.items(), key=lambda kv: -kv[1][1])}
        for k, (c, tms) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:60]:
            print(f"{tms:9.3f} ms {c:4d} x {k}", file=sys.stderr)
    elif detail and dist:
        step()                                                 # keep the ranks' collectives in lockstep with rank 0's extra step
    # forward + dgrad + wgrad per triplet-sample (a unimodal branch: ~1/3 of it)
    _, tf_peak, peak_src = load_peaks()
    flops = 94.77e9 * (GFLOP_PER_SAMPLE[MODEL_KIND] / CONV_GFLOP_PER_TRIPLET_SAMPLE) * world * B * S
    tfl = flops / (ms / 1e3) / 1e12
    rec = {"metric": "ELBO training triplets/sec" if MODEL_KIND == "multimodal" else f"ELBO training patches/sec (unimodal {MODEL_KIND})",
           "value": world * B / (ms / 1e3), "unit": UNIT if MODEL_KIND == "multimodal" else "patches/s",
           "n_gpus": world, "scaling": "weak", "steps": steps, "warmup": warmup, "ms_per_step": ms,
           "higher_is_better": True, "data": "synthetic", "dtype": "f16 operands / f32 accumulate, fp32 parameter gradients, fp32 Adam",
           "config": {"workload": f"{'cfg3 multimodal' if MODEL_KIND == 'multimodal' else 'cfg4 unimodal ' + MODEL_KIND} ELBO step, "
                                  f"{B} triplets per GPU, S={S}, 256x256, C=7, fused guarded Adam, "
                                  f"path={args.train_path}, gradient all-reduce over {world} GPU(s)",
                      "triplet_samples_per_s": world * B * S / (ms / 1e3)},
           "model_tflops": tfl,
           "roofline": {"bound": "tensor", "achieved": tfl / world, "peak": tf_peak, "unit": "TFLOP/s", "frac": tfl / world / tf_peak,
                        "peak_source": peak_src,
                        "note": "whole step (forward + hand-written backward + KL + all-reduce + Adam) against 94.77 GFLOP per "
                                "triplet-sample (SURVEY 8d: fwd + dgrad + wgrad), per GPU"},
           "allreduce_ms": ar_ms if dist else 0.0,
           "gpu_launches": launches, "loss": float(loss.detach()),
           "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    if prof_detail is not None:
        rec["kernel_ms_per_step"] = prof_detail
    del model, opt, xs
    if use_engine:
        del eng
    torch.cuda.empty_cache()
    return rec


def run_train(args):
    """`--workload train`: the cfg3 (or, with --model, cfg4) ELBO training step as its own bench line."""
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    use_engine = args.train_path == "engine"
    B = args.batch if args.batch != B_FULL else 8
    S = args.samples if args.samples != S_FULL else (30 if use_engine else 5)
    rec = train_leg(args, B, S, args.steps, args.warmup, detail=args.detail)
    if rank == 0:
        print(json.dumps(rec))
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=B_FULL)
    ap.add_argument("--samples", type=int, default=S_FULL)
    ap.add_argument("--group", type=int, default=30, help="MC samples walked together (30: one walk of the network per batch; "
                                                            "53 GB peak at B=256; measured 461 / 470 / 475 triplets/s at 10 / 15 / 30)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-x3", action="store_true", help="skip the one-step fp32-class (x3) sub-record")
    ap.add_argument("--no-train-leg", action="store_true", help="skip the cfg3 ELBO training sub-record of the default line")
    ap.add_argument("--detail", action="store_true", help="per-shape kernel table on stderr")
    ap.add_argument("--train-path", default="engine", choices=["engine", "layers"],
                    help="--workload train: S-batched TrainEngine (default) or the drop-in layer path")
    ap.add_argument("--model", default="multimodal", choices=["multimodal", "image", "bathy", "sss"],
                    help="multimodal (cfg2 / cfg3, the headline) or one unimodal ResNet50Custom branch (cfg4)")
    ap.add_argument("--workload", default="inference", choices=["inference", "train"],
                    help="inference = BASELINE cfg2 (headline); train = cfg3-style ELBO step (secondary)")
    args = ap.parse_args()
    global MODEL_KIND
    MODEL_KIND = args.model
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.workload == "train" and args.impl == "ours":
        return run_train(args)
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
