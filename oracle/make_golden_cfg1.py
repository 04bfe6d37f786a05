"""CPU ORACLE tooling - generates tests/golden/cfg1_philox.pt: the fp32 oracle's logits at a BASELINE config size.

BASELINE.json configs[0] (cfg1): multimodal BNN inference, batch 8 synthetic (image, bathymetry, side-scan) triplets of
256x256, 10 MC samples, 7 classes, random init + MOPED (delta = 0.1). The reference draws eps with torch's global RNG;
the CUDA path draws it in-kernel from Philox4x32-10 + Box-Muller, so "identical injected eps" at this size (733 M
normals = 2.9 GB, far too large to commit or to upload per test) means: the oracle regenerates the PRODUCTION eps
stream on the CPU (oracle/philox.py: counter = (element quad, sample id, layer id), key = seed) and injects it into
the fp32 restatement (oracle/bnn_oracle.py), and the GPU test runs the production path with the same seed. The logits
[S, B, C] are a few KB and are committed, so the GPU box does not spend minutes of CPU time per test run.

A second record does the same for one unimodal ResNet50Custom (image branch, cfg4's shape) with S = 4: the unimodal net
is the ill-conditioned case of DESIGN.md section 4.3 and is what the fp32-class validation mode has to hold up on.

Run:  python oracle/make_golden_cfg1.py          (about 1 minute on 8 cores; needs ~6 GB of RAM)
      python oracle/make_golden_cfg1.py --cfg2   (tests/golden/cfg2_philox.pt: the FULL cfg2 batch, B = 256 at 256x256, for
                                                  MC samples 0 and 29 of the S = 30 range; ~2 minutes, ~25 GB of RAM)
      python oracle/make_golden_cfg1.py --cfg2 --all   (cfg2_philox_s30.pt: all 30 samples + the statistics; ~25 minutes)
"""
from __future__ import annotations

import os
import sys
import time
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import bnn_oracle as O  # noqa: E402
import philox  # noqa: E402

OUT = HERE.parent / "tests" / "golden" / "cfg1_philox.pt"
SEED_EPS = 20261018


def philox_eps(model, layer_ids, seed: int, s: int) -> dict:
    """eps of MC sample `s` for every Bayesian layer, in the layout bnn_oracle.inject_eps expects ([S=1, ...])."""
    eps = {}
    for name, layer in O.bayesian_layers(model):
        lid = layer_ids[name]
        w = layer.mu_kernel if hasattr(layer, "mu_kernel") else layer.mu_weight
        e = {"w": torch.from_numpy(philox.philox_normal(w.numel(), seed, lid, s)).view(1, *w.shape), "b": None}
        if layer.mu_bias is not None:
            e["b"] = torch.from_numpy(philox.philox_normal(layer.mu_bias.numel(), seed, lid | 0x80000000, s)).view(1, -1)
        eps[name] = e
    return eps


@torch.no_grad()
def mc_logits_philox(model, inputs, S: int, seed: int, sample0: int = 0) -> torch.Tensor:
    layer_ids = {n: i for i, (n, _) in enumerate(O.bayesian_layers(model))}
    model.train()        # BatchNorm batch statistics in every MC pass (reference predictors.py:27)
    outs = []
    for s in range(S):
        t0 = time.time()
        O.inject_eps(model, philox_eps(model, layer_ids, seed, sample0 + s), 0)
        outs.append(model(*inputs).float())
        print(f"  pass {s + 1}/{S}: {time.time() - t0:.1f}s", flush=True)
    O.inject_eps(model, None, 0)
    return torch.stack(outs)


def main_cfg2():
    """BASELINE configs[1] at its full size: batch 256 of 256x256 triplets; two of the 30 MC samples (the first and the
    last sample id, i.e. the first sample of rank 0 and the last sample of the last rank under sample sharding)."""
    torch.set_num_threads(int(os.environ.get("ORACLE_THREADS", max(1, os.cpu_count() or 1))))
    B, SIZE, C = 256, 256, 7
    full = "--all" in sys.argv          # every sample id of the S = 30 range (~25 minutes): cfg2_philox_s30.pt
    sample_ids = tuple(range(30)) if full else (0, 29)
    gold = {"B": B, "size": SIZE, "C": C, "seed_w": 1234, "seed_x": 1234, "seed_eps": SEED_EPS, "sample_ids": sample_ids,
            "torch": torch.__version__}
    mm = O.define_models(C, unimodal=True)["multimodal_model"]
    gold["param_checksum"] = float(sum(p.detach().double().sum() for p in mm.parameters()))
    img, bathy, sss, labels = O.synthetic_batch(B, seed=gold["seed_x"], size=SIZE)
    outs = []
    for sid in sample_ids:
        print(f"multimodal cfg2 sample id {sid} ...", flush=True)
        outs.append(mc_logits_philox(mm, (img, bathy, sss), 1, SEED_EPS, sample0=sid)[0])
    gold["logits_fp32"] = torch.stack(outs)
    if full:
        gold["predictor_stats"] = O.predictor_stats(gold["logits_fp32"])          # inference/predictors.py:65-84
        gold["eval_stats"] = O.multimodal_eval_stats(gold["logits_fp32"])         # train/multimodal.py:287-310
        gold["labels"] = labels.clone()
    out = OUT.with_name("cfg2_philox_s30.pt" if full else "cfg2_philox.pt")
    torch.save(gold, out)
    print("wrote", out, "logits absmax", float(gold["logits_fp32"].abs().max()))


def main():
    if "--cfg2" in sys.argv:
        return main_cfg2()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    B, S, SIZE, C = 8, 10, 256, 7
    gold = {"B": B, "S": S, "size": SIZE, "C": C, "seed_w": 1234, "seed_x": 1234, "seed_eps": SEED_EPS,
            "torch": torch.__version__}
    models = O.define_models(C, unimodal=True)          # seed 1234: same weights as every other fixture
    mm = models["multimodal_model"]
    gold["layer_ids"] = {n: i for i, (n, _) in enumerate(O.bayesian_layers(mm))}
    gold["param_checksum"] = float(sum(p.detach().double().sum() for p in mm.parameters()))
    img, bathy, sss, labels = O.synthetic_batch(B, seed=gold["seed_x"], size=SIZE)
    gold["labels"] = labels.clone()
    print("multimodal cfg1 ...", flush=True)
    logits = mc_logits_philox(mm, (img, bathy, sss), S, SEED_EPS)
    gold["logits_fp32"] = logits.clone()
    gold["predictor_stats"] = O.predictor_stats(logits)               # inference/predictors.py:65-84
    gold["eval_stats"] = O.multimodal_eval_stats(logits)              # train/multimodal.py:287-310
    gold["kl"] = float(O.get_kl_loss(mm).detach())

    um = models["image_model"]
    Su = 4
    gold["uni_S"] = Su
    gold["uni_layer_ids"] = {n: i for i, (n, _) in enumerate(O.bayesian_layers(um))}
    gold["uni_param_checksum"] = float(sum(p.detach().double().sum() for p in um.parameters()))
    print("unimodal image branch ...", flush=True)
    ulogits = mc_logits_philox(um, (img,), Su, SEED_EPS + 1)
    gold["uni_logits_fp32"] = ulogits.clone()
    gold["uni_eval_stats"] = O.unimodal_eval_stats(ulogits)           # train/unimodal.py:268-308

    OUT.parent.mkdir(parents=True, exist_ok=True)
    torch.save(gold, OUT)
    print("wrote", OUT, "logits absmax", float(logits.abs().max()), "uni absmax", float(ulogits.abs().max()))


if __name__ == "__main__":
    main()
