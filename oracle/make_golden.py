"""CPU ORACLE tooling — generates tests/golden/*.pt by running THE REFERENCE'S OWN CODE in this
container (it cannot travel to the GPU box, so its outputs are committed as small fixtures).

What runs from /root/reference (loaded by file path, unmodified):
  src/Multimodal_AUV/models/base_models.py     ResNet50Custom, AdditiveAttention, MultiModalModel
  src/Multimodal_AUV/models/model_utils.py     define_models (calls dnn_to_bnn)
  src/Multimodal_AUV/inference/predictors.py   multimodal_predict_and_save
  src/Multimodal_AUV/train/multimodal.py       train_multimodal_model, evaluate_multimodal_model
  src/Multimodal_AUV/train/unimodal.py         train_unimodal_model, evaluate_unimodal_model
What is substituted (absent from the container / needs the network):
  bayesian_torch.models.dnn_to_bnn  -> oracle/bnn_oracle.py restatement (package not installable offline)
  torchvision ImageNet checkpoints  -> random init (resnet50(weights=None)), as BASELINE.json specifies
  matplotlib                        -> a stub module (only used for the confusion-matrix PNG)

The fixtures pin oracle/bnn_oracle.py's restatement of the model topology, the predictor statistics,
the evaluation statistics and the ELBO loss against the reference's real driver code, on identical
weights (seeded) and identical injected eps (seeded).

Run:  python oracle/make_golden.py        (about 2 minutes on 8 cores)
"""
from __future__ import annotations

import csv
import importlib.util
import os
import sys
import tempfile
import types
from pathlib import Path

import torch
import torch.nn as nn

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import bnn_oracle as O  # noqa: E402

REF = Path("/root/reference/src/Multimodal_AUV")
OUT = HERE.parent / "tests" / "golden"


def _load(name: str, path: Path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    """Import the reference's hot-path modules by path over the minimum set of stubs."""
    # bayesian_torch shim -> oracle restatement
    bt = types.ModuleType("bayesian_torch")
    btm = types.ModuleType("bayesian_torch.models")
    btd = types.ModuleType("bayesian_torch.models.dnn_to_bnn")
    btd.dnn_to_bnn, btd.get_kl_loss = O.dnn_to_bnn, O.get_kl_loss
    sys.modules.update({"bayesian_torch": bt, "bayesian_torch.models": btm, "bayesian_torch.models.dnn_to_bnn": btd})
    # matplotlib stub (confusion-matrix PNG only; the reference already wraps plotting in try/except)
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    plt.rcParams = {}

    # plt.subplots must SUCCEED: reference train/unimodal.py:330-347 reads `fig` in a `finally` without
    # initialising it, so a failing subplots() would send the whole evaluation into its bare `except`.
    plt.subplots = lambda *a, **k: (object(), object())
    plt.title = plt.savefig = plt.close = lambda *a, **k: None
    mpl.pyplot = plt
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt})
    # package skeleton so that `from Multimodal_AUV.x.y import z` resolves without the package __init__
    for pkg in ("Multimodal_AUV", "Multimodal_AUV.models", "Multimodal_AUV.train", "Multimodal_AUV.inference"):
        sys.modules[pkg] = types.ModuleType(pkg)
    base = _load("Multimodal_AUV.models.base_models", REF / "models" / "base_models.py")
    import torchvision
    base.resnet50 = lambda weights=None: torchvision.models.resnet50(weights=None)     # no download
    utils = _load("Multimodal_AUV.models.model_utils", REF / "models" / "model_utils.py")
    utils.resnet50 = base.resnet50
    _load("Multimodal_AUV.train.checkpointing", REF / "train" / "checkpointing.py")
    pred = _load("Multimodal_AUV.inference.predictors", REF / "inference" / "predictors.py")
    tmm = _load("Multimodal_AUV.train.multimodal", REF / "train" / "multimodal.py")
    tum = _load("Multimodal_AUV.train.unimodal", REF / "train" / "unimodal.py")
    return base, utils, pred, tmm, tum


class _EpsInjector:
    """Forward pre-hook: before the s-th forward of `model`, inject eps sample s into every layer."""

    def __init__(self, model, eps):
        self.model, self.eps, self.s = model, eps, 0
        self.h = model.register_forward_pre_hook(self)

    def __call__(self, module, args):
        O.inject_eps(self.model, self.eps, self.s)
        self.s += 1

    def remove(self):
        self.h.remove()
        O.inject_eps(self.model, None, 0)


class _ListLoader(list):
    """Minimal DataLoader stand-in (the drivers use iteration, len() and .batch_size)."""
    batch_size = None


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    base, utils, pred, tmm, tum = load_reference()
    B, S, SIZE, C = 2, 3, 64, 7
    SEED_W, SEED_X, SEED_EPS = 1234, 4321, 77
    OUT.mkdir(parents=True, exist_ok=True)
    gold = {"B": B, "S": S, "size": SIZE, "C": C, "seed_w": SEED_W, "seed_x": SEED_X, "seed_eps": SEED_EPS,
            "torch": torch.__version__}

    # ---- models built by the reference's define_models (random init + MOPED) -------------------
    torch.manual_seed(SEED_W)
    models = utils.define_models(torch.device("cpu"), C, dict(O.DEFAULT_PRIOR))
    mm = models["multimodal_model"]
    # the oracle's restated topology must accept the reference model's state_dict verbatim
    torch.manual_seed(0)
    o_mm = O.define_models(C, seed=None, unimodal=False)["multimodal_model"]
    o_mm.load_state_dict(mm.state_dict(), strict=True)
    gold["n_bayes_layers"] = len(O.bayesian_layers(mm))
    gold["param_checksum"] = float(sum(p.detach().double().sum() for p in mm.parameters()))

    img, bathy, sss, labels = O.synthetic_batch(B, seed=SEED_X, size=SIZE)
    eps = O.draw_eps(mm, S, SEED_EPS)

    # ---- (1) raw fp32 logits: reference model forward, S passes, injected eps ------------------
    mm.train()
    inj = _EpsInjector(mm, eps)
    with torch.no_grad():
        ref_logits = torch.stack([mm(img, bathy, sss) for _ in range(S)])
    inj.remove()
    gold["logits_fp32"] = ref_logits.clone()
    gold["kl"] = float(O.get_kl_loss(mm))
    gold["bn1_running_mean_after"] = mm.image_model_feat.bn1.running_mean.clone()

    # ---- (2) the shipped predictor (autocast bf16 on CPU): CSV rows ------------------------------
    mm.load_state_dict(o_mm.state_dict())      # reset BN running stats to the pre-forward state
    loader = _ListLoader([(img, bathy, sss, [f"img_{i}" for i in range(B)])])
    inj = _EpsInjector(mm, eps)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "pred.csv")
        # The shipped predictor prints `tensor.cpu().numpy()` of autocast outputs (predictors.py:74,80); on CPU
        # autocast means bf16, which numpy rejects (TypeError) - as shipped, the predictor only runs on CUDA
        # (fp16). Only for this print we let .numpy() upcast bf16; the CSV values come from .item() and are
        # untouched.
        _numpy = torch.Tensor.numpy
        torch.Tensor.numpy = lambda self, *a, **k: _numpy(self.float() if self.dtype == torch.bfloat16 else self, *a, **k)
        try:
            pred.multimodal_predict_and_save(mm, loader, torch.device("cpu"), p, num_mc_samples=S)
        finally:
            torch.Tensor.numpy = _numpy
        rows = list(csv.reader(open(p)))
    inj.remove()
    gold["predictor_csv_header"] = rows[0]
    gold["predictor_csv_rows"] = [[r[0], int(r[1]), float(r[2]), float(r[3])] for r in rows[1:]]

    # ---- (3) evaluate_multimodal_model (fp32, no autocast) ---------------------------------------
    mm.load_state_dict(o_mm.state_dict())
    batch = {"main_image": img, "label": labels, "bathy_image": bathy, "sss_image": sss}
    loader = _ListLoader([batch])
    loader.batch_size = B
    inj = _EpsInjector(mm, eps)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "logs", "eval.csv")
        os.makedirs(os.path.dirname(p))
        acc = tmm.evaluate_multimodal_model(mm, loader, torch.device("cpu"), epoch=0, total_num_epochs=20,
                                            num_mc=S, model_type="multimodal", csv_path=p)
        rows = list(csv.reader(open(p)))
    inj.remove()
    gold["eval_mm_accuracy"] = float(acc)
    gold["eval_mm_csv_header"] = rows[0]
    gold["eval_mm_csv_row"] = rows[1]

    # ---- (4) train_multimodal_model: one ELBO step with Adam -------------------------------------
    mm.load_state_dict(o_mm.state_dict())
    opt = torch.optim.Adam(mm.parameters(), lr=1e-4)

    class _W:
        def add_scalar(self, *a, **k):
            pass
    inj = _EpsInjector(mm, eps)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "logs", "train.csv")
        os.makedirs(os.path.dirname(p))
        loss, accuracy = tmm.train_multimodal_model(mm, loader, nn.CrossEntropyLoss(), opt, epoch=1, device=torch.device("cpu"),
                                                    model_type="multimodal", total_num_epochs=20, num_mc=S,
                                                    sum_writer=_W(), csv_path=p)
        rows = list(csv.reader(open(p)))
    inj.remove()
    gold["train_mm_return"] = (float(loss), float(accuracy))
    gold["train_mm_csv_row"] = rows[1]
    gold["labels"] = labels.clone()
    # a few updated parameters (pins gradient flow through mu and rho + the Adam step)
    sd = mm.state_dict()
    gold["train_mm_after"] = {k: sd[k].flatten()[:8].clone() for k in
                              ("fc2.mu_weight", "fc2.rho_weight", "fc2.mu_bias", "image_model_feat.conv1.mu_kernel",
                               "image_model_feat.conv1.rho_kernel", "sss_model_feat.layer4.2.conv3.rho_kernel")}

    # ---- (5) unimodal: evaluate + train on the image branch --------------------------------------
    um = models["image_model"]
    o_um = O.ResNet50Custom(3, C)
    O.dnn_to_bnn(o_um, O.DEFAULT_PRIOR)
    o_um.load_state_dict(um.state_dict(), strict=True)
    eps_u = O.draw_eps(um, S, SEED_EPS + 1)
    um.train()
    inj = _EpsInjector(um, eps_u)
    with torch.no_grad():
        gold["uni_logits_fp32"] = torch.stack([um(img) for _ in range(S)])
    inj.remove()
    um.load_state_dict(o_um.state_dict())
    inj = _EpsInjector(um, eps_u)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "logs", "ueval.csv")
        os.makedirs(os.path.dirname(p))
        acc = tum.evaluate_unimodal_model(um, loader, torch.device("cpu"), epoch=0, csv_path=p, total_num_epochs=20,
                                          num_mc=S, model_type="image")
        rows = list(csv.reader(open(p)))
    inj.remove()
    gold["eval_uni_accuracy"] = float(acc)
    gold["eval_uni_csv_header"] = rows[0]
    gold["eval_uni_csv_row"] = rows[1]

    # ---- (6) train_unimodal_model: one ELBO step with Adam on the image branch ---------------------
    um.load_state_dict(o_um.state_dict())
    opt_u = torch.optim.Adam(um.parameters(), lr=1e-4)
    inj = _EpsInjector(um, eps_u)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "logs", "utrain.csv")
        os.makedirs(os.path.dirname(p))
        acc_u, loss_u = tum.train_unimodal_model(um, loader, nn.CrossEntropyLoss(), opt_u, epoch=1, total_num_epochs=20, num_mc=S,
                                                 sum_writer=_W(), device=torch.device("cpu"), model_type="image", csv_path=p)
        rows = list(csv.reader(open(p)))
    inj.remove()
    gold["train_uni_return"] = (float(acc_u), float(loss_u))
    gold["train_uni_csv_header"] = rows[0]
    gold["train_uni_csv_row"] = rows[1]
    sd = um.state_dict()
    gold["train_uni_after"] = {k: sd[k].flatten()[:8].clone() for k in
                               ("model.fc.mu_weight", "model.fc.rho_weight", "model.fc.mu_bias", "model.conv1.mu_kernel",
                                "model.conv1.rho_kernel", "model.layer4.2.conv3.rho_kernel")}
    gold["train_uni_before"] = {k: o_um.state_dict()[k].flatten()[:8].clone() for k in gold["train_uni_after"]}

    torch.save(gold, OUT / "reference_small.pt")
    print("wrote", OUT / "reference_small.pt", {k: (tuple(v.shape) if torch.is_tensor(v) else v)
                                                for k, v in gold.items() if not k.endswith(("_after", "_before"))})


if __name__ == "__main__":
    main()
