"""The CPU oracle (oracle/bnn_oracle.py, oracle/philox.py) against closed forms, published known-answer
vectors and the golden fixtures produced by the reference's own code (oracle/make_golden.py)."""
import math
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import bnn_oracle as O
import philox

GOLD = Path(__file__).resolve().parent / "golden" / "reference_small.pt"


# ------------------------------------------------------------------ bayesian-torch restatement vs closed forms
def test_kl_div_matches_torch_distributions():
    torch.manual_seed(0)
    mu = torch.randn(64, 32, 3, 3, dtype=torch.float64) * 0.1
    rho = torch.randn(64, 32, 3, 3, dtype=torch.float64) - 3
    sigma = torch.log1p(torch.exp(rho))
    ref = torch.distributions.kl_divergence(torch.distributions.Normal(mu, sigma),
                                            torch.distributions.Normal(0.0, 1.0)).mean()
    got = O.kl_div(mu, sigma, torch.tensor(0.0, dtype=torch.float64), torch.tensor(1.0, dtype=torch.float64))
    assert abs(got.item() - ref.item()) < 1e-12 * abs(ref.item()) + 1e-12


def test_moped_rho_gives_delta_abs_w():
    w = torch.randn(1000) * 0.05
    rho = O.get_rho(w, 0.1)
    sigma = torch.log1p(torch.exp(rho.double()))
    assert torch.allclose(sigma, (0.1 * w.abs()).double() + 1e-20, rtol=1e-4, atol=1e-9)
    assert torch.isfinite(O.get_rho(torch.zeros(3), 0.1)).all()      # +1e-20 keeps rho finite (~ -46)
    assert abs(O.get_rho(torch.zeros(1), 0.1).item() - math.log(1e-20)) < 1e-3


def test_layers_follow_reparameterisation():
    torch.manual_seed(1)
    conv = O.Conv2dReparameterization(8, 16, 3, stride=2, padding=1, bias=False)
    conv.dnn_to_bnn_flag = True
    x = torch.randn(2, 8, 10, 10)
    eps = torch.randn_like(conv.mu_kernel)
    conv.injected_eps_kernel = eps
    w = conv.mu_kernel + torch.log1p(torch.exp(conv.rho_kernel)) * eps
    assert torch.equal(conv(x), F.conv2d(x, w, None, 2, 1))
    assert torch.equal(conv.eps_kernel, eps)          # buffer holds the eps used (capture point)
    lin = O.LinearReparameterization(12, 5)
    out, kl = lin(torch.randn(3, 12))                 # stand-alone layers return (out, kl)
    assert out.shape == (3, 5) and kl.ndim == 0
    conv.injected_eps_kernel = None
    assert not torch.equal(conv(x), conv(x))          # fresh eps per call


def test_dnn_to_bnn_converts_every_conv_and_linear():
    m = O.define_models(7, unimodal=False)["multimodal_model"]
    layers = O.bayesian_layers(m)
    assert len(layers) == 174                                                    # 159 conv + 15 linear
    assert sum(hasattr(l, "mu_kernel") for _, l in layers) == 159
    n_w = sum((l.mu_kernel if hasattr(l, "mu_kernel") else l.mu_weight).numel() for _, l in layers)
    n_b = sum(l.mu_bias.numel() for _, l in layers if l.mu_bias is not None)
    assert (n_w, n_b) == (73301280, 2859)
    assert all(l.dnn_to_bnn_flag for _, l in layers)
    assert not any(isinstance(x, (torch.nn.Conv2d, torch.nn.Linear)) for x in m.modules())
    # registration order (= get_kl_loss order): trunks, fc/fc1/fc2, then the three attentions
    # (reference models/base_models.py:57-70 registers fc* before attention_*)
    names = [n for n, _ in layers]
    assert names[0] == "image_model_feat.conv1" and names[-1] == "attention_sss.attention_mechanism"
    assert names.index("fc2") < names.index("attention_image.query_projection")
    kl = O.get_kl_loss(m)
    per = sum(l.kl_loss() for _, l in layers)
    assert torch.allclose(kl, per)


# ------------------------------------------------------------------ MC statistics identities
def test_uncertainty_identities():
    torch.manual_seed(2)
    lg = torch.randn(30, 64, 7) * 2
    p = O.predictor_stats(lg)
    m = O.multimodal_eval_stats(lg)
    u = O.unimodal_eval_stats(lg)
    ln7 = math.log(7)
    assert (p["aleatoric_uncertainty"] >= 0).all() and (p["aleatoric_uncertainty"] <= ln7 + 1e-5).all()
    assert (m["predictive_uncertainty"] <= ln7 + 1e-5).all()
    assert (m["model_uncertainty"] >= -1e-5).all()                 # MI >= 0 (Jensen), up to epsilon
    assert torch.allclose(p["predictive_uncertainty"], u["predictive_uncertainty"])
    same = torch.zeros(5, 4, 7) + torch.randn(1, 4, 7)             # identical samples -> zero epistemic part
    assert O.predictor_stats(same)["predictive_uncertainty"].abs().max() < 1e-12
    assert O.multimodal_eval_stats(same)["model_uncertainty"].abs().max() < 1e-6
    with pytest.warns(UserWarning):
        assert torch.isnan(O.predictor_stats(lg[:1])["predictive_uncertainty"]).all()   # S=1 -> NaN (predictors.py:73)


def test_elbo_losses():
    torch.manual_seed(3)
    lg = torch.randn(4, 6, 7)
    y = torch.randint(0, 7, (6,))
    kl = torch.tensor(1031.0)
    loss, ce, skl = O.elbo_loss_multimodal(lg, y, kl, batch_size=6, epoch=0, total_num_epochs=20)
    assert abs(skl.item() - 1031.0 / 6 * 2 ** -19) < 1e-9 and torch.allclose(loss, ce + skl)
    loss_u, ce_u, skl_u = O.elbo_loss_unimodal(lg, y, kl, 6, 0, 20)
    assert torch.allclose(ce, ce_u) and abs(skl_u.item() - 1031.0 / 6) < 1e-4


# ------------------------------------------------------------------ Philox: published known answers
def test_philox4x32_10_known_answer_vectors():
    """Random123 kat_vectors (Salmon et al.): philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(*[[c] for c in ctr], *key)
        assert tuple(int(g[0]) for g in got) == want


def test_philox_normal_stream():
    z = philox.philox_normal(400001, seed=11, layer_id=3, sample_id=9)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3
    assert np.array_equal(z[:1000], philox.philox_normal(1000, 11, 3, 9))       # prefix-stable
    assert not np.array_equal(z[:1000], philox.philox_normal(1000, 11, 3, 10))  # disjoint per sample id
    assert not np.array_equal(z[:1000], philox.philox_normal(1000, 11, 4, 9))   # ... and per layer


# ------------------------------------------------------------------ golden fixtures from the reference's own code
@pytest.fixture(scope="module")
def gold():
    if not GOLD.exists():
        pytest.skip("tests/golden/reference_small.pt missing (python oracle/make_golden.py)")
    return torch.load(GOLD, weights_only=False)


@pytest.fixture(scope="module")
def oracle_model(gold):
    torch.manual_seed(gold["seed_w"])
    # reference define_models builds the three unimodal nets first, then the three trunks (RNG order)
    models = O.define_models(gold["C"], seed=None, unimodal=True)
    return models


def test_oracle_topology_reproduces_reference_logits(gold, oracle_model):
    mm = oracle_model["multimodal_model"]
    assert abs(float(sum(p.double().sum() for p in mm.parameters())) - gold["param_checksum"]) < 1e-6
    img, bathy, sss, _ = O.synthetic_batch(gold["B"], seed=gold["seed_x"], size=gold["size"])
    eps = O.draw_eps(mm, gold["S"], gold["seed_eps"])
    sd = {k: v.clone() for k, v in mm.state_dict().items()}
    lg = O.mc_logits(mm, (img, bathy, sss), gold["S"], eps)
    assert torch.allclose(lg, gold["logits_fp32"], rtol=1e-4, atol=1e-6)
    assert abs(O.get_kl_loss(mm).item() - gold["kl"]) < 1e-4 * gold["kl"]
    assert torch.allclose(mm.image_model_feat.bn1.running_mean, gold["bn1_running_mean_after"], atol=1e-7)
    mm.load_state_dict(sd)


def test_oracle_eval_stats_reproduce_reference_csv(gold):
    lg = gold["logits_fp32"]
    st = O.multimodal_eval_stats(lg)
    row = gold["eval_mm_csv_row"]
    hdr = gold["eval_mm_csv_header"]
    assert hdr[:6] == ["Epoch", "Model Type", "Test Loss", "Test Accuracy", "Predictive Uncertainty", "Model Uncertainty"]
    acc = (st["predicted"] == gold["labels"]).float().mean().item()
    assert abs(acc - float(row[3])) < 1e-6 and abs(acc - gold["eval_mm_accuracy"]) < 1e-6
    assert abs(st["predictive_uncertainty"].mean().item() - float(row[4])) < 1e-5
    assert abs(st["model_uncertainty"].mean().item() - float(row[5])) < 1e-5
    kl_scaled = gold["kl"] / 1 * O.kl_weight(0, 20)           # / len(dataloader) = 1   (multimodal.py:293)
    ce = F.cross_entropy(st["output_mean"], gold["labels"]).item()
    assert abs(float(row[6]) - kl_scaled) < 1e-6 * max(1.0, kl_scaled) and abs(float(row[7]) - ce) < 1e-5
    assert abs(float(row[2]) - (ce + kl_scaled)) < 1e-5


def test_oracle_unimodal_stats_reproduce_reference_csv(gold):
    lg = gold["uni_logits_fp32"]
    st = O.unimodal_eval_stats(lg)
    row = gold["eval_uni_csv_row"]
    acc = (st["predicted"] == gold["labels"]).float().mean().item()
    assert abs(acc - float(row[3])) < 1e-6
    assert abs(st["predictive_uncertainty"].mean().item() - float(row[4])) < 1e-6
    assert abs(st["aleatoric_uncertainty"].mean().item() - float(row[5])) < 1e-5


def test_oracle_train_step_reproduces_reference(gold, oracle_model):
    """One ELBO step (CE(mean_s logits) + KL/B * 2^(e+1)/2^E, Adam) restated with the oracle == reference driver."""
    mm = oracle_model["multimodal_model"]
    img, bathy, sss, labels = O.synthetic_batch(gold["B"], seed=gold["seed_x"], size=gold["size"])
    eps = O.draw_eps(mm, gold["S"], gold["seed_eps"])
    opt = torch.optim.Adam(mm.parameters(), lr=1e-4)
    mm.train()
    outs, kls = [], []
    for s in range(gold["S"]):
        O.inject_eps(mm, eps, s)
        outs.append(mm(img, bathy, sss))
        kls.append(O.get_kl_loss(mm))
    O.inject_eps(mm, None, 0)
    loss, ce, skl = O.elbo_loss_multimodal(torch.stack(outs), labels, torch.mean(torch.stack(kls), dim=0),
                                           gold["B"], epoch=1, total_num_epochs=20)
    loss.backward()
    opt.step()
    row = gold["train_mm_csv_row"]
    assert abs(float(row[5]) - skl.item()) < 1e-6 * max(1.0, abs(skl.item()))
    assert abs(float(row[6]) - ce.item()) < 1e-4
    assert abs(gold["train_mm_return"][0] - loss.item() / gold["B"]) < 1e-4      # loss / total samples (multimodal.py:170)
    sd = mm.state_dict()
    for k, v in gold["train_mm_after"].items():
        assert torch.allclose(sd[k].flatten()[:8], v, rtol=1e-3, atol=1e-7), k
