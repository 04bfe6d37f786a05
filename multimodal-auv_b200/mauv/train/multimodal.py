"""Drop-in for the reference's train/multimodal.py: `train_multimodal_model` (:25-202) and
`evaluate_multimodal_model` (:204-369) with the same signatures, CSV columns and return values.

Evaluation runs the S MC passes S-batched through engine.MCEngine, the statistics (mean logits, argmax,
H[mean p], mean H[p], MI) in the K5 kernel and the KL sum in the K4 kernel; the only host work per batch is one
[B, 3+C] device->host copy (the reference synchronises per element). Training keeps the reference's loop
structure (S stochastic passes, ELBO = CE(mean logits) + mean(KL)/batch_size * 2^(e+1)/2^E, finite-grad guard, Adam)
over the drop-in Bayesian layers, whose forward/backward dispatch to the CUDA kernels.
"""
from __future__ import annotations

import csv
import logging
import os
from pathlib import Path
from typing import Optional, Tuple

import numpy as np
import torch
from torch import nn

from .. import ops
from ..bayesian import get_kl_loss
from ..engine import MCEngine


def _save_model(model, csv_path, tag):
    """reference train/checkpointing.py:7-44 layout: <dirname(dirname(csv))>/models/bayesian_model_type{tag}.pth"""
    try:
        base = os.path.dirname(os.path.dirname(csv_path))
        d = os.path.join(base, "models")
        os.makedirs(d, exist_ok=True)
        torch.save(model.state_dict(), os.path.join(d, f"bayesian_model_type{tag}.pth"))
    except Exception as e:  # checkpointing must never kill the loop (reference behaviour)
        logging.error(f"save_model failed: {e}")


def _select_patches(batch, device, bathy_patch_type, sss_patch_type):
    """reference train/multimodal.py:87-102: base tensors + optional patch dicts, fallback to the base image."""
    inputs = batch["main_image"].to(device, non_blocking=True)
    labels = batch["label"].long().to(device, non_blocking=True)
    bathy_tensor = batch["bathy_image"].to(device, non_blocking=True)
    sss_image = batch["sss_image"].to(device, non_blocking=True)
    patch_bathy = {k: v.to(device) for k, v in batch.get("patch_bathy", {}).items()}
    patch_sss = {k: v.to(device) for k, v in batch.get("patch_sss", {}).items()}
    patch_bathy["patch_30_bathy"] = bathy_tensor
    patch_sss["patch_30_sss"] = sss_image
    return inputs, labels, patch_bathy.get(bathy_patch_type, bathy_tensor), patch_sss.get(sss_patch_type, sss_image)


def _patch_sizes(bathy_patch_type, sss_patch_type):
    sss = sss_patch_type.replace("patch_", "").replace("_sss", "") if sss_patch_type else "none"
    bathy = bathy_patch_type.replace("patch_", "").replace("_bathy", "") if bathy_patch_type else "none"
    return sss, bathy


def train_multimodal_model(multimodal_model: nn.Module, dataloader, criterion: nn.Module,
                           optimizer: torch.optim.Optimizer, epoch: int, device: torch.device, model_type: str,
                           total_num_epochs: int, num_mc: int, sum_writer, bathy_patch_type: Optional[str] = None,
                           sss_patch_type: Optional[str] = None, csv_path: Optional[str] = "") -> Tuple[float, float]:
    multimodal_model.train()
    csv_path = str(Path(csv_path))
    file_exists = os.path.isfile(csv_path)
    sss_patch_size, bathy_patch_size = _patch_sizes(bathy_patch_type, sss_patch_type)
    try:
        with open(csv_path, mode="a", newline="") as csvfile:
            csv_writer = csv.writer(csvfile)
            if not file_exists:
                csv_writer.writerow(["Epoch", "Model type", "Loss", "Accuracy", "lr", "kl loss", "cross entropy loss",
                                     "SSS Patch Type", "Channel Patch Type"])
            total_loss, correct, total = 0, 0, 0
            kl_weight = (2 ** (epoch + 1)) / (2 ** total_num_epochs)
            module = multimodal_model.module if isinstance(
                multimodal_model, (nn.parallel.DistributedDataParallel, nn.DataParallel)) else multimodal_model
            engine = train_engine_for(module, criterion)
            sync_group = ddp_sync_group(multimodal_model, engine)
            for i, batch in enumerate(dataloader):
                logging.info(f"Train batch {i+1}/{len(dataloader)} - Model: {model_type}")
                inputs, labels, bathy_patch, sss_patch = _select_patches(batch, device, bathy_patch_type, sss_patch_type)
                if engine is not None:
                    # S-batched step (train_engine.py): one grouped forward + one grouped backward for all num_mc passes
                    res = engine.step((inputs, bathy_patch, sss_patch), labels, num_mc, kl_weight / dataloader.batch_size)
                    if sync_group is not False:
                        # the engine writes `.grad` directly, so DistributedDataParallel's reducer never sees these
                        # gradients: average them over the ranks here (one all-reduce of the flat buffer), BEFORE the
                        # finite check and the optimizer step, as DDP's backward hooks would have
                        engine.allreduce_grads(sync_group)
                    output, cross_entropy_loss, loss = res["mean_logit"], res["ce"], res["loss"]
                    scaled_kl = res["kl"] / dataloader.batch_size * kl_weight
                    if not bool(torch.isfinite(loss)):
                        logging.warning(f"Skipping batch {i} due to NaN/Inf loss: {loss}")
                        engine.zero_grad()       # the engine has already accumulated this batch's gradients: drop them
                        continue
                    # NaN/Inf guard + Adam in one device pass (mauv.optim.FusedAdam; reference :141-145)
                    if bool(engine.optimizer_step(optimizer)):
                        engine.zero_grad()
                    else:
                        logging.warning("Skipping optimizer step due to NaN/Inf gradients")
                else:
                    output_ensemble = [module(inputs, bathy_patch, sss_patch) for _ in range(num_mc)]
                    # KL(q||p) does not depend on eps: the reference's S get_kl_loss calls return the same value S times
                    # (train/multimodal.py:114,124 mean over them); one fused launch computes it once.
                    kl = get_kl_loss(module)
                    output = torch.mean(torch.stack(output_ensemble), dim=0)
                    scaled_kl = kl / dataloader.batch_size * kl_weight
                    cross_entropy_loss = criterion(output, labels)
                    loss = cross_entropy_loss + scaled_kl
                    if torch.any(torch.isnan(loss)) or torch.any(torch.isinf(loss)):
                        logging.warning(f"Skipping batch {i} due to NaN/Inf loss: {loss}")
                        continue
                    loss.backward()
                    if not any(torch.any(torch.isnan(p.grad)) or torch.any(torch.isinf(p.grad))
                               for p in multimodal_model.parameters() if p.grad is not None):
                        optimizer.step()
                        optimizer.zero_grad()
                    else:
                        logging.warning("Skipping optimizer step due to NaN/Inf gradients")
                total_loss += loss.item()
                _, predicted = torch.max(output, 1)
                correct += (predicted == labels).sum().item()
                total += labels.size(0)
                sum_writer.add_scalar("Loss/train", loss, i)
                logging.info(f"[Epoch {epoch} | Batch {i}] Loss: {loss.item():.4f}, KL: {scaled_kl.item():.4f}, "
                             f"Accuracy: {correct / total:.4f}")
            train_accuracy = correct / total
            train_loss = total_loss / total
            lr = optimizer.param_groups[0]["lr"]
            csv_writer.writerow([epoch, model_type, train_loss, train_accuracy, lr, scaled_kl.item(),
                                 cross_entropy_loss.item(), sss_patch_size, bathy_patch_size])
        if epoch % 5 == 0:
            _save_model(multimodal_model, csv_path, f"{model_type}_bathy_patch{bathy_patch_size}_sss_patch{sss_patch_size}")
    except Exception:
        _save_model(multimodal_model, csv_path, f"{model_type}_bathy_patch{bathy_patch_size}_sss_patch{sss_patch_size}")
        logging.error(f"Error at epoch {epoch}", exc_info=True)
        train_loss, train_accuracy = 0.0, 0.0
    return train_loss, train_accuracy


def ddp_sync_group(model: nn.Module, engine):
    """The process group over which the S-batched engine's gradients must be averaged, or False when no exchange is
    needed. The reference's loops rely on the wrapper (`model(x)` through DistributedDataParallel synchronises gradients
    in backward); TrainEngine bypasses the wrapper's forward, so the drop-in loops do the exchange themselves whenever the
    model they were handed is DDP-wrapped and torch.distributed runs with more than one rank. nn.DataParallel (what
    reference utils/device.py:19 uses) is single-process: the engine runs the whole minibatch on the module's device."""
    if engine is None or not isinstance(model, nn.parallel.DistributedDataParallel):
        return False
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(model.process_group) < 2:
        return False
    if engine._flat is None:
        engine.flatten_grads()
    return model.process_group


def train_engine_for(module: nn.Module, criterion) -> Optional["TrainEngine"]:
    """The S-batched TrainEngine for `module` (cached on it), or None when the step cannot be expressed by it: a criterion
    other than plain mean-reduced nn.CrossEntropyLoss, a topology other than the reference's two model classes, or
    MAUV_TRAIN_PATH=layers. None selects the drop-in layer path (same CUDA kernels, one pass at a time)."""
    from ..train_engine import TrainEngine
    from .._lib import MauvError
    if os.environ.get("MAUV_TRAIN_PATH", "engine") == "layers":
        return None
    plain_ce = (type(criterion) is nn.CrossEntropyLoss and criterion.weight is None and criterion.reduction == "mean"
                and getattr(criterion, "label_smoothing", 0.0) == 0.0 and criterion.ignore_index == -100)
    if not plain_ce:
        return None
    eng = module.__dict__.get("_mauv_train_engine")
    if eng is None:
        try:
            eng = TrainEngine(module)
        except MauvError:
            return None
        module.__dict__["_mauv_train_engine"] = eng
    return eng


def evaluate_multimodal_model(multimodal_model: nn.Module, dataloader, device: torch.device, epoch: int,
                              total_num_epochs: int, num_mc: int, model_type: str,
                              bathy_patch_type: Optional[str] = None, sss_patch_type: Optional[str] = None,
                              csv_path: Optional[str] = ""):
    multimodal_model.train()  # BN batch statistics per MC pass (reference :232)
    csv_path = str(Path(csv_path))
    file_exists = os.path.isfile(csv_path)
    try:
        with open(csv_path, mode="a", newline="") as csvfile:
            csv_writer = csv.writer(csvfile)
            if not file_exists:
                csv_writer.writerow(["Epoch", "Model Type", "Test Loss", "Test Accuracy", "Predictive Uncertainty",
                                     "Model Uncertainty", "Scaled KL", "Cross Entropy Loss", "bathy Patch Type",
                                     "SSS Patch Type"])
            module = multimodal_model.module if isinstance(
                multimodal_model, (nn.parallel.DistributedDataParallel, nn.DataParallel)) else multimodal_model
            from ..engine import eval_precision
            engine = MCEngine(module, precision=eval_precision("multimodal"))
            total_loss, correct, total = 0, 0, 0
            all_predicted, all_labels, all_pu, all_mu = [], [], [], []
            kl_weight = (2 ** (epoch + 1)) / (2 ** total_num_epochs)
            with torch.no_grad():
                kl_value = get_kl_loss(module).item()          # eps-independent: once per evaluation, not S x batches
                for i, batch in enumerate(dataloader):
                    logging.info(f"Eval, batch: {i+1}/{len(dataloader)}, model: {model_type}")
                    inputs, labels, bathy_patch, sss_patch = _select_patches(batch, device, bathy_patch_type,
                                                                             sss_patch_type)
                    logits = engine.forward_mc((inputs, bathy_patch, sss_patch), num_mc)
                    st = ops.mc_reduce(logits, 1e-8)                                     # epsilon = 1e-8 (:257)
                    C = logits.shape[-1]
                    host = torch.cat([st["mean_logit"], st["argmax_logit"].to(torch.float32).unsqueeze(1),
                                      st["pred_entropy"].unsqueeze(1), st["mutual_info"].unsqueeze(1)], dim=1).cpu()
                    output_mean, predicted = host[:, :C], host[:, C].to(torch.int64)
                    labels_h = labels.cpu()
                    kl_scaled = kl_value / len(dataloader) * kl_weight                   # (:293-294)
                    cross_entropy_loss = torch.nn.functional.cross_entropy(output_mean, labels_h).item()
                    total_loss += cross_entropy_loss + kl_scaled
                    correct += (predicted == labels_h).sum().item()
                    total += labels_h.size(0)
                    all_pu.extend(host[:, C + 1].numpy())
                    all_mu.extend(host[:, C + 2].numpy())
                    all_predicted.extend(predicted.numpy())
                    all_labels.extend(labels_h.numpy())
            test_accuracy = correct / total
            test_loss = total_loss / len(dataloader)
            predictive_uncertainty_mean = np.mean(all_pu)
            model_uncertainty_mean = np.mean(all_mu)
            _confusion_matrix_png(all_labels, all_predicted, csv_path, model_type, epoch)
            csv_writer.writerow([epoch + 1, model_type, test_loss, test_accuracy, predictive_uncertainty_mean,
                                 model_uncertainty_mean, kl_scaled, cross_entropy_loss,
                                 bathy_patch_type or "patch_30_bathy", sss_patch_type or "patch_30_sss"])
            logging.info(f"Epoch {epoch + 1}: Test Loss: {test_loss:.4f}, Accuracy: {test_accuracy:.4f}, "
                         f"Total Uncertainty: {predictive_uncertainty_mean:.4f}, Epistemic: {model_uncertainty_mean:.4f}")
    except Exception as e:
        logging.error(f"Critical error at epoch {epoch}: {e}", exc_info=True)
        test_accuracy = 0.0
    return test_accuracy


def _confusion_matrix_png(all_labels, all_predicted, csv_path, model_type, epoch):
    """Reference :322-347 — plotting is outside the hot path and optional (matplotlib may be absent)."""
    fig = None
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        from sklearn.metrics import ConfusionMatrixDisplay, confusion_matrix
        cm = confusion_matrix(all_labels, all_predicted)
        fig, ax = plt.subplots(figsize=(8, 8))
        ConfusionMatrixDisplay(confusion_matrix=cm).plot(cmap="Blues", ax=ax)
        plt.title(f"Confusion Matrix for Epoch {epoch}")
        folder = os.path.join(os.path.dirname(csv_path), "confusion_matrices")
        os.makedirs(folder, exist_ok=True)
        plt.savefig(os.path.join(folder, f"conf_matrix_model_{model_type}_{epoch}.png"))
    except Exception as e:
        logging.warning(f"Confusion matrix not saved due to plotting error: {e}")
    finally:
        if fig is not None:
            import matplotlib.pyplot as plt
            plt.close(fig)
