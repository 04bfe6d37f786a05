"""Drop-in for the reference's train/unimodal.py: `train_unimodal_model` (:21-175) and
`evaluate_unimodal_model` (:178-365) for the single-branch ResNet50Custom (image / bathy / sss), same
signatures, CSV columns and return values ((train_accuracy, train_loss); accuracy)."""
from __future__ import annotations

import csv
import logging
import os
from typing import Optional

import numpy as np
import torch
from torch import nn

from .. import ops
from ..bayesian import get_kl_loss
from ..engine import MCEngine
from .multimodal import _confusion_matrix_png, _save_model, ddp_sync_group, train_engine_for


def _pick_input(batch, device, model_type):
    inputs = batch["main_image"].to(device, non_blocking=True)
    labels = batch["label"].long().to(device, non_blocking=True)
    if model_type == "image":
        return inputs, labels
    if model_type == "sss":
        return batch["sss_image"].to(device, non_blocking=True), labels
    if model_type == "bathy":
        return batch["bathy_image"].to(device, non_blocking=True), labels
    logging.error(f"Unknown model_type: {model_type}")
    raise ValueError(f"Unknown model_type: {model_type}")


def train_unimodal_model(model: nn.Module, dataloader, criterion: nn.Module, optimizer: torch.optim.Optimizer,
                         epoch: int, total_num_epochs: int, num_mc: int, sum_writer, device: torch.device,
                         model_type: str = "image", csv_path: str = "", patch_type: Optional[str] = None):
    model.train()
    model.to(device)
    kl_weight = (2 ** (epoch + 1)) / (2 ** total_num_epochs)
    file_exists = os.path.isfile(csv_path)
    try:
        with open(csv_path, mode="a", newline="") as csvfile:
            writer = csv.writer(csvfile)
            if not file_exists:
                writer.writerow(["Epoch", "Model type", "Loss", "Accuracy", "lr"])
            total_loss, correct, total = 0, 0, 0
            module = model.module if isinstance(model, (nn.parallel.DistributedDataParallel, nn.DataParallel)) else model
            engine = train_engine_for(module, criterion)
            sync_group = ddp_sync_group(model, engine)
            for i, batch in enumerate(dataloader):
                logging.info(f"Train batch {i+1}/{len(dataloader)} - Model: {model_type}")
                model_input, labels = _pick_input(batch, device, model_type)
                if engine is not None:
                    engine.zero_grad()
                    res = engine.step((model_input,), labels, num_mc, kl_weight / dataloader.batch_size)
                    output, loss = res["mean_logit"], res["loss"]
                    if sync_group is not False:
                        engine.allreduce_grads(sync_group)       # DDP's reducer is bypassed by the engine (see ddp_sync_group)
                    # The reference's unimodal loop has no NaN/Inf guard (train/unimodal.py:140-142). The engine's
                    # gradients travel in scaled fp16, so one overflow would poison the Adam moments for good: skip the
                    # update for such a batch (what train/multimodal.py:141-145 does) instead of applying it.
                    if not bool(engine.optimizer_step(optimizer)):
                        logging.warning("Skipping optimizer step due to NaN/Inf gradients")
                else:
                    optimizer.zero_grad()
                    outputs = [model(model_input) for _ in range(num_mc)]
                    kl = get_kl_loss(model)                     # eps-independent: S identical terms in the reference
                    output = torch.mean(torch.stack(outputs), dim=0)
                    scaled_kl = kl / dataloader.batch_size
                    cross_entropy_loss = criterion(output, labels)
                    loss = cross_entropy_loss + (kl_weight * scaled_kl)
                    loss.backward()
                    optimizer.step()
                _, predicted = output.float().max(1)
                total_loss += loss.item()
                correct += (predicted == labels).sum().item()
                total += labels.size(0)
                sum_writer.add_scalar("Loss/train", loss, i)
            train_accuracy = correct / total
            train_loss = total_loss / total
            lr = optimizer.param_groups[0]["lr"]
            writer.writerow([epoch + 1, model_type, train_loss, train_accuracy, lr])
        if epoch % 5 == 0:
            _save_model(model, csv_path, model_type)
    except Exception:
        _save_model(model, csv_path, model_type)
        logging.error(f"Error at epoch {epoch}", exc_info=True)
        train_accuracy, train_loss = 0.0, 0.0
    return train_accuracy, train_loss


def evaluate_unimodal_model(model: nn.Module, dataloader, device: torch.device, epoch: int, csv_path: str,
                            total_num_epochs: int, num_mc: int, model_type: str = "image",
                            patch_type: Optional[str] = None):
    model.train()
    kl_weight = (2 ** (epoch + 1)) / (2 ** total_num_epochs)
    file_exists = os.path.isfile(csv_path)
    try:
        with open(csv_path, mode="a", newline="") as csvfile:
            writer = csv.writer(csvfile)
            if not file_exists:
                writer.writerow(["Epoch", "Model Type", "Test Loss", "Test Accuracy", "predictive_uncertainty",
                                 "model_uncertainty"])
            from ..engine import eval_precision
            module = model.module if isinstance(model, (nn.parallel.DistributedDataParallel, nn.DataParallel)) else model
            engine = MCEngine(module, precision=eval_precision("unimodal"))       # fp32-class arithmetic (see eval_precision)
            correct, total, total_loss = 0, 0, 0
            all_pu, all_au, all_predicted, all_labels = [], [], [], []
            with torch.no_grad():
                kl_value = get_kl_loss(model).item()
                for i, batch in enumerate(dataloader):
                    model_input, labels = _pick_input(batch, device, model_type)
                    logits = engine.forward_mc((model_input,), num_mc)
                    st = ops.mc_reduce(logits, 1e-7)                                   # epsilon = 1e-7 (:305)
                    C = logits.shape[-1]
                    host = torch.cat([st["mean_logit"], st["argmax_logit"].to(torch.float32).unsqueeze(1),
                                      st["var_mean"].unsqueeze(1), st["aleatoric"].unsqueeze(1)], dim=1).cpu()
                    output_mean, predicted = host[:, :C], host[:, C].to(torch.int64)   # argmax softmax == argmax logits
                    labels_h = labels.cpu()
                    scaled_kl = kl_value / dataloader.batch_size
                    loss = torch.nn.functional.cross_entropy(output_mean, labels_h).item() + kl_weight * scaled_kl
                    total_loss += loss
                    correct += (predicted == labels_h).sum().item()
                    total += labels_h.size(0)
                    all_pu.extend(host[:, C + 1].numpy())
                    all_au.extend(host[:, C + 2].numpy())
                    all_predicted.extend(predicted.numpy())
                    all_labels.extend(labels_h.numpy())
            accuracy = correct / total
            avg_loss = total_loss / total
            avg_pu = np.mean(all_pu) if all_pu else 0.0
            avg_au = np.mean(all_au) if all_au else 0.0
            _confusion_matrix_png(all_labels, all_predicted, csv_path, model_type, epoch)
            writer.writerow([epoch + 1, model_type, avg_loss, accuracy, avg_pu, avg_au])
    except Exception:
        _save_model(model, csv_path, model_type)
        logging.error(f"Error at epoch {epoch}", exc_info=True)
        accuracy = 0.0
    return accuracy
