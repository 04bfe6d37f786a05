"""models/model_utils.py of the reference (define_models :10-45, load_pretrained_resnet_as_feature_extractor
:52-64) over mauv.bayesian instead of bayesian_torch."""
import logging
import os
from typing import Any, Dict, Tuple

import torch
import torch.nn as nn

from ..bayesian import dnn_to_bnn
from .base_models import Identity, MultiModalModel, ResNet50Custom, _resnet50_backbone


def load_pretrained_resnet_as_feature_extractor(input_channels: int = 3) -> nn.Module:
    model = _resnet50_backbone()
    if input_channels == 1:
        model.conv1 = nn.Conv2d(1, 64, kernel_size=(7, 7), stride=(2, 2), padding=(3, 3), bias=False)
    model.fc = Identity()
    return model


def define_models(device: torch.device, num_classes: int,
                  const_bnn_prior_parameters: Dict[str, Any]) -> Dict[str, nn.Module]:
    try:
        image_model = ResNet50Custom(input_channels=3, num_classes=num_classes)
        bathy_model = ResNet50Custom(input_channels=3, num_classes=num_classes)
        sss_model = ResNet50Custom(input_channels=1, num_classes=num_classes)
        dnn_to_bnn(image_model, const_bnn_prior_parameters)
        dnn_to_bnn(bathy_model, const_bnn_prior_parameters)
        dnn_to_bnn(sss_model, const_bnn_prior_parameters)
        image_model_feat = load_pretrained_resnet_as_feature_extractor()
        bathy_model_feat = load_pretrained_resnet_as_feature_extractor()
        sss_model_feat = load_pretrained_resnet_as_feature_extractor(input_channels=1)
        multimodal_model = MultiModalModel(image_model_feat, bathy_model_feat, sss_model_feat, num_classes)
        dnn_to_bnn(multimodal_model, const_bnn_prior_parameters)
        return {
            "image_model": image_model, "bathy_model": bathy_model, "sss_model": sss_model,
            "multimodal_model": multimodal_model, "image_model_feat": image_model_feat,
            "bathy_model_feat": bathy_model_feat, "sss_model_feat": sss_model_feat,
        }
    except Exception as e:
        logging.error(f"Error defining models: {e}", exc_info=True)
        raise


def load_models(model_paths: Dict[str, str], device: torch.device,
                num_classes: int) -> Tuple[nn.Module, nn.Module, nn.Module]:
    """models/model_utils.py:66-101 of the reference: the three (deterministic) ResNet-50 feature extractors, each filled
    from `model_paths[key]` (keys "image", "channels", "sss") when that file exists. A missing path is a warning and a
    failing load an error log, never an exception - the reference's behaviour; only a failure to build the trunks raises."""
    try:
        loaded = {"image": load_pretrained_resnet_as_feature_extractor(),
                  "channels": load_pretrained_resnet_as_feature_extractor(),
                  "sss": load_pretrained_resnet_as_feature_extractor(input_channels=1)}
        for key, model in loaded.items():
            path = model_paths.get(key)
            try:
                if path and os.path.exists(path):
                    model.load_state_dict(torch.load(path, map_location=device))
                    logging.info(f"{key.capitalize()} model loaded successfully from {path}.")
                else:
                    logging.warning(f"Path not found for model: {key} -> {path}")
            except Exception as inner_e:
                logging.error(f"Failed to load {key} model from {path}: {inner_e}", exc_info=True)
        return loaded["image"], loaded["channels"], loaded["sss"]
    except Exception as e:
        logging.error(f"Error loading models: {e}", exc_info=True)
        raise


_BRANCHES = ("image_model_feat", "bathy_model_feat", "sss_model_feat")


def load_reference_weights(multimodal_model: nn.Module, weights, num_classes: int = 7, map_location="cpu"):
    """Load a checkpoint written by the reference into the mauv multimodal model: a `.pth` from
    train/checkpointing.py:40 (`torch.save(model.state_dict())`, possibly with the `module.` prefix of DataParallel) or
    the published `pytorch_model.bin`, whose trunks are stored one level deeper (`image_model_feat.model.conv1...`).
    Same key normalisation as Examples/Example_Inference_model.py:82-112: strip `module.`, drop the `.model.` level of the
    three trunks, and leave the output layer (`fc2.*`) at its fresh initialisation when num_classes differs from the 7
    classes the published weights were trained on. The Bayesian layers carry the same parameter names as
    bayesian-torch's (mu_kernel / rho_kernel / mu_weight / rho_weight / mu_bias / rho_bias), so nothing else is remapped.
    -> (missing_keys, unexpected_keys) as `load_state_dict(strict=False)` reports them."""
    if isinstance(weights, (str, bytes, os.PathLike)) or hasattr(weights, "read"):
        state = torch.load(weights, map_location=map_location)
    else:
        state = weights
    if not hasattr(state, "items"):
        raise TypeError(f"load_reference_weights: expected a path, a file object or a state_dict, got {type(weights).__name__}")
    fixed = {}
    for key, value in state.items():
        if key.startswith("module."):
            key = key[len("module."):]
        for branch in _BRANCHES:
            deep = branch + ".model."
            if key.startswith(deep):
                key = branch + "." + key[len(deep):]
                break
        if num_classes != 7 and key.startswith("fc2."):
            logging.info(f"load_reference_weights: skipping '{key}' (checkpoint has 7 classes, model has {num_classes})")
            continue
        fixed[key] = value
    result = multimodal_model.load_state_dict(fixed, strict=False)
    for key in result.missing_keys:
        logging.warning(f"load_reference_weights: missing in the checkpoint: {key}")
    for key in result.unexpected_keys:
        logging.warning(f"load_reference_weights: not used by the model: {key}")
    return list(result.missing_keys), list(result.unexpected_keys)
