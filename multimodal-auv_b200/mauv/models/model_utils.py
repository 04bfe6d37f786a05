"""models/model_utils.py of the reference (define_models :10-45, load_pretrained_resnet_as_feature_extractor
:52-64) over mauv.bayesian instead of bayesian_torch."""
import logging
from typing import Any, Dict

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
