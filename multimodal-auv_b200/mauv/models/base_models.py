"""Network topology of the reference (models/base_models.py:7-90), same class names, constructor
signatures, attribute names and state_dict keys. The only deviation: ImageNet weights are used when
torchvision can load them from the local cache and random init otherwise (no network on the target box;
BASELINE.json specifies random-init weights for every benchmark)."""
import logging

import torch
import torch.nn as nn
import torch.nn.functional as F
from torchvision.models import ResNet50_Weights, resnet50


def _resnet50_backbone() -> nn.Module:
    """torchvision ResNet-50 with the ImageNet checkpoint when it is already in the local torch hub cache
    (the reference downloads it: models/base_models.py:15); random init otherwise - never touches the network."""
    import os
    ckpt = os.path.join(torch.hub.get_dir(), "checkpoints", os.path.basename(ResNet50_Weights.IMAGENET1K_V1.url))
    if os.path.exists(ckpt):
        return resnet50(weights=ResNet50_Weights.IMAGENET1K_V1)
    logging.warning("ImageNet ResNet-50 checkpoint not in the local cache; using random init")
    return resnet50(weights=None)


class ResNet50Custom(nn.Module):
    def __init__(self, input_channels, num_classes):
        super().__init__()
        self.input_channels = input_channels
        self.model = _resnet50_backbone()
        self.model.conv1 = nn.Conv2d(input_channels, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.model.fc = nn.Linear(self.model.fc.in_features, num_classes)

    def forward(self, x):
        return self.model(x)

    def get_feature_size(self):
        return self.model.fc.in_features


class Identity(nn.Module):
    def forward(self, x):
        return x


class AdditiveAttention(nn.Module):
    def __init__(self, d_model, hidden_dim=128):
        super().__init__()
        self.query_projection = nn.Linear(d_model, hidden_dim)
        self.key_projection = nn.Linear(d_model, hidden_dim)
        self.value_projection = nn.Linear(d_model, hidden_dim)
        self.attention_mechanism = nn.Linear(hidden_dim, hidden_dim)

    def forward(self, query):
        keys = self.key_projection(query)
        values = self.value_projection(query)
        queries = self.query_projection(query)
        attention_scores = torch.tanh(queries + keys)
        attention_weights = F.softmax(self.attention_mechanism(attention_scores), dim=1)
        return values * attention_weights


class MultiModalModel(nn.Module):
    def __init__(self, image_model_feat, bathy_model_feat, sss_model_feat, num_classes,
                 attention_type="scaled_dot_product"):
        super().__init__()
        self.image_model_feat = image_model_feat
        self.bathy_model_feat = bathy_model_feat
        self.sss_model_feat = sss_model_feat
        self.fc = nn.Linear(384, 1284)
        self.fc1 = nn.Linear(1284, 32)
        num_classes = int(num_classes)
        self.fc2 = nn.Linear(32, num_classes)
        self.attention_type = attention_type
        self.attention_image = AdditiveAttention(2048)
        self.attention_bathy = AdditiveAttention(2048)
        self.attention_sss = AdditiveAttention(2048)

    def forward(self, inputs, bathy_tensor, sss_image):
        image_features = self.image_model_feat(inputs)
        bathy_features = self.bathy_model_feat(bathy_tensor)
        sss_features = self.sss_model_feat(sss_image)
        image_features_attended = self.attention_image(image_features)
        bathy_features_attended = self.attention_bathy(bathy_features)
        sss_features_attended = self.attention_sss(sss_features)
        combined_features = torch.cat([image_features_attended, bathy_features_attended, sss_features_attended], dim=1)
        outputs_1 = self.fc(combined_features)
        output_2 = self.fc1(outputs_1)
        return self.fc2(output_2)
