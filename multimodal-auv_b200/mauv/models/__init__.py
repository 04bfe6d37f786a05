from .base_models import AdditiveAttention, Identity, MultiModalModel, ResNet50Custom  # noqa: F401
from .model_utils import define_models, load_pretrained_resnet_as_feature_extractor  # noqa: F401
