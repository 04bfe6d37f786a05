"""mauv — B200-native Monte-Carlo Bayesian hot path of Multimodal-AUV (drop-in Python surface).

    mauv.bayesian      dnn_to_bnn / get_kl_loss / *Reparameterization   (replaces bayesian_torch 0.5.0)
    mauv.models        base_models.py / model_utils.py of the reference (same classes, same signatures)
    mauv.inference     multimodal_predict_and_save                        (inference/predictors.py)
    mauv.train         train_/evaluate_ multimodal / unimodal            (train/multimodal.py, train/unimodal.py)
    mauv.engine        the S-batched execution plan those drivers use
    mauv.ops           tensor-level wrappers over the C-ABI (include/mauv_b200.h)

All arithmetic on the path runs in libmauv_b200.so (hand-written sm_100a CUDA). No CPU fallback.
"""
from . import _lib  # noqa: F401
from .bayesian import (Conv2dReparameterization, LinearReparameterization, dnn_to_bnn, get_kl_loss,  # noqa: F401
                       manual_seed)

__version__ = "0.1.0"
