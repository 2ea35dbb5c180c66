"""ffr_net_b200 — B200-native (sm_100a) implementation of the FFR-Net hot path.

Drop-in module surface of the reference (haoosz/FFR-Net):
    pretrain.model_ir_se50  ->  ffr_net_b200.backbone   (Backbone, ir_se_50_512, l2_norm)
    models.recnet           ->  ffr_net_b200.recnet     (RecNet, selfSimilarity, cosine_sim, init_weights)
    lfw.lfw_eval (scoring)  ->  ffr_net_b200.scoring    (pair_cosine, KFold, get_fold_accuracy, ...)
Everything below these modules runs in libffr_sm100.so (hand-written CUDA, C ABI in include/ffr_sm100.h).
"""
from .backbone import Backbone, ir_se_50_512, l2_norm  # noqa: F401
from .recnet import RecNet, selfSimilarity, cosine_sim, init_weights  # noqa: F401

__all__ = ["Backbone", "ir_se_50_512", "l2_norm", "RecNet", "selfSimilarity", "cosine_sim", "init_weights"]
