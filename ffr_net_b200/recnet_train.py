"""Training-mode RecNet (batch-statistics BatchNorm, CosFace head, autograd) — under construction.

The eval/no-grad forward lives in recnet.py and runs fully on the sm_100a library. Until the backward kernels
(dgrad/wgrad implicit GEMMs, BN/PReLU/reflection backward) land, the training entry points fail loudly rather than
fall back to another implementation."""


def forward_train(model, input, label):
    raise NotImplementedError(
        "ffr_net_b200.RecNet: training-mode / label forward is not implemented yet in the sm_100a library "
        "(eval forward with label=None is); there is deliberately no PyTorch fallback")


def cosine_sim(x1, x2, dim=1):
    raise NotImplementedError("ffr_net_b200.cosine_sim: not implemented yet in the sm_100a library")


def self_similarity(x):
    raise NotImplementedError("ffr_net_b200.selfSimilarity: not implemented yet in the sm_100a library")
