"""Fused set-abstraction / feature-propagation path (filled in once the MLP kernels land)."""


def sa_supported(module, xyz, features):
    return False


def fp_supported(module, unknown, known, unknow_feats, known_feats):
    return False


def sa_forward(module, xyz, features, inds):
    raise NotImplementedError


def fp_forward(module, unknown, known, unknow_feats, known_feats):
    raise NotImplementedError
