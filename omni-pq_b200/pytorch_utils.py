"""Parameter containers for the shared MLPs of the PointNet++ path.

Drop-in for the reference's `pointnet2/pytorch_utils.py` (SharedMLP :11-36, BatchNorm* :39-64,
Conv1d/2d/3d :67-247, FC :250-282, BNMomentumScheduler :285-end): same class names, constructor
arguments, child-module names and initialisation, hence the same `state_dict` keys
(`layer{i}.conv.weight`, `layer{i}.bn.bn.{weight,bias,running_mean,running_var,num_batches_tracked}`).

These classes only *hold* parameters and define the unfused, layer-by-layer semantics (they are plain
`nn.Sequential`s and can still be called).  On the hot path `pointnet2_modules.PointnetSAModuleVotes`
and `PointnetFPModule` never call them: they hand the conv weights and BatchNorm tensors of a
`SharedMLP` to the fused kernels in libpn2_b200.so (see `fused.py`).
"""
import torch.nn as nn

_CONV = {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}
_NORM = {1: nn.BatchNorm1d, 2: nn.BatchNorm2d, 3: nn.BatchNorm3d}


class _BNBase(nn.Sequential):
    """A one-child Sequential whose child is called `bn` (so keys read `...bn.bn.weight`)."""

    def __init__(self, in_size, batch_norm=None, name=""):
        super().__init__()
        norm = batch_norm(in_size)
        nn.init.ones_(norm.weight)
        nn.init.zeros_(norm.bias)
        self.add_module(name + "bn", norm)


def _bn_class(dim):
    def __init__(self, in_size, *, name=""):
        _BNBase.__init__(self, in_size, batch_norm=_NORM[dim], name=name)

    return type(f"BatchNorm{dim}d", (_BNBase,), {"__init__": __init__, "__doc__": f"BatchNorm{dim}d wrapper"})


BatchNorm1d, BatchNorm2d, BatchNorm3d = _bn_class(1), _bn_class(2), _bn_class(3)
_BN_WRAP = {1: BatchNorm1d, 2: BatchNorm2d, 3: BatchNorm3d}


class _ConvBase(nn.Sequential):
    """[bn, act,] conv [, bn, act]: pre- or post-activation unit.  The conv has no bias when bn."""

    def __init__(self, in_size, out_size, kernel_size, stride, padding, activation, bn, init, conv=None,
                 batch_norm=None, bias=True, preact=False, name=""):
        super().__init__()
        unit = conv(in_size, out_size, kernel_size=kernel_size, stride=stride, padding=padding,
                    bias=bias and not bn)
        init(unit.weight)
        if unit.bias is not None:
            nn.init.zeros_(unit.bias)
        tail = []
        if bn:
            tail.append((name + "bn", batch_norm(in_size if preact else out_size)))
        if activation is not None:
            tail.append((name + "activation", activation))
        order = tail + [(name + "conv", unit)] if preact else [(name + "conv", unit)] + tail
        for key, mod in order:
            self.add_module(key, mod)


def _conv_class(dim):
    ones, zeros = (1,) * dim, (0,) * dim
    if dim == 1:
        ones, zeros = 1, 0

    def __init__(self, in_size, out_size, *, kernel_size=ones, stride=ones, padding=zeros,
                 activation=nn.ReLU(inplace=True), bn=False, init=nn.init.kaiming_normal_, bias=True,
                 preact=False, name=""):
        _ConvBase.__init__(self, in_size, out_size, kernel_size, stride, padding, activation, bn, init,
                           conv=_CONV[dim], batch_norm=_BN_WRAP[dim], bias=bias, preact=preact, name=name)

    return type(f"Conv{dim}d", (_ConvBase,), {"__init__": __init__, "__doc__": f"Conv{dim}d + BN + activation"})


Conv1d, Conv2d, Conv3d = _conv_class(1), _conv_class(2), _conv_class(3)


class SharedMLP(nn.Sequential):
    """Stack of 1x1 Conv2d(+BN+ReLU) units named `layer0..`; widths given by `args`."""

    def __init__(self, args, *, bn=False, activation=nn.ReLU(inplace=True), preact=False, first=False, name=""):
        super().__init__()
        for i in range(len(args) - 1):
            plain = first and preact and i == 0  # a pre-activated first layer gets neither bn nor act
            self.add_module(name + f"layer{i}",
                            Conv2d(args[i], args[i + 1], bn=bn and not plain,
                                   activation=None if plain else activation, preact=preact))


class FC(nn.Sequential):
    def __init__(self, in_size, out_size, *, activation=nn.ReLU(inplace=True), bn=False, init=None,
                 preact=False, name=""):
        super().__init__()
        fc = nn.Linear(in_size, out_size, bias=not bn)
        if init is not None:
            init(fc.weight)
        if fc.bias is not None:
            nn.init.zeros_(fc.bias)
        tail = []
        if bn:
            tail.append((name + "bn", BatchNorm1d(in_size if preact else out_size)))
        if activation is not None:
            tail.append((name + "activation", activation))
        for key, mod in (tail + [(name + "fc", fc)] if preact else [(name + "fc", fc)] + tail):
            self.add_module(key, mod)


def set_bn_momentum_default(bn_momentum):
    def fn(m):
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)):
            m.momentum = bn_momentum

    return fn


class BNMomentumScheduler(object):
    def __init__(self, model, bn_lambda, last_epoch=-1, setter=set_bn_momentum_default):
        if not isinstance(model, nn.Module):
            raise RuntimeError("Class '{}' is not a PyTorch nn Module".format(type(model).__name__))
        self.model, self.setter, self.lmbd = model, setter, bn_lambda
        self.step(last_epoch + 1)
        self.last_epoch = last_epoch

    def step(self, epoch=None):
        if epoch is None:
            epoch = self.last_epoch + 1
        self.last_epoch = epoch
        self.model.apply(self.setter(self.lmbd(epoch)))
