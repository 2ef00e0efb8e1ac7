"""``import lab as B`` (train.py:12): the two calls nzdownscale makes -- ``B.to_numpy`` (train.py:370) and ``B.sigmoid``
(train.py:648) -- for torch tensors / numpy arrays."""
import numpy as _np
import torch as _torch


def to_numpy(x):
    if isinstance(x, _torch.Tensor):
        return x.detach().cpu().numpy()
    if isinstance(x, (list, tuple)):
        return type(x)(to_numpy(v) for v in x)
    return _np.asarray(x)


def sigmoid(x):
    if isinstance(x, _torch.Tensor):
        return _torch.sigmoid(x)
    x = _np.asarray(x)
    return 1.0 / (1.0 + _np.exp(-x))
