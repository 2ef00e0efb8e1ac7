"""``import deepsensor.torch`` (train.py:19): upstream selects the torch backend by side effect.  torch is the only
backend here; importing this module loads the package and checks that the CUDA library can be opened when a GPU is
present (so a missing build fails at import, not at the first batch)."""
import torch as _torch

import deepsensornz_b200 as _pkg  # noqa: F401
from deepsensornz_b200 import _cabi as _cabi
from deepsensornz_b200 import ConvNP  # noqa: F401  (upstream re-exports the model classes here as well)

if _torch.cuda.is_available():
    _cabi.lib()
