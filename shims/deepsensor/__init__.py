"""``deepsensor`` namespace shim over deepsensornz_b200 (see shims/README.md).

Mirrors the module paths nzdownscale imports (train.py:19-26, validate_ERA.py:6-7); everything on the ConvNP hot path
is re-exported from deepsensornz_b200, whose arithmetic runs in libconvnp_b200.so."""
from . import backend  # noqa: F401  (train.py:262 reaches deepsensor.backend.nps through the package attribute)

__version__ = "0.3.6+b200"
