"""deepsensor.backend.nps: the one attribute nzdownscale uses is ``num_params`` (train.py:262)."""
from deepsensornz_b200.model import num_params  # noqa: F401
