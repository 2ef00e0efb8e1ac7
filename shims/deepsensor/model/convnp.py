"""deepsensor.model.convnp (train.py:23, validate_ERA.py:6, validate_WRF.py:7)."""
from deepsensornz_b200.convnp import ConvNP, GaussianPrediction  # noqa: F401
