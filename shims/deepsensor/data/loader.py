"""deepsensor.data.loader (train.py:20)."""
from deepsensornz_b200.loader import InvalidSamplingStrategyError, TaskLoader  # noqa: F401
