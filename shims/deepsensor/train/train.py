"""deepsensor.train.train (train.py:24, validate.py:28)."""
from deepsensornz_b200.train import set_gpu_default_device, train_epoch  # noqa: F401
