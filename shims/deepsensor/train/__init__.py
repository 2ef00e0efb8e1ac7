from .train import set_gpu_default_device, train_epoch  # noqa: F401
