"""deepsensornz_b200: a B200-native ConvNP forward/backward behind the DeepSensor API that
deepsensorNZ (nzdownscale) drives.  See DESIGN.md for the path, the boundary and the kernels."""
from .task import Task, Masked, concat_tasks, convert_task_to_nps_args  # noqa: F401
from .model import ConvNPConfig, ConvNPModule, num_params  # noqa: F401
from .convnp import ConvNP  # noqa: F401
from .train import train_epoch, set_gpu_default_device  # noqa: F401

__all__ = ["Task", "Masked", "concat_tasks", "ConvNPConfig", "ConvNPModule", "num_params", "ConvNP",
           "train_epoch", "set_gpu_default_device"]
