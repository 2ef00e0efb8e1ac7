"""deepsensor.data.task (train.py:26)."""
from deepsensornz_b200.task import Masked, Task, concat_tasks, convert_task_to_nps_args  # noqa: F401
