"""deepsensor.data: Task / TaskLoader are on the hot path; the xarray ETL helpers are not (out of
scope) and fail with a clear message when called."""
from .loader import TaskLoader  # noqa: F401
from .task import Task, concat_tasks  # noqa: F401


def _needs_upstream(name):
    def f(*a, **k):
        raise ImportError(f"deepsensor.data.{name} is part of DeepSensor's xarray ETL, which deepsensornz_b200 does not "
                          "replace (only the ConvNP hot path is: SURVEY.md section 8); install upstream deepsensor for it")
    f.__name__ = name
    return f


construct_circ_time_ds = _needs_upstream("construct_circ_time_ds")   # preprocess.py:25, validate_ERA.py:7
