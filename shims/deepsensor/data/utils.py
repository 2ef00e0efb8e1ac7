"""deepsensor.data.utils (preprocess.py:24): xarray ETL, out of scope -- see deepsensor/data/__init__.py."""
from . import _needs_upstream

construct_x1x2_ds = _needs_upstream("utils.construct_x1x2_ds")
