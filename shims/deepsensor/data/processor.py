"""deepsensor.data.processor (preprocess.py:23): the xarray normaliser is ETL, out of scope."""


class DataProcessor:
    def __init__(self, *a, **k):
        raise ImportError("deepsensor.data.processor.DataProcessor is DeepSensor's xarray ETL, which deepsensornz_b200 does "
                          "not replace (only the ConvNP hot path is); install upstream deepsensor for preprocessing")
