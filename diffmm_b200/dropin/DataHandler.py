import _path  # noqa: F401
from diffmm_b200.DataHandler import DataHandler, DiffusionData, DiffusionLoader, TestData, TrainData  # noqa: F401
