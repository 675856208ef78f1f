import _path  # noqa: F401
from diffmm_b200.Conf import *  # noqa: F401,F403
from diffmm_b200.Conf import BaseConfig, Config, DataConfig, HyperConfig, TrainConfig, load_config  # noqa: F401
