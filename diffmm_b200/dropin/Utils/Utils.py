# The reference's Utils/Utils.py has no __all__, so ``from Utils.Utils import *`` (Main.py:10, Model.py:7)
# also exports torch, F, Tensor and np; Main.py:321 relies on ``F`` arriving this way.
import sys, os  # noqa: E401
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _path  # noqa: F401,E402
import numpy as np  # noqa: F401,E402
import torch  # noqa: F401,E402
import torch.nn.functional as F  # noqa: F401,E402
from torch import Tensor  # noqa: F401,E402
from diffmm_b200.Utils.Utils import InfoNCE, bpr_loss, l2_reg_loss  # noqa: F401,E402
