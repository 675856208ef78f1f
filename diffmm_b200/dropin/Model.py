import _path  # noqa: F401
from diffmm_b200.Model import Denoise, GaussianDiffusion, GCNOutput, Model, init  # noqa: F401
