import sys, os  # noqa: E401
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _path  # noqa: F401,E402
from diffmm_b200.Utils.Log import Log  # noqa: F401,E402
