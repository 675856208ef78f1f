"""TEST INFRASTRUCTURE ONLY — never imported by the product path.

Imports the *unmodified* reference (sun2ot/DiffMM) on CPU so that golden
vectors can be generated from it (oracle/gen_golden.py), so the numpy oracle
(oracle/diffmm_oracle.py) can be pinned against it, and so bench.py can time
it as the CPU arm.  The modules come from /root/reference where that is
mounted (this container) and otherwise from oracle/_ref/, the git-ignored
unmodified copy made by oracle/build_ref.py that travels to the GPU box with
the snapshot.  Only tests/, smoke() and bench.py's reference / cpu_baseline
legs may import this module.

What the shim does:
  * Conf.py:62-66 uses dataclass-instance defaults, which Python >= 3.11
    rejects; the source is exec'd with ``X: T = T()`` rewritten to
    ``field(default_factory=T)`` and registered as ``sys.modules['Conf']``.
  * Main.py / Model.py call ``.cuda()`` unconditionally (Main.py:88-108,
    Model.py:397); on a CPU-only host these are patched to identity.
  * Main.py reads module globals ``main_log`` and ``config`` (Main.py:23,99).
"""
import dataclasses  # noqa: F401
import os
import re
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root() -> str:
    env = os.environ.get("DIFFMM_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_HERE, "_ref")):
        if os.path.isfile(os.path.join(cand, "Model.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()


class _QuietLog:
    def __init__(self):
        self.lines = []

    def info(self, msg):
        self.lines.append(str(msg))


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "Model.py"))


def load_reference(force_cpu: bool = False):
    """Returns a namespace with the reference modules (Conf, Model, DataHandler, Utils, Main).

    force_cpu: make the reference take its CPU path even on a machine with a GPU (its modules pick
    ``cuda:{gpu}`` whenever torch.cuda.is_available(), Model.py:149,227): torch.cuda.is_available is patched to
    False for the rest of the process — use only in a process that does no GPU work (bench.py --impl reference)."""
    if not available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    import torch
    from torch import nn

    sys.dont_write_bytecode = True
    # make sure the reference's flat modules win over any drop-in dir on sys.path
    for name in ("Conf", "Model", "DataHandler", "Main", "Utils", "Utils.Utils", "Utils.Log"):
        sys.modules.pop(name, None)
    if REFERENCE_ROOT in sys.path:
        sys.path.remove(REFERENCE_ROOT)
    sys.path.insert(0, REFERENCE_ROOT)

    src = "import dataclasses\n" + re.sub(
        r"(\w+): (\w+Config) = \2\(\)",
        r"\1: \2 = dataclasses.field(default_factory=\2)",
        open(os.path.join(REFERENCE_ROOT, "Conf.py")).read(),
    )
    conf = types.ModuleType("Conf")
    conf.__file__ = os.path.join(REFERENCE_ROOT, "Conf.py")
    sys.modules["Conf"] = conf  # must be registered before exec: @dataclass looks the module up
    exec(compile(src, "Conf_shim", "exec"), conf.__dict__)

    if force_cpu:
        torch.cuda.is_available = lambda: False
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self

    import Model as ref_model  # noqa: E402
    import DataHandler as ref_data  # noqa: E402
    import Utils.Utils as ref_utils  # noqa: E402
    import Main as ref_main  # noqa: E402

    ref_main.main_log = _QuietLog()
    ns = types.SimpleNamespace(Conf=conf, Model=ref_model, DataHandler=ref_data,
                               Utils=ref_utils, Main=ref_main)
    return ns


def make_config(ref, name="tiktok", **over):
    """Builds a reference Config; ``over`` keys are 'section.key'."""
    cfg = ref.Conf.Config()
    cfg.data.name = name
    for k, v in over.items():
        sec, key = k.split(".")
        setattr(getattr(cfg, sec), key, v)
    return cfg
