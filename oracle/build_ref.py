"""TEST / BASELINE INFRASTRUCTURE ONLY — never imported by the product path.

Recipe for ``oracle/_ref/``: an UNMODIFIED copy of the reference's Python modules (sun2ot/DiffMM), taken from where
they lie under /root/reference, so that the real reference implementation can be timed as the CPU arm of bench.py
(``--impl reference`` and ``cpu_baseline``) and executed by the compat-mode test on machines where /root/reference
does not exist (the GPU box: oracle/_ref/ is git-ignored but travels with the gpurun snapshot, exactly like the built
``.so``).  Nothing under oracle/_ref/ is ever committed; the reference has no build step, so "building" it is this copy.

    python oracle/build_ref.py            # idempotent; no-op when /root/reference is absent

Also copies the small real datasets shipped with the reference (Datasets/tiktok, 4.5 MB; Datasets/baby interactions)
for the real-data parity runs.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("DIFFMM_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
MODULES = ["Conf.py", "DataHandler.py", "Main.py", "Model.py", "Params.py", os.path.join("Utils", "Utils.py"),
           os.path.join("Utils", "Log.py")]
DATA = [os.path.join("Datasets", "tiktok"), os.path.join("Datasets", "baby")]


def build(verbose: bool = True) -> bool:
    """Returns True when oracle/_ref holds the reference afterwards."""
    if not os.path.isfile(os.path.join(SRC, "Model.py")):
        return os.path.isfile(os.path.join(DST, "Model.py"))
    for rel in MODULES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.isfile(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not (os.path.isfile(d) and filecmp.cmp(s, d, shallow=False)):
            shutil.copyfile(s, d)
    for rel in DATA:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.isdir(s):
            continue
        os.makedirs(d, exist_ok=True)
        for f in sorted(os.listdir(s)):
            if f.startswith("."):
                continue
            if not (os.path.isfile(os.path.join(d, f)) and filecmp.cmp(os.path.join(s, f), os.path.join(d, f), shallow=False)):
                shutil.copyfile(os.path.join(s, f), os.path.join(d, f))
    if verbose:
        print(f"oracle/_ref: reference modules + datasets from {SRC}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
