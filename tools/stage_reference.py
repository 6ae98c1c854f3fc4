#!/usr/bin/env python
"""Stage the UNMODIFIED reference modules the hot path imports into baseline/_ref/ (git-ignored; travels to the GPU box with
gpurun) so that bench.py's CPU arm times the reference itself, not the oracle port.

    python tools/stage_reference.py          (called by __graft_entry__.build() when /root/reference exists)

Only the files model/nerf.py's import closure needs are taken, byte for byte; three absent third-party modules that are
not on the arithmetic path (h5py, hdf5plugin, imageio[.v3]) are provided as empty stubs under _stubs/.
The reference is not a pip package (no setup.py / pyproject.toml), so `pip install --target baseline/_ref` does not apply.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("BENERF_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["spline.py", "run_nerf_helpers.py", "undistort.py", "model/__init__.py", "model/nerf.py", "model/optimize.py",
         "model/component.py", "model/embedder.py", "utils/__init__.py", "utils/math_utils.py", "utils/img_utils.py",
         "utils/event_utils.py", "loss/__init__.py", "loss/imgloss.py"]
STUBS = {"h5py/__init__.py": "File = object\n", "hdf5plugin/__init__.py": "",
         "imageio/__init__.py": "from . import v3\n", "imageio/v3.py": "def imwrite(*a, **k):\n    return None\n\n\ndef imread(*a, **k):\n    return None\n"}


def stage():
    if not os.path.isdir(REF):
        return False
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(src):
            shutil.copyfile(src, dst)
        elif rel.endswith("__init__.py"):
            open(dst, "w").close()                     # namespace package upstream
        else:
            raise FileNotFoundError(src)
    for rel, body in STUBS.items():
        dst = os.path.join(DST, "_stubs", rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(dst, "w") as f:
            f.write(body)
    return True


def import_staged():
    """Import the staged reference (None when it has not been staged).  The stubs are appended to sys.path so that a real
    h5py / imageio, where installed, wins."""
    if not os.path.exists(os.path.join(DST, "model", "nerf.py")):
        return None
    from argparse import Namespace
    if DST not in sys.path:
        sys.path.insert(0, DST)
    stubs = os.path.join(DST, "_stubs")
    if stubs not in sys.path:
        sys.path.append(stubs)
    import spline, run_nerf_helpers                       # noqa: E401
    from model import nerf, optimize
    from utils import math_utils, img_utils
    from loss import imgloss
    return Namespace(spline=spline, helpers=run_nerf_helpers, nerf=nerf, optimize=optimize, math_utils=math_utils,
                     img_utils=img_utils, imgloss=imgloss)


if __name__ == "__main__":
    print("staged" if stage() else f"{REF} not present: nothing staged")
