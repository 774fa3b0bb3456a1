"""TEST INFRASTRUCTURE ONLY -- recipe that stages the UNMODIFIED reference next to the oracle.

The reference (AsteriosPar/mmWave_MSc) is pure Python: there is nothing to compile, but its hot-path modules can
run anywhere numpy / scikit-learn are.  `/root/reference` exists only in the build container, so this recipe copies
the five files of the path from where they lie into `oracle/_ref/src/` -- git-ignored (no reference source enters
the history), NOT gpurun-ignored (it travels to the GPU box like the built libmmw.so).  What uses it there:

  * tests/test_gpu_reference_driver.py  -- the reference's own offline_main.py drives the GPU drop-in modules;
  * bench.py --impl reference           -- the reference's own Utils / Tracking timed on the host cores
                                           (`cpu_baseline.kind` = "reference");
  * oracle/ref_harness.py               -- falls back to this copy when /root/reference is absent.

A manifest with the sha256 of every file is written next to the copies so that "unmodified" can be checked.
Run by `__graft_entry__.build()`; a no-op when the reference is not present (the GPU box).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

SRC = os.environ.get("MMW_REFERENCE_SRC", "/root/reference/src")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "src")
FILES = ["constants.py", "Utils.py", "Tracking.py", "offline_main.py", "ReadDataIWR1443.py"]


def staged_dir() -> str:
    """Directory holding the reference modules: the original when present, else the staged copy, else ''."""
    if os.path.isfile(os.path.join(SRC, "Tracking.py")):
        return SRC
    if os.path.isfile(os.path.join(DST, "Tracking.py")):
        return DST
    return ""


def make_ref() -> bool:
    if not os.path.isfile(os.path.join(SRC, "Tracking.py")):
        return False
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        manifest[f] = hashlib.sha256(open(os.path.join(DST, f), "rb").read()).hexdigest()
    with open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "sha256": manifest}, fh, indent=1)
    return True


if __name__ == "__main__":
    print("staged" if make_ref() else "reference not present at %s: nothing staged" % SRC)
