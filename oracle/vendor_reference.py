#!/usr/bin/env python
"""ORACLE TOOLING (test infrastructure, NOT product code).

Makes the REAL reference importable on the GPU box: copies the Python package `/root/reference/lightx2v` (sources only, unmodified) to
`baseline/_ref/lightx2v`.  `baseline/_ref/` is git-ignored (reference sources never enter the history) but not gpurun-ignored, so it
travels to the GPU box with the snapshot, like the built `.so`.  Run by `__graft_entry__.build()` whenever
`/root/reference` exists; a no-op elsewhere.  The tests and bench legs that use it (`oracle/ref_loader.py`) skip / fall back to the
pinned restatement when `baseline/_ref` is absent.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("LIGHTX2V_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def vendor() -> bool:
    src_pkg = os.path.join(SRC, "lightx2v")
    if not os.path.isdir(src_pkg):
        return False
    dst_pkg = os.path.join(DST, "lightx2v")
    if os.path.isdir(dst_pkg):
        shutil.rmtree(dst_pkg)
    shutil.copytree(src_pkg, dst_pkg, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.sh"))
    with open(os.path.join(DST, "VENDORED_FROM"), "w") as f:
        f.write(f"{src_pkg} (unmodified copy made by oracle/vendor_reference.py; git-ignored)\n")
    return True


if __name__ == "__main__":
    ok = vendor()
    print("vendored" if ok else f"{SRC} not present: nothing to do", DST)
    sys.exit(0)
