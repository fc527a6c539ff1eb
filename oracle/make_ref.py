"""Recipe: stage the UNMODIFIED reference (bryanlimy/V1T) under oracle/_ref/ so it travels to the GPU box.

TEST / BASELINE INFRASTRUCTURE — nothing under v1t_b200/ imports it.  The reference is pure Python (no build
step), so "building" it means copying the package and the two caller scripts, byte for byte, from where they lie:

    /root/reference/src/v1t      ->  oracle/_ref/src/v1t        (Model, PoissonLoss, Recorder, ...)
    /root/reference/train.py     ->  oracle/_ref/train.py       (train_step, train.py:42-81)
    /root/reference/ensemble.py  ->  oracle/_ref/ensemble.py    (EnsembleModel, ensemble.py:30-151)

oracle/_ref/ is git-ignored (the history stays free of reference sources) but NOT gpurun-ignored, so the GPU box
gets it with the snapshot: `bench.py --impl reference`, the `gpu_eager_baseline` leg of the main bench line and the
live-reference tests (tests/test_live_reference.py) then run the reference's own code there.  Run by
__graft_entry__.build() whenever /root/reference is present; a no-op when the copy is already identical.

    python oracle/make_ref.py [--check]
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("V1T_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
ITEMS = [("src/v1t", "src/v1t"), ("train.py", "train.py"), ("ensemble.py", "ensemble.py")]


def _same_tree(a: str, b: str) -> bool:
    if os.path.isfile(a):
        return os.path.isfile(b) and filecmp.cmp(a, b, shallow=False)
    if not os.path.isdir(b):
        return False
    cmp = filecmp.dircmp(a, b, ignore=["__pycache__"])
    if cmp.left_only or cmp.right_only or cmp.funny_files:
        return False
    _, mismatch, errors = filecmp.cmpfiles(a, b, cmp.common_files, shallow=False)
    if mismatch or errors:
        return False
    return all(_same_tree(os.path.join(a, d), os.path.join(b, d)) for d in cmp.common_dirs)


def stage(check_only: bool = False) -> bool:
    """Returns True when oracle/_ref holds an identical copy afterwards."""
    if not os.path.isdir(os.path.join(SRC, "src", "v1t")):
        return os.path.isdir(os.path.join(DST, "src", "v1t"))  # GPU box: use what travelled
    ok = True
    for rel_src, rel_dst in ITEMS:
        a, b = os.path.join(SRC, rel_src), os.path.join(DST, rel_dst)
        if _same_tree(a, b):
            continue
        ok = False
        if check_only:
            continue
        if os.path.isdir(b):
            shutil.rmtree(b)
        elif os.path.exists(b):
            os.remove(b)
        os.makedirs(os.path.dirname(b), exist_ok=True)
        if os.path.isdir(a):
            shutil.copytree(a, b, ignore=shutil.ignore_patterns("__pycache__"))
        else:
            shutil.copy2(a, b)
    return ok or not check_only


if __name__ == "__main__":
    good = stage(check_only="--check" in sys.argv)
    print(f"oracle/_ref {'ok' if good else 'STALE'} ({DST})")
    sys.exit(0 if good else 1)
