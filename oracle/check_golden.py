#!/usr/bin/env python
"""Re-derives every fixture under tests/golden/ from the reference (authoring container only: needs /root/reference)
and compares the arrays with the committed files -- test infrastructure, like the rest of oracle/.

    python oracle/check_golden.py       # exit 0: every key of every .npz is reproduced bit for bit

The generators write into tests/golden/; the committed files are saved first and restored afterwards, so the work
tree is left as it was (the .npz containers differ in zip time stamps even when the arrays are equal).
"""
import glob, os, shutil, subprocess, sys, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def same(a, b):
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    return np.array_equal(a, b, equal_nan=True) if a.dtype.kind in "fc" else np.array_equal(a, b)


def main():
    if not os.path.exists("/root/reference/functions.py"):
        print("check_golden: /root/reference is not here (this check runs in the authoring container only)")
        return 0
    keep = tempfile.mkdtemp(prefix="golden_keep_")
    files = sorted(glob.glob(os.path.join(GOLD, "*.npz")))
    for f in files:
        shutil.copy(f, keep)
    bad = 0
    try:
        for gen in ("make_golden.py", "make_golden_next.py"):
            subprocess.run([sys.executable, os.path.join(HERE, gen)], check=True, stdout=subprocess.DEVNULL)
        for f in files:
            new, old = np.load(f, allow_pickle=True), np.load(os.path.join(keep, os.path.basename(f)), allow_pickle=True)
            diff = sorted(set(new.files) ^ set(old.files)) + [k for k in set(new.files) & set(old.files) if not same(new[k], old[k])]
            print(f"{os.path.basename(f):20s} {len(old.files):4d} keys  {'ok' if not diff else 'DIFFERENT: ' + ', '.join(diff[:6])}")
            bad += len(diff)
    finally:
        for f in files:
            shutil.copy(os.path.join(keep, os.path.basename(f)), f)
        shutil.rmtree(keep, ignore_errors=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
