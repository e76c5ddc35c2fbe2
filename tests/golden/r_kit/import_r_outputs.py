"""Step 3 of the R pinning kit: packs what make_golden.R wrote (tests/golden/r_kit/out/<case>/) into
tests/golden/<case>_R.npz.  tests/test_golden_r.py then holds the oracle (and the KDE / JSD restatement) to those numbers."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.dirname(HERE)


def read_case(d):
    out = {}
    with open(os.path.join(d, "manifest.txt")) as fh:
        for line in fh:
            name, nr, nc = line.split()
            a = np.fromfile(os.path.join(d, f"{name}.f64"), dtype="<f8")
            out[name] = a.reshape((int(nc), int(nr))).T.copy()  # column-major on disk
    return out


if __name__ == "__main__":
    root = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "out")
    for case in sorted(os.listdir(root)):
        arrs = read_case(os.path.join(root, case))
        path = os.path.join(GOLDEN, f"{case}_R.npz")
        np.savez_compressed(path, **arrs)
        print("wrote", path, sorted(arrs))
