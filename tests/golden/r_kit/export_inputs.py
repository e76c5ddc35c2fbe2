"""Step 1 of the R pinning kit: writes the inputs of every golden fixture (tests/golden/*.npz) in a form base R reads
without packages -- raw little-endian column-major doubles (<name>.f64), one name per line (<name>.txt) and a
manifest (manifest.txt: name nrow ncol) -- under tests/golden/r_kit/in/<case>/.
    python tests/golden/r_kit/export_inputs.py
Step 2 (on a machine with R and a checkout of eso28599/resnmtf):
    Rscript tests/golden/r_kit/make_golden.R <reference checkout> tests/golden/r_kit/in tests/golden/r_kit/out
Step 3: python tests/golden/r_kit/import_r_outputs.py   -> tests/golden/<case>_R.npz, read by tests/test_golden_r.py."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.dirname(HERE)
CASES = ("readme_toy", "single_view", "three_views")


def write_case(name, out_root):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    d = os.path.join(out_root, name)
    os.makedirs(d, exist_ok=True)
    V = int(z["n_views"])
    lines = []

    def put(key, arr):
        a = np.asarray(arr, dtype="<f8")
        a2 = a.reshape(a.shape[0], -1) if a.ndim else a.reshape(1, 1)
        np.asfortranarray(a2).T.tofile(os.path.join(d, f"{key}.f64"))  # .T of a Fortran array = column-major bytes
        lines.append(f"{key} {a2.shape[0]} {a2.shape[1]}")

    put("n_views", np.array([[float(V)]]))
    put("k", np.asarray(z["k"], dtype=float).reshape(-1, 1))
    for v in range(V):
        put(f"x{v}", z[f"x{v}"])
        put(f"f0_{v}", z[f"f0_{v}"])
        put(f"s0_{v}", z[f"s0_{v}"])
        put(f"g0_{v}", z[f"g0_{v}"])
        for kind in ("rn", "cn"):
            with open(os.path.join(d, f"{kind}{v}.txt"), "w") as fh:
                fh.write("\n".join(str(s) for s in z[f"{kind}{v}"]) + "\n")
    for m in ("phi", "xi", "psi"):
        put(m, z[m])
    with open(os.path.join(d, "manifest.txt"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return d


if __name__ == "__main__":
    root = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "in")
    for c in CASES:
        print("wrote", write_case(c, root))
