"""Loads a golden fixture (tests/golden/*.npz, written by tests/golden/make_golden.py) as a Problem."""
import os

import numpy as np

from helpers import Problem

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ("readme_toy", "single_view", "three_views")
N_SWEEPS = 10


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    V = int(z["n_views"])
    prob = Problem([z[f"x{v}"] for v in range(V)], [int(k) for k in z["k"]],
                   [z[f"f0_{v}"] for v in range(V)], [z[f"s0_{v}"] for v in range(V)],
                   [z[f"g0_{v}"] for v in range(V)], phi=z["phi"], xi=z["xi"], psi=z["psi"],
                   row_names=[[str(s) for s in z[f"rn{v}"]] for v in range(V)],
                   col_names=[[str(s) for s in z[f"cn{v}"]] for v in range(V)])
    return prob, z, V
