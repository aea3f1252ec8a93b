"""Generates tests/golden/large_*.npz: teacher-forced rho of the REAL reference (oracle/_ref, canonical build) on a few spatial
nodes of BASELINE.json's large configurations, deep into the history -- C3 (2d2v 32^2 x 128^2) at n = 100 / 400 / 800, C5-16 and
C5-32 (3d3v) at n = 25 -- sizes whose free run the CPU cannot afford.  The input history is oracle_py.exact_history: built from
correctly rounded +,-,*,/ only, hence bit-reproducible on any machine and NOT stored; the files hold the sampled nodes and the
reference's rho there.  Build container only (needs /root/reference); the vectors are committed.

    python tests/golden/make_large_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from bench import make_workload  # noqa: E402
from numericalflowiteration_b200 import n_nodes  # noqa: E402
from oracle.oracle_py import Reference, exact_history  # noqa: E402

PLAN = {"C3": [100, 400, 800], "C5-16": [25], "C5-32": [25]}
N_NODES = 8


def main():
    ref = Reference()
    for name, depths in PLAN.items():
        conf, f0, _, desc = make_workload(name, 1)
        hist = exact_history(conf, max(depths))
        nn = n_nodes(conf)
        nodes = np.array(sorted({(i * nn) // N_NODES + 5 * i for i in range(N_NODES)}))  # spread over the grid, varying x, y, z
        out = {"workload": desc, "depths": np.array(depths), "nodes": nodes, "amp": 1e-2,
               "history_sha256_first_level": np.frombuffer(__import__("hashlib").sha256(hist[:64].tobytes()).digest(), dtype=np.uint8)}
        for n in depths:
            out[f"rho_n{n}"] = np.array([ref.rho(conf, f0, n, hist, int(l), int(l) + 1)[int(l)] for l in nodes])
            print(name, n, out[f"rho_n{n}"][:3])
        np.savez_compressed(os.path.join(HERE, f"large_{name}.npz"), **out)


if __name__ == "__main__":
    main()
