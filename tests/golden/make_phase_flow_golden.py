"""tests/golden/phase_flow_1d.npz from the REAL reference (nufi::dim1::eval_phase_flow<double,4>, nufi/rho.hpp:98-131, through
oracle/_ref): feet of the characteristics through 24 phase-space points on the golden 1d two-stream history, at n = 0, 1 (the
reference traces nothing for n <= 1), 2, Nt/2 and Nt-1.  Build container only.      python tests/golden/make_phase_flow_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cases import load_golden  # noqa: E402
from oracle.oracle_py import Reference  # noqa: E402


def main():
    conf, f0, g = load_golden("1d-two-stream")
    ref = Reference()
    rng = np.random.default_rng(20261018)
    pts = rng.uniform([conf.x_min - 3.0, -4.0], [conf.x_max + 3.0, 4.0], size=(24, 2))  # some outside the box: periodic wrap
    steps = np.array([0, 1, 2, conf.Nt // 2, conf.Nt - 1])
    feet = np.stack([ref.phase_flow(conf, int(n), g["coeffs"], pts) for n in steps])
    np.savez_compressed(os.path.join(HERE, "phase_flow_1d.npz"), pts=pts, steps=steps, feet=feet)
    print("feet", feet.shape, feet[-1, :2])


if __name__ == "__main__":
    main()
