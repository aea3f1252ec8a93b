"""Generates tests/golden/orders.npz: the REAL reference's templates (oracle/_ref/libnufi_ref.so, see oracle/ref_harness.cpp
ref_*_order_*) at spline orders 3, 5, 6, 8 -- basis values and derivatives, teacher-forced rho, interpolate -- on small
1d/2d/3d cases.  Every reference driver runs order 4 (tests/golden/<case>.npz); these vectors pin the order-generic code of the
oracle and, through it, the generic-order kernels of libnufi_b200.  Build container only; the vectors are committed.

    python tests/golden/make_order_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cases import ORDER_CASES  # noqa: E402
from oracle.oracle_py import Oracle, Reference  # noqa: E402

ORDERS = [3, 5, 6, 8]


def main():
    ref, orc = Reference(), Oracle()
    out = {"orders": np.array(ORDERS)}
    xs = np.array([0.0, 0.125, 0.3, 0.5, 0.77, 0.999])
    for o in ORDERS:
        out[f"basis_o{o}"] = np.stack([np.stack([ref.basis_order(o, der, x) for x in xs]) for der in (0, 1, 2)])
        for name, (mk, f0, n_lev) in ORDER_CASES.items():
            conf = mk()
            # input history: the oracle's own free run at this order (any smooth history would do; it is stored)
            coeffs, _, _ = orc.run(conf, f0, n_lev, order=o)
            steps = [1, n_lev // 2, n_lev]
            out[f"coeffs_{name}_o{o}"] = coeffs
            out[f"steps_{name}_o{o}"] = np.array(steps)
            out[f"rho_{name}_o{o}"] = np.stack([ref.rho_order(conf, f0, o, n, coeffs) for n in steps])
            vals = ref.rho_order(conf, f0, o, n_lev, coeffs)
            out[f"level_{name}_o{o}"] = ref.interpolate_order(conf, o, vals)  # LSMR collocation solve of the reference
    out["xs"] = xs
    np.savez_compressed(os.path.join(HERE, "orders.npz"), **out)
    print("wrote orders.npz", os.path.getsize(os.path.join(HERE, "orders.npz")), "bytes")


if __name__ == "__main__":
    main()
