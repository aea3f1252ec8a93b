"""Generates tests/golden/*.npz from the REAL reference (oracle/_ref/libnufi_ref.so = the reference's own headers
compiled in place from /root/reference by oracle/Makefile, canonical -O2 -ffp-contract=off build).  Run in the build
container only (/root/reference does not exist on the GPU box); the vectors are committed.

    python tests/golden/make_golden.py

Per case: the free-running reference history (reference eval_rho + reference LSMR interpolate; the Poisson stage is the
DHT restatement because FFTW is absent -- see oracle/ref_harness.cpp), the electric-energy trace, teacher-forced rho at
a few steps, point values of eval_ftilda/eval_f and of the field, and the cubic basis.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cases import CASES  # noqa: E402
from oracle.oracle_py import Reference  # noqa: E402

GOLDEN_CASES = ["1d-two-stream", "1d-landau", "2d-landau", "2d-two-stream", "3d-landau", "3d-bump"]


def conf_fields(conf):
    return {name: getattr(conf, name) for name, _ in conf._fields_}


def main():
    ref = Reference()
    rng = np.random.default_rng(20261017)
    for name in GOLDEN_CASES:
        mk, f0 = CASES[name]
        conf = mk()
        d = conf.dim
        coeffs, energy, _ = ref.run(conf, f0, conf.Nt)
        steps = np.array([0, 1, 2, conf.Nt // 2, conf.Nt - 1, conf.Nt])
        rho = np.stack([ref.rho(conf, f0, int(n), coeffs) for n in steps])
        # random phase-space points (some outside the box: periodic wrap), teacher-forced values
        lo = [conf.x_min - 3.0] * d + [-3.0] * d
        hi = [conf.x_max + 3.0] * d + [3.0] * d
        pts = rng.uniform(lo, hi, size=(24, 2 * d))
        n_pt = conf.Nt - 1
        ftilda = np.array([ref.ftilda(conf, f0, n_pt, coeffs, p) for p in pts])
        fval = np.array([ref.ftilda(conf, f0, n_pt - 1, coeffs, p, full=True) for p in pts])
        st = coeffs.size // conf.Nt
        level = coeffs[(conf.Nt - 1) * st: conf.Nt * st]
        ders = [tuple(int(i == j) for j in range(d)) for i in range(d)]
        field = np.array([[ref.field(conf, level, p[:d], der) for der in [(0,) * d] + ders] for p in pts])
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            conf_names=np.array(list(conf_fields(conf).keys())), conf_values=np.array(list(conf_fields(conf).values()), dtype=np.float64),
            f0_kind=np.int64(f0.kind), f0_p=np.array(list(f0.p)), coeffs=coeffs, energy=energy, rho_steps=steps, rho=rho,
            pts=pts, n_pt=np.int64(n_pt), ftilda=ftilda, f=fval, field=field)
        print(name, "levels", conf.Nt, "stride", st, "energy[0]", energy[0])
    # basis values and the reference's default-constructed configs / committed f0 (asis build)
    xs = np.linspace(0, 1, 17)[:-1]
    basis = np.stack([np.stack([ref.basis(der, x) for x in xs]) for der in (0, 1)])
    np.savez_compressed(os.path.join(HERE, "basis4.npz"), xs=xs, basis=basis)
    from numericalflowiteration_b200 import Config1D, Config2D, Config3D

    asis = Reference("_asis")
    out = {}
    for C in (Config1D, Config2D, Config3D):
        c = asis.default_conf(C())
        out[f"conf{c.dim}d"] = np.array(list(conf_fields(c).values()), dtype=np.float64)
        p = rng.uniform(-2, 2, size=(8, 2 * c.dim))
        out[f"f0pts{c.dim}d"] = p
        out[f"f0{c.dim}d"] = np.array([asis.f0(c, *q) for q in p])
    np.savez_compressed(os.path.join(HERE, "defaults.npz"), **out)
    print("defaults written")


if __name__ == "__main__":
    main()
