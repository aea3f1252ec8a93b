"""Shared small configurations for the parity tests (sizes the CPU oracle finishes in seconds)."""
import math

from numericalflowiteration_b200 import Config1D, Config2D, Config3D, F0


def conf1d(**kw):
    base = dict(Nx=64, Nu=48, Nt=40)
    base.update(kw)
    return Config1D(**base)


def conf2d(**kw):
    base = dict(Nx=16, Ny=12, Nu=10, Nv=6, Nt=24)
    base.update(kw)
    return Config2D(**base)


def conf3d(**kw):
    L = 10 * math.pi
    base = dict(Nx=8, Ny=6, Nz=5, Nu=4, Nv=3, Nw=5, Nt=16, u_min=-6, u_max=6, v_min=-6, v_max=6, w_min=-6, w_max=6,
                x_max=L, y_max=L, z_max=L)
    base.update(kw)
    return Config3D(**base)


CASES = {
    "1d-two-stream": (conf1d, F0(1, 0.01, 0.5)),
    "1d-landau": (conf1d, F0(0, 0.01, 0.5)),
    "2d-landau": (conf2d, F0(0, 0.05, 0.5)),
    "2d-two-stream": (conf2d, F0(1, 0.05, 0.5, 2.4)),
    "3d-landau": (conf3d, F0(0, 0.001, 0.2)),
    "3d-bump": (lambda **kw: conf3d(u_min=-9, u_max=0, v_min=-9, v_max=0, w_min=-9, w_max=0,
                                    x_max=20 * math.pi / 3, y_max=20 * math.pi / 3, z_max=20 * math.pi / 3, **kw),
                F0(2, 0.03, 0.3)),
    "3d-two-stream": (conf3d, F0(1, 0.001, 0.2, 2.4)),
}


# spline orders other than 4 (tests/golden/orders.npz): (config, f0, history levels); every N_d >= 8 = the largest order
ORDER_CASES = {
    "1d": (lambda: conf1d(Nx=32, Nu=24, Nt=12), F0(1, 0.01, 0.5), 12),
    "2d": (lambda: conf2d(Nx=10, Ny=9, Nu=6, Nv=5, Nt=8), F0(0, 0.05, 0.5), 8),
    "3d": (lambda: conf3d(Nx=8, Ny=9, Nz=8, Nu=3, Nv=4, Nw=3, Nt=6), F0(0, 0.001, 0.2), 6),
}


def rel_linf(a, b):
    import numpy as np

    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(np.asarray(b))))


def load_golden(name):
    """A committed reference fixture (tests/golden/<name>.npz, made by tests/golden/make_golden.py)."""
    import os

    import numpy as np

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    mk, f0 = CASES[name]
    conf = mk()
    for k, v in zip(g["conf_names"], g["conf_values"]):  # the fixture's config must be the case's config
        assert float(getattr(conf, str(k))) == float(v), (name, k)
    assert int(g["f0_kind"]) == f0.kind and list(g["f0_p"]) == list(f0.p)
    return conf, f0, g
