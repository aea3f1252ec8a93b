"""CPU: the oracle (C restatement, oracle/nufi_oracle.c) against the committed golden vectors produced by the REAL
reference headers (tests/golden/make_golden.py), against the real reference live when oracle/_ref is built, and against
the known answers the reference's own test programs use (bin/test_poisson.cpp, bin/test_fields.cpp, step-0 closed form)."""
import math
import os

import numpy as np
import pytest

from cases import CASES, load_golden, rel_linf
from numericalflowiteration_b200 import Config1D, Config2D, Config3D, F0, stride_t

GOLDEN = ["1d-two-stream", "1d-landau", "2d-landau", "2d-two-stream", "3d-landau", "3d-bump"]
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("name", GOLDEN)
def test_teacher_forced_rho_bit_exact(name, oracle):
    """Same history in, same rho out: the restatement keeps the reference's expression order, so the canonical
    (-O2, no FMA contraction) build reproduces the reference's rho bit for bit."""
    conf, f0, g = load_golden(name)
    for n, want in zip(g["rho_steps"], g["rho"]):
        got = oracle.rho(conf, f0, int(n), g["coeffs"])
        assert np.array_equal(got, want), (name, int(n), np.max(np.abs(got - want)))


@pytest.mark.parametrize("name", GOLDEN)
def test_point_values_bit_exact(name, oracle):
    conf, f0, g = load_golden(name)
    d = conf.dim
    n_pt = int(g["n_pt"])
    st = stride_t(conf)
    level = g["coeffs"][(conf.Nt - 1) * st: conf.Nt * st]
    ders = [(0,) * d] + [tuple(int(i == j) for j in range(d)) for i in range(d)]
    for p, ft, fv, fld in zip(g["pts"], g["ftilda"], g["f"], g["field"]):
        assert oracle.ftilda(conf, f0, n_pt, g["coeffs"], p) == ft
        assert oracle.ftilda(conf, f0, n_pt - 1, g["coeffs"], p, full=True) == fv
        for der, want in zip(ders, fld):
            assert oracle.field(conf, level, p[:d], der) == want


@pytest.mark.parametrize("name", GOLDEN)
def test_free_run_matches_reference_history(name, oracle):
    """The whole CPU loop.  The oracle's interpolate is an exact solve where the reference iterates LSMR to eps, so
    histories agree to solver tolerance, not bitwise."""
    conf, f0, g = load_golden(name)
    coeffs, energy, _ = oracle.run(conf, f0, conf.Nt)
    assert rel_linf(coeffs, g["coeffs"]) <= 1e-11
    assert np.max(np.abs(energy - g["energy"]) / np.abs(g["energy"])) <= 1e-10


def test_basis_golden(oracle):
    g = np.load(os.path.join(HERE, "golden", "basis4.npz"))
    for der in (0, 1):
        for x, want in zip(g["xs"], g["basis"][der]):
            assert np.array_equal(oracle.basis(4, der, x), want)
    # closed forms
    t = 0.3
    n = oracle.basis(4, 0, t)
    assert np.allclose(n, [(1 - t) ** 3 / 6, (3 * t ** 3 - 6 * t ** 2 + 4) / 6, (-3 * t ** 3 + 3 * t ** 2 + 3 * t + 1) / 6, t ** 3 / 6], rtol=0, atol=2e-16)
    assert abs(n.sum() - 1) <= 4e-16
    for order in (2, 3, 5, 6):  # generic order: partition of unity, derivative sums to zero
        assert abs(oracle.basis(order, 0, t).sum() - 1) <= 1e-15
        assert abs(oracle.basis(order, 1, t).sum()) <= 1e-14


def test_defaults_match_reference_constructors():
    """Config{1,2,3}D() == the reference's default-constructed config_t<double> (config.hpp:55-70, 117-138, 196-219),
    and F0.default(dim) is the f0 line committed in config.hpp (evaluated by the as-is reference build)."""
    from oracle.oracle_py import Oracle

    g = np.load(os.path.join(HERE, "golden", "defaults.npz"))
    orc = Oracle()
    for C in (Config1D, Config2D, Config3D):
        c = C()
        got = np.array([getattr(c, n) for n, _ in c._fields_], dtype=np.float64)
        assert np.array_equal(got, g[f"conf{c.dim}d"]), c.dim
        f0 = F0.default(c.dim)
        for p, want in zip(g[f"f0pts{c.dim}d"], g[f"f0{c.dim}d"]):
            assert orc.f0(c, f0, *p) == want


def test_step0_closed_form(oracle):
    """rho^0 = 1 - int f0 dv = -alpha cos(kx) => electric energy alpha^2 L / (4 k^2) = 4 pi 1e-4 for the 1d default
    (the reference prints 1.2566370614359521e-03)."""
    conf = Config1D(Nt=1)
    for f0 in (F0(0, 0.01, 0.5),):
        rho = oracle.rho(conf, f0, 0, np.zeros(stride_t(conf)))
        x = conf.x_min + conf.dx * np.arange(conf.Nx)
        assert np.max(np.abs(rho + 0.01 * np.cos(0.5 * x))) <= 1e-14
        _, e = oracle.poisson(conf, rho)
        assert abs(e - 4e-4 * math.pi) <= 1e-15
        assert abs(e - 1.2566370614359521e-03) <= 1e-15


def test_poisson_known_answer(oracle):
    """bin/test_poisson.cpp:37-50: rho = cos 2z + sin 8y + cos 42x on [0,2pi]^3, 128 x 64 x 32
    => phi = cos 2z / 4 + sin 8y / 64 + cos 42x / 42^2."""
    conf = Config3D(Nx=128, Ny=64, Nz=32, x_max=2 * math.pi, y_max=2 * math.pi, z_max=2 * math.pi)
    z, y, x = np.meshgrid(conf.z_min + conf.dz * np.arange(conf.Nz), conf.y_min + conf.dy * np.arange(conf.Ny),
                          conf.x_min + conf.dx * np.arange(conf.Nx), indexing="ij")
    rho = np.cos(2 * z) + np.sin(8 * y) + np.cos(42 * x)
    want = np.cos(2 * z) / 4 + np.sin(8 * y) / 64 + np.cos(42 * x) / (42 * 42)
    phi, energy = oracle.poisson(conf, rho)
    assert np.sum(np.abs(phi - want.ravel())) / np.sum(np.abs(want)) <= 1e-13
    # energy = 1/2 int |grad phi|^2 = V/2 * (1/2)(1/4 + 1/64 + 1/42^2)
    v = (2 * math.pi) ** 3
    assert abs(energy - v / 4 * (1 / 4 + 1 / 64 + 1 / 42 ** 2)) <= 1e-12 * energy


def test_fields_known_answer(oracle):
    """bin/test_fields.cpp:94-126: interpolate sin 3x + sin 3y on a periodic grid, evaluate back at the nodes
    (reproduction to solver tolerance) and between them (O(h^4))."""
    conf = Config2D(Nx=64, Ny=48, x_min=0.0, x_max=2 * math.pi, y_min=-math.pi, y_max=math.pi)
    y, x = np.meshgrid(conf.y_min + conf.dy * np.arange(conf.Ny), conf.x_min + conf.dx * np.arange(conf.Nx), indexing="ij")
    vals = np.sin(3 * x) + np.sin(3 * y)
    level = oracle.interpolate(conf, vals)
    back = np.array([oracle.field(conf, level, (xx, yy)) for xx, yy in zip(x.ravel()[::7], y.ravel()[::7])])
    assert np.max(np.abs(back - vals.ravel()[::7])) <= 1e-13
    rng = np.random.default_rng(3)
    pts = rng.uniform([0, -math.pi], [2 * math.pi, math.pi], size=(200, 2))
    err = max(abs(oracle.field(conf, level, p) - (math.sin(3 * p[0]) + math.sin(3 * p[1]))) for p in pts)
    assert err <= 2e-4  # ~ (3h)^4 / 384-ish at h = 2 pi / 48
    derr = max(abs(oracle.field(conf, level, p, (1, 0)) - 3 * math.cos(3 * p[0])) for p in pts)
    assert derr <= 5e-3


def test_partial_sums_add_up(oracle):
    """Flat-q partials (GPU convention, no leading 1) over a ragged partition add up to the CPU-convention rho."""
    conf, f0, g = load_golden("2d-landau")
    n = conf.Nt // 2
    nq = conf.Nx * conf.Ny * conf.Nu * conf.Nv
    cuts = [0, 17, nq // 3 + 5, nq - 1, nq]
    total = np.zeros(conf.Nx * conf.Ny)
    for a, b in zip(cuts[:-1], cuts[1:]):
        oracle.rho_partial(conf, f0, n, g["coeffs"], a, b, rho=total)
    assert rel_linf(1 + total, oracle.rho(conf, f0, n, g["coeffs"])) <= 1e-13


# ---- live against the real reference (only where oracle/_ref is built, i.e. where /root/reference exists) ----
@pytest.mark.parametrize("name", list(CASES))
def test_oracle_vs_live_reference(name, oracle, reference):
    mk, f0 = CASES[name]
    conf = mk()
    cr, er, _ = reference.run(conf, f0, conf.Nt)
    for n in (0, 1, conf.Nt // 2, conf.Nt):
        assert np.array_equal(oracle.rho(conf, f0, n, cr), reference.rho(conf, f0, n, cr))
    co, eo, _ = oracle.run(conf, f0, conf.Nt)
    assert rel_linf(co, cr) <= 1e-11
    assert np.max(np.abs(eo - er) / np.abs(er)) <= 1e-10
    st = stride_t(conf)
    # exact collocation solve vs the reference's LSMR
    phi, _ = oracle.poisson(conf, oracle.rho(conf, f0, 3, cr))
    assert rel_linf(oracle.interpolate(conf, phi), reference.interpolate(conf, phi)) <= 1e-12
    assert st == cr.size // conf.Nt


def test_phase_flow_bit_exact(oracle):
    """orc_phase_flow_1d == the reference's eval_phase_flow (nufi/rho.hpp:98-131) on the committed fixture, n <= 1 quirk included."""
    import os

    conf, f0, g = load_golden("1d-two-stream")
    pf = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "phase_flow_1d.npz"))
    for n, want in zip(pf["steps"], pf["feet"]):
        got = oracle.phase_flow(conf, int(n), g["coeffs"], pf["pts"])
        assert np.array_equal(got, want), int(n)
    assert np.array_equal(pf["feet"][0][:, 1], pf["pts"][:, 1]) and np.array_equal(pf["feet"][1][:, 1], pf["pts"][:, 1])  # n <= 1: u untouched


# ------------------------------------------------------------------ spline orders other than 4 (tests/golden/orders.npz)
@pytest.mark.parametrize("order", [3, 5, 6, 8])
def test_generic_order_golden(order, oracle):
    """The reference is generic in the spline order (nufi/splines.hpp:39-110) and instantiates its kernels for 3..8
    (nufi/cuda_kernel.cu:191-203).  Vectors from the REAL reference templates at orders 3, 5, 6, 8
    (tests/golden/make_order_golden.py): basis values/derivatives and teacher-forced rho bit for bit, interpolate to LSMR's
    tolerance -- including odd orders on even grids, where the collocation system is singular and LSMR returns the
    minimum-norm least-squares solution."""
    from cases import ORDER_CASES

    g = np.load(os.path.join(HERE, "golden", "orders.npz"))
    for der in (0, 1, 2):
        for x, want in zip(g["xs"], g[f"basis_o{order}"][der]):
            assert np.array_equal(oracle.basis(order, der, x), want), (order, der, x)
    for name, (mk, f0, n_lev) in ORDER_CASES.items():
        conf = mk()
        coeffs = g[f"coeffs_{name}_o{order}"]
        for n, want in zip(g[f"steps_{name}_o{order}"], g[f"rho_{name}_o{order}"]):
            got = oracle.rho(conf, f0, int(n), coeffs, order=order)
            assert np.array_equal(got, want), (name, order, int(n), np.max(np.abs(got - want)))
        level = oracle.interpolate(conf, g[f"rho_{name}_o{order}"][-1], order=order)
        assert rel_linf(level, g[f"level_{name}_o{order}"]) <= 1e-11, (name, order)


@pytest.mark.parametrize("order", [3, 5, 7])
def test_generic_order_live_reference(order, oracle, reference):
    """Same check against the reference compiled here (orders beyond the committed vectors included)."""
    from cases import ORDER_CASES

    if not reference.has_orders:
        pytest.skip("oracle/_ref built before the order-generic entry points existed")
    for name, (mk, f0, n_lev) in ORDER_CASES.items():
        conf = mk()
        coeffs, _, _ = oracle.run(conf, f0, n_lev, order=order)
        assert np.isfinite(coeffs).all()
        want = reference.rho_order(conf, f0, order, n_lev, coeffs)
        assert np.array_equal(oracle.rho(conf, f0, n_lev, coeffs, order=order), want), (name, order)
        assert rel_linf(oracle.interpolate(conf, want, order=order), reference.interpolate_order(conf, order, want)) <= 1e-11


# ------------------------------------------------------------------ large configurations, deep histories (tests/golden/large_*.npz)
def _large(name):
    import sys

    sys.path.insert(0, os.path.dirname(HERE))
    from bench import make_workload
    from oracle.oracle_py import exact_history

    g = np.load(os.path.join(HERE, "golden", f"large_{name}.npz"))
    conf, f0, _, desc = make_workload(name, 1)
    assert str(g["workload"]) == desc
    return conf, f0, g, exact_history


@pytest.mark.parametrize("name,depth", [("C3", 100), ("C3", 400), ("C5-16", 25), ("C5-32", 25)])
def test_large_teacher_forced_rho_bit_exact(name, depth, oracle):
    """BASELINE.json's 2d2v / 3d3v configurations at full size, deep into the history: the reference's rho (real reference,
    tests/golden/make_large_golden.py) on nodes spread over the grid, input = the bit-reproducible exact_history.  (C3 at n = 800
    is left to the GPU suite: 1e8 point-steps of the canonical single-thread-per-node oracle.)"""
    conf, f0, g, exact_history = _large(name)
    hist = exact_history(conf, depth)
    import hashlib

    assert np.array_equal(np.frombuffer(hashlib.sha256(hist[:64].tobytes()).digest(), dtype=np.uint8), g["history_sha256_first_level"])
    for l, want in zip(g["nodes"], g[f"rho_n{depth}"]):
        assert oracle.rho(conf, f0, depth, hist, int(l), int(l) + 1)[int(l)] == want, (name, depth, int(l))
