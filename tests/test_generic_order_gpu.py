"""GPU: spline orders other than 4 (the reference instantiates its kernels for orders 3..8, nufi/cuda_kernel.cu:191-203,
373-385, 575-587; nufi/splines.hpp:39-110 is generic) and the 1d scheduler's separate metrics grid
(nufi/cuda_scheduler.hpp:65-85, nufi/cuda_kernel.cu:55-70), through the C ABI, against the oracle -- which
tests/test_oracle_golden.py pins bit for bit against the real reference templates at these orders."""
import numpy as np
import pytest

from cases import ORDER_CASES, conf1d, rel_linf
from numericalflowiteration_b200 import Config1D, CudaScheduler, F0, n_quad, stride_t

pytestmark = pytest.mark.gpu
RHO_TOL, ENERGY_TOL, COEFF_TOL = 1e-10, 1e-8, 1e-11
ORDERS = [3, 5, 6, 7, 8]


@pytest.mark.parametrize("order", ORDERS)
@pytest.mark.parametrize("name", list(ORDER_CASES))
def test_rho_teacher_forced_generic_order(name, order, oracle):
    mk, f0, n_lev = ORDER_CASES[name]
    conf = mk()
    coeffs, _, _ = oracle.run(conf, f0, n_lev, order=order)
    with CudaScheduler(conf, f0, order=order) as s:
        assert s.stride_t == stride_t(conf, order)
        s.upload_history(coeffs, n_lev)
        for n in (0, 1, 2, n_lev - 1, n_lev):
            got = s.eval_rho(n)
            want = oracle.rho(conf, f0, n, coeffs, order=order)
            err = rel_linf(got, want)
            assert err <= RHO_TOL, (name, order, n, err, s.last_variant)
        assert f"order{order}" in s.last_variant
        # a level goes up and comes back unchanged (reference layout with the order-1 halo)
        st = stride_t(conf, order)
        assert np.array_equal(s.download_phi(2), coeffs[2 * st:3 * st])


@pytest.mark.parametrize("order", [3, 5, 6])
@pytest.mark.parametrize("tail", [1, 2], ids=["cufft", "fused-1cta"])
@pytest.mark.parametrize("name", list(ORDER_CASES))
def test_free_run_generic_order(name, order, tail, oracle):
    """The fused step at order != 4: the tail's collocation symbol sum_i N_i(0) w^i (pseudo-inverse at the Nyquist mode for
    odd orders on even grids, what the reference's LSMR converges to) and the order-1 halo."""
    mk, f0, n_lev = ORDER_CASES[name]
    conf = mk()
    coeffs, energy, _ = oracle.run(conf, f0, n_lev, order=order)
    with CudaScheduler(conf, f0, order=order) as s:
        s.set_tail_variant(tail)
        for n in range(n_lev):
            s.step(n)
        got = s.download_energy(0, n_lev)
        last = s.download_phi(n_lev - 1)
    assert np.max(np.abs(got - energy) / np.abs(energy)) <= ENERGY_TOL, (name, order)
    st = stride_t(conf, order)
    assert rel_linf(last, coeffs[(n_lev - 1) * st:n_lev * st]) <= 1e-8


@pytest.mark.parametrize("order", [3, 6])
@pytest.mark.parametrize("name", list(ORDER_CASES))
def test_metrics_and_sampling_generic_order(name, order, oracle):
    mk, f0, n_lev = ORDER_CASES[name]
    conf = mk()
    d = conf.dim
    coeffs, _, _ = oracle.run(conf, f0, n_lev, order=order)
    n = n_lev - 1
    nq = n_quad(conf)
    rng = np.random.default_rng(5)
    with CudaScheduler(conf, f0, order=order) as s:
        s.upload_history(coeffs, n_lev)
        s.compute_metrics(n, 0, nq)
        m = np.zeros(4)
        s.download_metrics(m)
        want = oracle.metrics(conf, f0, n, coeffs, 0, nq, order=order)
        assert np.max(np.abs(m - want) / np.abs(want)) <= 1e-11, (name, order, m, want)
        lo = [conf.x_min - 2.0] * d + [-3.0] * d
        hi = [conf.x_max + 2.0] * d + [3.0] * d
        pts = rng.uniform(lo, hi, size=(40, 2 * d))
        got = s.eval_f(n, pts, full=True)
        ref = np.array([oracle.ftilda(conf, f0, n, coeffs, p, order=order, full=True) for p in pts])
        assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))
        st = stride_t(conf, order)
        level = coeffs[n * st:(n + 1) * st]
        for axis in range(-1, d):
            der = tuple(int(a == axis) for a in range(d))
            vals = s.eval_field(n, pts[:, :d], axis)
            refv = np.array([oracle.field(conf, level, p[:d], der, order=order) for p in pts])
            assert np.max(np.abs(vals - refv)) <= 1e-12 * max(np.max(np.abs(refv)), 1e-300), (name, order, axis)


@pytest.mark.parametrize("order", [4, 5])
def test_metrics_on_a_separate_grid_1d(order, oracle):
    """cuda_scheduler(conf, conf_metrics): the metrics are integrated over the (x,u) nodes of conf_metrics -- nodes
    x_min + ix*dx, u_min + iu*du + du/2, weight du*dx, all from conf_metrics -- with eval_f on the field grid of conf
    (nufi/cuda_kernel.cu:55-79).  The metrics grid is NOT aligned with the field grid here (37 nodes against 64 cells)."""
    conf, f0 = conf1d(), F0(1, 0.01, 0.5)
    n_lev = 20
    coeffs, _, _ = oracle.run(conf, f0, n_lev, order=order)
    n = n_lev - 1
    cm = Config1D(Nx=37, Nu=29, Nt=conf.Nt, u_min=-7.0, u_max=8.0)
    want = np.zeros(4)
    for ix in range(cm.Nx):  # the reference kernel's arithmetic, one point at a time, flat-q order
        for iu in range(cm.Nu):
            x = cm.x_min + ix * cm.dx
            u = cm.u_min + iu * cm.du + cm.du / 2
            f = oracle.ftilda(conf, f0, n, coeffs, (x, u), order=order, full=True)
            w = cm.du * cm.dx
            want += [w * f, w * f * f, w * (u * u * f / 2), (-w * f * np.log(f)) if f > 0 else 0.0]
    with CudaScheduler(conf, f0, order=order) as s:
        s.upload_history(coeffs, n_lev)
        s.set_metrics_grid(cm)
        with pytest.raises(Exception):
            s.compute_metrics(n, 0, cm.Nx * cm.Nu + 1)  # the range is checked against the metrics grid
        got = np.zeros(4)
        for a, b in ((0, 400), (400, cm.Nx * cm.Nu)):  # ragged split: download accumulates
            s.compute_metrics(n, a, b)
            s.download_metrics(got)
        assert np.max(np.abs(got - want) / np.abs(want)) <= 1e-11, (got, want)
        s.set_metrics_grid(None)  # back to the scheduler's own grid
        s.compute_metrics(n, 0, n_quad(conf))
        own = np.zeros(4)
        s.download_metrics(own)
        want_own = oracle.metrics(conf, f0, n, coeffs, 0, n_quad(conf), order=order)
        assert np.max(np.abs(own - want_own) / np.abs(want_own)) <= 1e-11
