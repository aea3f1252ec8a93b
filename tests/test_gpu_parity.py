"""GPU parity tests proper: libnufi_b200.so (through its C ABI) against the CPU oracle on the same inputs.

Tolerances are the north star's: rho relative L-infinity <= 1e-10 per step on the SAME history
(teacher-forced), electric-energy trace relative error <= 1e-8 over a free run."""
import numpy as np
import pytest

from cases import CASES, conf1d, conf2d, conf3d, rel_linf
from numericalflowiteration_b200 import CudaScheduler, F0, RangeError, n_quad, stride_t

pytestmark = pytest.mark.gpu

RHO_TOL = 1e-10
ENERGY_TOL = 1e-8
COEFF_TOL = 1e-11


@pytest.fixture(scope="module")
def histories(oracle):
    """Free-running oracle histories (the reference CPU loop) for every case."""
    out = {}
    for name, (mk, f0) in CASES.items():
        conf = mk()
        n_lev = conf.Nt
        coeffs, energy, _ = oracle.run(conf, f0, n_lev)
        out[name] = (conf, f0, coeffs, energy)
    return out


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("xpp", [0, 1], ids=["bspline-levels", "xpp-levels"])
def test_rho_teacher_forced(name, variant, xpp, histories, oracle, monkeypatch):
    """variant: 1 global-memory kernel, 2 shared-memory staged kernel; xpp: level format of 2d/3d histories (B-spline
    window vs per-row cubics in x)."""
    conf, f0, coeffs, _ = histories[name]
    if conf.dim == 1 and xpp:
        pytest.skip("1d has a single level format")
    monkeypatch.setenv("NUFI_B200_XPP", str(xpp))
    with CudaScheduler(conf, f0) as s:
        s.set_variant(variant)
        s.upload_history(coeffs, conf.Nt)
        for n in (0, 1, 2, 5, conf.Nt - 1, conf.Nt):
            got = s.eval_rho(n)
            want = oracle.rho(conf, f0, n, coeffs)
            err = rel_linf(got, want)
            assert err <= RHO_TOL, (name, variant, n, err, s.last_variant)


@pytest.mark.parametrize("name", list(CASES))
def test_cluster_pairs_share_the_history_stream(name, histories, oracle, monkeypatch):
    """NUFI_B200_CLUSTER=2: the staged kernel runs as clusters of two CTAs, each fetching half of every history chunk and
    multicasting it to both (opt-in).  Same rho as the unpaired kernel bit for bit (the arithmetic and the
    order of every sum are unchanged), ragged cases included: a CTA whose partner has more rounds, an idle partner, and the
    fused step reading the self-validating slots of a paired launch."""
    conf, f0, coeffs, _ = histories[name]
    with CudaScheduler(conf, f0) as s:
        s.set_variant(2)
        s.upload_history(coeffs, conf.Nt)
        n = conf.Nt - 1
        ref = s.eval_rho(n)
        ref_variant = s.last_variant
        monkeypatch.setenv("NUFI_B200_CLUSTER", "2")
        got = s.eval_rho(n)
        if "-mc2" not in s.last_variant:
            pytest.skip(f"cluster pairs unavailable on this device ({s.last_variant})")
        assert "-mc2" not in ref_variant
        assert np.array_equal(got, ref), (name, s.last_variant)
        assert rel_linf(got, oracle.rho(conf, f0, n, coeffs)) <= RHO_TOL
        nq = n_quad(conf)
        s.compute_rho(n, nq // 3, nq // 3 + max(1, nq // 50))  # a few tiles only: most pairs idle, some half idle
        part = np.zeros(s.n_nodes)
        s.download_rho(part)
        want = oracle.rho_partial(conf, f0, n, coeffs, nq // 3, nq // 3 + max(1, nq // 50))
        assert np.max(np.abs(part - want)) <= 1e-12 * max(np.max(np.abs(want)), 1e-300) + 1e-15
        s.step(n)  # fused step on the paired kernel
        lvl_pair = s.download_phi(n)
        monkeypatch.setenv("NUFI_B200_CLUSTER", "1")
        s.step(n)
        assert np.array_equal(lvl_pair, s.download_phi(n))


@pytest.mark.parametrize("name", ["1d-two-stream", "2d-landau", "3d-landau"])
def test_partial_ranges_accumulate(name, histories, oracle):
    """compute_rho/download_rho keep the reference GPU convention: partial = -dV*sum f, download accumulates;
    arbitrary [q_begin,q_end) cuts (ragged ends inside a node's velocity range) sum to the whole."""
    conf, f0, coeffs, _ = histories[name]
    nq = n_quad(conf)
    rng = np.random.default_rng(7)
    cuts = sorted(set([0, nq] + [int(c) for c in rng.integers(1, nq - 1, size=5)]))
    n = conf.Nt // 2
    with CudaScheduler(conf, f0) as s:
        s.upload_history(coeffs, conf.Nt)
        total = np.zeros(s.n_nodes)
        for a, b in zip(cuts[:-1], cuts[1:]):
            s.compute_rho(n, a, b)
            part = np.zeros(s.n_nodes)
            s.download_rho(part)
            want = oracle.rho_partial(conf, f0, n, coeffs, a, b)
            scale = np.max(np.abs(want))
            assert np.max(np.abs(part - want)) <= 1e-12 * max(scale, 1e-300) + 1e-15, (name, a, b)
            total += part
        s.compute_rho(n, 5, 5)  # empty range -> zero partial
        z = np.zeros(s.n_nodes)
        s.download_rho(z)
        assert not z.any()
        want = oracle.rho(conf, f0, n, coeffs)
        assert rel_linf(1.0 + total, want) <= RHO_TOL


@pytest.mark.parametrize("tail", [1, 2], ids=["cufft", "fused-1cta"])
@pytest.mark.parametrize("name", ["1d-landau", "2d-landau", "3d-landau", "3d-bump"])
def test_field_tail(name, tail, histories, oracle):
    """solve + interpolate on the device vs poisson.cpp / fields.hpp restated (and LSMR via oracle/_ref in
    test_oracle_vs_ref): coefficients and electric energy."""
    conf, f0, coeffs, _ = histories[name]
    n = 3
    rho = oracle.rho(conf, f0, n, coeffs)
    phi, e_want = oracle.poisson(conf, rho)
    level_want = oracle.interpolate(conf, phi)
    with CudaScheduler(conf, f0) as s:
        s.set_tail_variant(tail)
        e_got = s.solve_interpolate(n, rho=rho)
        level_got = s.download_phi(n)
        assert s.last_tail_variant == ("cufft" if tail == 1 else "fused-1cta")
    assert abs(e_got - e_want) <= 1e-12 * abs(e_want)
    assert rel_linf(level_got, level_want) <= COEFF_TOL


@pytest.mark.parametrize("tail", [0, 1], ids=["tail-auto", "tail-cufft"])
@pytest.mark.parametrize("name", ["1d-two-stream", "1d-landau", "2d-landau", "3d-landau", "3d-bump"])
def test_free_run_energy_trace(name, tail, histories, monkeypatch):
    """The fused step() loop, no host round trip, against the reference CPU loop."""
    conf, f0, coeffs, energy = histories[name]
    monkeypatch.setenv("NUFI_B200_XPP", str(tail))  # tail-auto with B-spline levels, tail-cufft with xpp levels
    with CudaScheduler(conf, f0) as s:
        s.set_tail_variant(tail)
        for n in range(conf.Nt):
            s.step(n)
        got = s.download_energy(0, conf.Nt)
        last = s.download_phi(conf.Nt - 1)
    rel = np.max(np.abs(got - energy) / np.abs(energy))
    assert rel <= ENERGY_TOL, (name, rel)
    st = stride_t(conf)
    assert rel_linf(last, coeffs[(conf.Nt - 1) * st: conf.Nt * st]) <= 1e-8


@pytest.mark.parametrize("xpp", [0, 1])
def test_upload_download_roundtrip_and_errors(xpp, histories, monkeypatch):
    conf, f0, coeffs, _ = histories["2d-landau"]
    st = stride_t(conf)
    monkeypatch.setenv("NUFI_B200_XPP", str(xpp))
    with CudaScheduler(conf, f0) as s:
        s.upload_phi(2, coeffs)
        assert np.array_equal(s.download_phi(2), coeffs[2 * st:3 * st])
        with pytest.raises(RangeError):
            s.compute_rho(conf.Nt + 1, 0, 10)  # "Time-step out of range." (cuda_kernel.cu:115-116)
        with pytest.raises(RangeError):
            s.compute_rho(5, 0, 10)  # levels 0,1,3,4 never uploaded
        with pytest.raises(RangeError):
            s.compute_rho(0, 0, n_quad(conf) + 1)
    with pytest.raises(ValueError):
        CudaScheduler(conf, f0, order=9)  # orders 3..8 exist (tests/test_generic_order_gpu.py)
    with pytest.raises(ValueError):
        CudaScheduler(conf, F0(7))


@pytest.mark.parametrize("name", ["1d-two-stream", "2d-landau", "3d-landau"])
def test_metrics(name, histories, oracle):
    conf, f0, coeffs, _ = histories[name]
    nq = n_quad(conf)
    n = conf.Nt - 1
    with CudaScheduler(conf, f0) as s:
        s.upload_history(coeffs, conf.Nt)
        got = np.zeros(4)
        for a, b in ((0, nq // 3), (nq // 3, nq)):
            s.compute_metrics(n, a, b)
            s.download_metrics(got)
        s.compute_metrics(0, 0, nq)
        got0 = np.zeros(4)
        s.download_metrics(got0)
    want = oracle.metrics(conf, f0, n, coeffs, 0, nq)
    want0 = oracle.metrics(conf, f0, 0, coeffs, 0, nq)
    assert np.max(np.abs(got - want) / np.abs(want)) <= 1e-11
    assert np.max(np.abs(got0 - want0) / np.abs(want0)) <= 1e-11


def test_run_to_run_deterministic(histories):
    conf, f0, coeffs, _ = histories["2d-landau"]
    with CudaScheduler(conf, f0) as s:
        s.upload_history(coeffs, conf.Nt)
        a = s.eval_rho(conf.Nt)
        b = s.eval_rho(conf.Nt)
    assert np.array_equal(a, b)


def test_multi_round_ragged_tiles(oracle):
    """More CTA-rounds than SMs (several rounds per CTA, tile changes inside a CTA), tiles that straddle grid rows
    (Nx = 40), a last tile that is only partly filled, non-power-of-two periodic wrap."""
    conf = conf2d(Nx=40, Ny=25, Nu=20, Nv=12, Nt=8)
    f0 = F0(0, 0.05, 0.5)
    coeffs, _, _ = oracle.run(conf, f0, conf.Nt)
    with CudaScheduler(conf, f0) as s:
        s.upload_history(coeffs, conf.Nt)
        for variant in (1, 2):
            s.set_variant(variant)
            got = s.eval_rho(conf.Nt)
            assert rel_linf(got, oracle.rho(conf, f0, conf.Nt, coeffs)) <= RHO_TOL, s.last_variant


@pytest.mark.parametrize("nx", [12, 16])
def test_multi_period_jump_in_one_step(nx, oracle):
    """A point that crosses more than one whole period in a single drift (dt*|u|/dx > Nx).  Power-of-two grids wrap by
    masking; other grids clamp, flag and redo the point on the robust path -- both must agree with the reference's
    floor()-based wrap."""
    conf = conf1d(Nx=nx, Nu=64, Nt=6, dt=4.0, u_min=-6.0, u_max=6.0)  # |u| ~ 4 carries weight and jumps > Lx
    f0 = F0(0, 0.01, 0.5)
    coeffs, _, _ = oracle.run(conf, f0, conf.Nt)
    assert conf.dt * 4.0 / conf.dx > conf.Nx
    with CudaScheduler(conf, f0) as s:
        s.upload_history(coeffs, conf.Nt)
        for n in (1, 3, conf.Nt):
            got = s.eval_rho(n)
            assert rel_linf(got, oracle.rho(conf, f0, n, coeffs)) <= RHO_TOL, (nx, n)


@pytest.mark.parametrize("shape", ["1d-256", "1d-512", "2d-32x32", "2d-64x32", "3d-16^3", "3d-8x16x4"])
def test_fused_step_grid_sizes(shape, oracle):
    """The fused step (slot reduction + field tail in one CTA, or the cuFFT tail above 4096 nodes / 256 per dimension) on
    the grid sizes of the benchmark configurations; few velocities so the CPU oracle stays cheap."""
    import math

    L = 10 * math.pi
    if shape.startswith("1d"):
        conf, f0 = conf1d(Nx=int(shape[3:]), Nu=32, Nt=6), F0(1, 0.01, 0.5)
    elif shape.startswith("2d"):
        nx, ny = (int(v) for v in shape[3:].split("x"))
        conf, f0 = conf2d(Nx=nx, Ny=ny, Nu=6, Nv=4, Nt=5), F0(0, 0.05, 0.5)
    else:
        dims = (16, 16, 16) if shape == "3d-16^3" else (8, 16, 4)
        conf, f0 = conf3d(Nx=dims[0], Ny=dims[1], Nz=dims[2], Nu=3, Nv=2, Nw=2, Nt=4), F0(0, 0.001, 0.2)
    coeffs, energy, _ = oracle.run(conf, f0, conf.Nt)
    with CudaScheduler(conf, f0) as s:
        for n in range(conf.Nt):
            s.step(n)
        got = s.download_energy(0, conf.Nt)
        last = s.download_phi(conf.Nt - 1)
        rho = s.eval_rho(conf.Nt)
    assert np.max(np.abs(got - energy) / np.abs(energy)) <= ENERGY_TOL, (shape, s.last_tail_variant)
    st = stride_t(conf)
    assert rel_linf(last, coeffs[(conf.Nt - 1) * st: conf.Nt * st]) <= 1e-8
    assert rel_linf(rho, oracle.rho(conf, f0, conf.Nt, coeffs)) <= 1e-9  # own (free-running) history vs oracle history


@pytest.mark.parametrize("tn", [16, 8, 4, 1])
@pytest.mark.parametrize("name", ["1d-two-stream", "2d-landau", "3d-landau", "3d-bump"])
def test_tile_nodes_lane_layouts(name, tn, histories, oracle):
    """Every lane layout (nodes per tile 32 ... 1: 32/TN lanes per node with neighbouring velocities, combined by a shuffle
    tree before the slot is written) gives the reference's rho: whole range, a ragged q-range, the fused step and the
    peer-exchange path with a world of one."""
    conf, f0, coeffs, energy = histories[name]
    n = conf.Nt - 1
    want = oracle.rho(conf, f0, n, coeffs)
    with CudaScheduler(conf, f0, device=0) as s:
        s.set_tile_nodes(tn)
        s.upload_history(coeffs, n)
        got = s.eval_rho(n)
        assert f"/tn{tn}" in s.last_variant
        assert rel_linf(got, want) <= RHO_TOL
        nq = s.n_quad
        q0, q1 = nq // 3 + 5, (2 * nq) // 3 + 1
        part = np.zeros(s.n_nodes)
        s.compute_rho(n, q0, q1)
        s.download_rho(part)
        ref_part = oracle.rho_partial(conf, f0, n, coeffs, q0, q1)
        assert np.max(np.abs(part - ref_part)) <= RHO_TOL * np.max(np.abs(want))
        s.step(n)
        e = s.download_energy(n, n + 1)[0]
        assert abs(e - energy[n]) <= 1e-8 * abs(energy[n])
        s.peer_attach(0, 1, s.peer_export(1))
        s.peer_step(n)
        e2 = s.download_energy(n, n + 1)[0]
        assert abs(e2 - e) <= 1e-12 * abs(e)
        assert not s.peer_timed_out()


@pytest.mark.parametrize("name", ["C3", "C5-16", "C5-32"])
def test_large_configurations_deep_history_against_reference_golden(name):
    """Teacher-forced rho at BASELINE.json's full 2d2v / 3d3v sizes, deep into the history (C3: n = 100, 400, 800; C5-16, C5-32:
    n = 25), against the REAL reference's values on nodes spread over the grid (tests/golden/large_*.npz).  The input history is
    regenerated bit for bit from exact arithmetic (oracle_py.exact_history), so nothing large is stored."""
    import os
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from bench import make_workload
    from oracle.oracle_py import exact_history

    g = np.load(os.path.join(root, "tests", "golden", f"large_{name}.npz"))
    conf, f0, _, desc = make_workload(name, 1)
    assert str(g["workload"]) == desc
    depths = [int(d) for d in g["depths"]]
    hist = exact_history(conf, max(depths))
    nodes = g["nodes"].astype(int)
    with CudaScheduler(conf, f0, device=0) as s:
        s.upload_history(hist, max(depths))
        for n in depths:
            got = s.eval_rho(n)[nodes]
            want = g[f"rho_n{n}"]
            err = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
            print(f"{name} n={n} [{s.last_variant}]: rho rel-Linf over {len(nodes)} sampled nodes {err:.2e}")
            assert err <= RHO_TOL, (name, n, err)
