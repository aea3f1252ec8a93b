"""GPU: the 'next' rows of SURVEY section 8f -- checkpoint/restart from the coefficient history and batched sampling of the
field and of f for plots."""
import numpy as np
import pytest

from cases import CASES, rel_linf
from numericalflowiteration_b200 import CudaScheduler, RangeError, history_io, stride_t

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["1d-two-stream", "2d-landau", "3d-bump"])
def test_checkpoint_restart_is_bit_identical(name, tmp_path):
    """Run k steps, dump the history (binary and the reference's isolated-step text format), restore into a NEW handle,
    continue: levels and energies equal the uninterrupted run bit for bit."""
    mk, f0 = CASES[name]
    conf = mk()
    k = conf.Nt // 2
    with CudaScheduler(conf, f0) as s:
        for n in range(conf.Nt):
            s.step(n)
        want_hist = s.download_history(conf.Nt)
        want_energy = s.download_energy(0, conf.Nt)
    with CudaScheduler(conf, f0) as s:
        for n in range(k):
            s.step(n)
        history_io.write_binary(tmp_path / "ckpt.bin", conf, s.download_history(k), k)
        history_io.write_text_isolated(tmp_path / "ckpt.txt", conf, s.download_history(k), k)
    st = stride_t(conf)
    for loader in ("bin", "txt"):
        if loader == "bin":
            hdr, hist = history_io.read_binary(tmp_path / "ckpt.bin")
            assert hdr["n_levels"] == k
        else:
            hist = history_io.read_text(tmp_path / "ckpt.txt", k * st, header_lines=7)
        with CudaScheduler(conf, f0) as s:
            s.upload_history(hist, k)
            for n in range(k, conf.Nt):
                s.step(n)
            got_hist = s.download_history(conf.Nt)
            got_energy = s.download_energy(k, conf.Nt)
        assert np.array_equal(got_hist, want_hist), (name, loader)
        assert np.array_equal(got_energy, want_energy[k:]), (name, loader)


@pytest.mark.parametrize("name", ["1d-two-stream", "2d-landau", "3d-bump"])
def test_step_host_is_the_fused_step_with_a_host_history(name, oracle):
    """nufi_b200_step_host (the reference GPU drivers' loop body in one call, history kept by the caller on the host): a free run
    through it fills the host array with the same levels, bit for bit, as the device-resident fused steps, returns the same
    energies and the rho of every step (checked against the oracle's eval_rho on that history)."""
    mk, f0 = CASES[name]
    conf = mk()
    st = stride_t(conf)
    with CudaScheduler(conf, f0) as s:
        for n in range(conf.Nt):
            s.step(n)
        want_hist = s.download_history(conf.Nt)
        want_energy = s.download_energy(0, conf.Nt)
    host = np.zeros((conf.Nt + 1) * st)
    with CudaScheduler(conf, f0) as s:
        rho = np.zeros(s.n_nodes)
        energies = []
        for n in range(conf.Nt):
            energies.append(s.step_host(n, host, rho))
            if n in (1, conf.Nt - 1):
                assert rel_linf(rho, oracle.rho(conf, f0, n, host)) <= 1e-10
        with pytest.raises(RangeError):
            s.step_host(conf.Nt + 1, host)
    assert np.array_equal(host[: conf.Nt * st], want_hist[: conf.Nt * st]), name
    assert np.array_equal(np.array(energies), want_energy), name
    # a fresh handle continues from the host history alone: the levels it has not seen come in through upload_phi / step_host
    with CudaScheduler(conf, f0) as s:
        k = conf.Nt // 2
        s.upload_history(host, k)
        redo = host.copy()
        redo[k * st:] = 0
        for n in range(k, conf.Nt):
            s.step_host(n, redo)
    assert np.array_equal(redo[: conf.Nt * st], host[: conf.Nt * st]), name


@pytest.mark.parametrize("name", ["1d-two-stream", "2d-landau", "3d-landau"])
@pytest.mark.parametrize("xpp", [0, 1])
def test_sampling_matches_reference_point_functions(name, xpp, oracle, monkeypatch):
    """eval_f / eval_ftilda at random phase-space points and eval<...> of phi and its first derivatives at random positions
    (some outside the box: periodic wrap) against the reference functions (oracle)."""
    mk, f0 = CASES[name]
    conf = mk()
    if conf.dim == 1 and xpp:
        pytest.skip("1d has a single level format")
    monkeypatch.setenv("NUFI_B200_XPP", str(xpp))
    d = conf.dim
    coeffs, _, _ = oracle.run(conf, f0, conf.Nt)
    rng = np.random.default_rng(11)
    lo = [conf.x_min - 3.0] * d + [-3.0] * d
    hi = [conf.x_max + 3.0] * d + [3.0] * d
    pts = rng.uniform(lo, hi, size=(64, 2 * d))
    n = conf.Nt - 1
    st = stride_t(conf)
    level = coeffs[n * st:(n + 1) * st]
    with CudaScheduler(conf, f0) as s:
        s.upload_history(coeffs, conf.Nt)
        f_full = s.eval_f(n, pts, full=True)
        f_tilda = s.eval_f(n, pts, full=False)
        f_zero = s.eval_f(0, pts, full=True)
        phi = s.eval_field(n, pts[:, :d])
        grads = [s.eval_field(n, pts[:, :d], axis) for axis in range(d)]
    want_full = np.array([oracle.ftilda(conf, f0, n, coeffs, p, full=True) for p in pts])
    want_tilda = np.array([oracle.ftilda(conf, f0, n, coeffs, p) for p in pts])
    want_zero = np.array([oracle.f0(conf, f0, *p) for p in pts])
    assert rel_linf(f_full, want_full) <= 1e-11
    assert rel_linf(f_tilda, want_tilda) <= 1e-11
    assert rel_linf(f_zero, want_zero) <= 1e-13
    want_phi = np.array([oracle.field(conf, level, p[:d]) for p in pts])
    assert rel_linf(phi, want_phi) <= 1e-12
    for axis, g in enumerate(grads):
        der = tuple(int(axis == j) for j in range(d))
        want = np.array([oracle.field(conf, level, p[:d], der) for p in pts])
        assert rel_linf(g, want) <= 1e-12


def test_eval_phase_flow_matches_reference():
    """nufi_b200_eval_phase_flow against the committed outputs of the reference's eval_phase_flow (nufi/rho.hpp:98-131;
    tests/golden/phase_flow_1d.npz), including its n <= 1 behaviour (nothing traced, x only reduced into the box)."""
    import os

    from cases import load_golden

    conf, f0, g = load_golden("1d-two-stream")
    pf = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "phase_flow_1d.npz"))
    with CudaScheduler(conf, f0, device=0) as s:
        s.upload_history(g["coeffs"], conf.Nt)
        for n, want in zip(pf["steps"], pf["feet"]):
            got = s.eval_phase_flow(int(n), pf["pts"])
            dx = np.abs(got[:, 0] - want[:, 0])
            dx = np.minimum(dx, conf.Lx - dx)  # a foot within rounding of the box edge may come back on the other side
            assert np.max(dx) <= 1e-10 * conf.Lx, (int(n), float(np.max(dx)))
            assert np.max(np.abs(got[:, 1] - want[:, 1])) <= 1e-10 * np.max(np.abs(want[:, 1])), int(n)
            assert np.all(got[:, 0] >= 0) and np.all(got[:, 0] <= conf.Lx)
        with pytest.raises(RangeError):
            s.eval_phase_flow(conf.Nt + 1, pf["pts"])
