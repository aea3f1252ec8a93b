"""GPU: the C++ host side.  The drop-in drivers under bin/ (the reference's bin/test_nufi_{cpu,gpu}_{1,2,3}d loops written
against include/nufi/*.hpp over the C ABI) must reproduce the reference CPU loop's electric-energy trace (<= 1e-8), through
the reference scheduler's five-method round trip, through the fused device-resident step, and through the CPU-shaped
per-node eval_rho loop.  Also: the stand-alone poisson / interpolate entry points and the one-process multi-GPU group."""
import math
import os
import subprocess

import numpy as np
import pytest

from cases import CASES, rel_linf
from numericalflowiteration_b200 import Config1D, Config2D, Config3D, CudaGroup, CudaScheduler, F0, _lib, device_count, stride_t

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin", "build")
ENERGY_TOL = 1e-8


def landau_conf(dim, steps):
    """What `--landau --steps N` selects in bin/nufi_drivers.hpp."""
    if dim == 1:
        return Config1D(Nt=steps), F0(0, 0.01, 0.5)
    if dim == 2:
        return Config2D(Nt=steps), F0(0, 0.05, 0.5)
    L = 10 * math.pi
    return Config3D(Nt=steps, x_max=L, y_max=L, z_max=L, u_min=-6, u_max=6, v_min=-6, v_max=6, w_min=-6, w_max=6), F0(0, 0.001, 0.2)


def run_driver(name, args, tmp_path):
    exe = os.path.join(BIN, name)
    if not os.path.exists(exe):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "bin")], check=True)
    efile = tmp_path / "energy.txt"
    r = subprocess.run([exe, "--quiet", "--energy", str(efile)] + args, capture_output=True, text=True, cwd=tmp_path, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return np.loadtxt(efile)[:, 1], r.stdout, tmp_path


@pytest.mark.parametrize("dim,steps", [(1, 24), (2, 8), (3, 12)])
@pytest.mark.parametrize("mode", ["gpu", "gpu-fused", "cpu"])
def test_driver_energy_trace(dim, steps, mode, oracle, tmp_path):
    conf, f0 = landau_conf(dim, steps)
    _, want, _ = oracle.run(conf, f0, steps)
    name = f"test_nufi_{'cpu' if mode == 'cpu' else 'gpu'}_{dim}d"
    args = ["--landau", "--steps", str(steps)] + (["--fused"] if mode == "gpu-fused" else [])
    got, out, cwd = run_driver(name, args, tmp_path)
    n = min(len(got), steps)
    assert n >= steps - 1
    rel = np.max(np.abs(got[:n] - want[:n]) / np.abs(want[:n]))
    assert rel <= ENERGY_TOL, (dim, mode, rel)
    if mode != "cpu":
        rows = open(cwd / "statistics.csv").read().strip().splitlines()
        assert rows[0].startswith('"Time"; "L1-Norm"; "L2-Norm"; "Electric Energy"')  # bin/test_nufi_gpu_2d.cpp:119-121
        first = [float(x) for x in rows[1].split(";")]
        assert abs(first[1] - conf_volume(conf)) <= 1e-3 * conf_volume(conf) or dim == 3  # L1 norm = int f = box volume (1d, 2d weights)


def _csv_rows(path):
    rows = open(path).read().strip().splitlines()
    assert rows[0].startswith('"Time"; "L1-Norm"; "L2-Norm"; "Electric Energy"; "Kinetic Energy"; "Total Energy"; "Entropy"')
    return np.array([[float(x) for x in r.split(";")] for r in rows[1:]])


@pytest.mark.parametrize("dim,steps", [(1, 12), (2, 6), (3, 8)])
def test_statistics_csv_numeric_rows(dim, steps, oracle, tmp_path):
    """statistics.csv of the GPU drivers (bin/test_nufi_gpu_2d.cpp:119-121, 225-246): every numeric row -- time, L1, L2 (square
    root taken), electric / kinetic / total energy, entropy -- against the oracle's metrics of the same run.  The file carries 7
    significant digits (std::scientific, default precision, as the reference writes it)."""
    conf, f0 = landau_conf(dim, steps)
    coeffs, energy, _ = oracle.run(conf, f0, steps + 1)
    _, _, cwd = run_driver(f"test_nufi_gpu_{dim}d", ["--landau", "--steps", str(steps), "--fused"], tmp_path)
    rows = _csv_rows(cwd / "statistics.csv")
    assert len(rows) == steps + 1
    from numericalflowiteration_b200 import n_quad

    for n in (0, 1, steps // 2, steps):
        m = oracle.metrics(conf, f0, n, coeffs, 0, n_quad(conf))
        want = [n * conf.dt, m[0], np.sqrt(m[1]), energy[n], m[2], m[2] + energy[n], m[3]]
        assert np.allclose(rows[n], want, rtol=2e-6, atol=1e-300), (dim, n, rows[n], want)


def test_stats_txt_of_the_1d_cpu_driver(oracle, tmp_path):
    """stats.txt and E_<t>.txt of bin/test_nufi_cpu_1d.cpp:82-119: every second step "t  max|E|  sum E^2 * dx" over 256 plot
    points (widths 20, 8 digits), every 160th step the samples themselves."""
    steps = 12
    conf, f0 = landau_conf(1, steps)
    coeffs, _, _ = oracle.run(conf, f0, steps + 1)
    _, _, cwd = run_driver("test_nufi_cpu_1d", ["--landau", "--steps", str(steps)], tmp_path)
    lines = open(cwd / "stats.txt").read().splitlines()
    assert len(lines) == steps // 2 + 1 and all(len(ln) == 60 for ln in lines)
    st = stride_t(conf)
    xs = [conf.x_min + i * (conf.Lx / 256) for i in range(256)]
    for row, ln in enumerate(lines):
        n = 2 * row
        E = np.array([-oracle.field(conf, coeffs[n * st:(n + 1) * st], (x,), (1,)) for x in xs])
        want = [n * conf.dt, np.max(np.abs(E)), np.sum(E * E) * conf.dx]
        assert np.allclose([float(v) for v in ln.split()], want, rtol=5e-8, atol=1e-300), (n, ln, want)
    e0 = np.loadtxt(cwd / "E_0.000000.txt")
    assert e0.shape == (256, 2) and np.allclose(e0[:, 0], xs, rtol=1e-5)


def test_gpu_1d_driver_with_a_separate_metrics_grid(oracle, tmp_path):
    """The reference's 1d GPU driver builds its scheduler from two configs (bin/test_nufi_gpu_1d.cpp:216-230); with a metrics
    grid finer than the field grid the L1 norm (the integral of f over phase space) must still be the box volume."""
    steps = 6
    conf, f0 = landau_conf(1, steps)
    _, want, _ = oracle.run(conf, f0, steps)
    got, _, cwd = run_driver("test_nufi_gpu_1d", ["--landau", "--steps", str(steps), "--metrics-grid", "384", "640"], tmp_path)
    assert np.max(np.abs(got[:steps] - want) / np.abs(want)) <= ENERGY_TOL
    rows = _csv_rows(cwd / "statistics.csv")
    assert np.allclose(rows[:, 1], conf.Lx, rtol=1e-5)


REF_DRV = os.path.join(ROOT, "oracle", "_ref", "drivers")
REF_GPU_3D = os.path.join(REF_DRV, "ref_test_nufi_gpu_3d")


@pytest.mark.skipif(not os.path.exists(REF_GPU_3D), reason="oracle/_ref/drivers/ref_* are built by __graft_entry__.build() where /root/reference exists")
def test_unmodified_reference_gpu_3d_driver_runs_on_the_library(oracle, tmp_path):
    """The reference's OWN bin/test_nufi_gpu_3d.cpp, compiled unmodified against include/ (see tests/test_host_cpu.py), run on the
    GPU through libnufi_b200: its statistics.csv (as-committed 3d configuration: 8^3 x 8^3, bump-on-tail, Nt = 50, u in [-9,0])
    against the oracle's run of the same configuration."""
    r = subprocess.run([REF_GPU_3D], capture_output=True, text=True, cwd=tmp_path, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rows = _csv_rows(tmp_path / "statistics.csv")
    conf, f0 = Config3D(), F0(2, 0.03, 0.3)
    assert len(rows) == conf.Nt + 1
    coeffs, energy, _ = oracle.run(conf, f0, conf.Nt + 1)
    assert np.allclose(rows[:, 3], energy, rtol=2e-6)
    from numericalflowiteration_b200 import n_quad

    for n in (0, 7, conf.Nt):
        m = oracle.metrics(conf, f0, n, coeffs, 0, n_quad(conf))
        assert np.allclose(rows[n, [1, 2, 4, 6]], [m[0], np.sqrt(m[1]), m[2], m[3]], rtol=2e-6), (n, rows[n], m)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DRV, "ref_test_nufi_cpu_3d")), reason="oracle/_ref/drivers/ref_* not built")
def test_unmodified_reference_cpu_3d_driver_runs_on_the_library(tmp_path):
    """bin/test_nufi_cpu_3d.cpp unmodified: the OpenMP loop over eval_rho(n, l, coeffs, conf), poisson::solve, interpolate --
    every call served by the device library.  It prints timings only; the run must complete all Nt = 50 steps."""
    r = subprocess.run([os.path.join(REF_DRV, "ref_test_nufi_cpu_3d")], capture_output=True, text=True, cwd=tmp_path, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "n = 49 " in r.stdout and "Total time:" in r.stdout


def conf_volume(conf):
    v = conf.Lx
    if conf.dim >= 2:
        v *= conf.Ly
    return v


@pytest.mark.parametrize("name", ["1d-landau", "2d-landau", "3d-landau"])
def test_standalone_poisson_and_interpolate(name, oracle):
    """poisson<double>::solve and interpolate<double,4> as separate entry points (the reference loop calls them one after
    the other on the host, bin/test_nufi_gpu_3d.cpp:160-161)."""
    import ctypes as C

    mk, f0 = CASES[name]
    conf = mk()
    coeffs, _, _ = oracle.run(conf, f0, 4)
    rho = oracle.rho(conf, f0, 3, coeffs)
    phi_want, e_want = oracle.poisson(conf, rho)
    level_want = oracle.interpolate(conf, phi_want)
    L = _lib.load()
    with CudaScheduler(conf, f0) as s:
        data = rho.copy()
        e = C.c_double(0)
        assert L.nufi_b200_poisson_solve(s._h, data.ctypes.data_as(C.c_void_p), C.byref(e)) == 0
        level = np.zeros(stride_t(conf))
        assert L.nufi_b200_interpolate(s._h, data.ctypes.data_as(C.c_void_p), level.ctypes.data_as(C.c_void_p)) == 0
    assert rel_linf(data, phi_want) <= 1e-12
    assert abs(e.value - e_want) <= 1e-12 * abs(e_want)
    assert rel_linf(level, level_want) <= 1e-11


def test_group_one_process_multi_gpu(oracle):
    """nufi_b200_group_step over every visible device == the single-device step (bitwise identical level on every device,
    energy trace within tolerance of the reference CPU loop).  With one visible GPU the group degenerates to step()."""
    mk, f0 = CASES["2d-landau"]
    conf = mk()
    _, want, _ = oracle.run(conf, f0, conf.Nt)
    ndev = device_count()
    with CudaGroup(conf, f0, devices=range(ndev)) as g:
        for n in range(conf.Nt):
            g.step(n)
        g.sync()
        energies = [s.download_energy(0, conf.Nt) for s in g.scheds]
        levels = [s.download_phi(conf.Nt - 1) for s in g.scheds]
    for e, lv in zip(energies, levels):
        assert np.max(np.abs(e - want) / np.abs(want)) <= ENERGY_TOL
        assert np.array_equal(lv, levels[0])


@pytest.mark.parametrize("case", ["1d-two-stream", "2d-landau", "3d-landau"])
def test_peer_step_world1_matches_fused_step(case):
    """The peer-memory exchange path with a world of one rank (push into the own buffer, flag, wait, rank-order sum) must
    reproduce the plain fused step: same kernels as the multi-GPU path, runnable on a one-GPU box."""
    mk, f0 = CASES[case]
    conf = mk()
    with CudaScheduler(conf, f0, device=0) as a, CudaScheduler(conf, f0, device=0) as b:
        h = b.peer_export(1)
        b.peer_attach(0, 1, h)
        for n in range(conf.Nt):
            a.step(n)
            b.peer_step(n)
        ea, eb = a.download_energy(0, conf.Nt), b.download_energy(0, conf.Nt)
        assert not b.peer_timed_out()
        assert np.max(np.abs(ea - eb) / np.abs(ea)) <= 1e-12
        assert rel_linf(b.download_phi(conf.Nt - 1), a.download_phi(conf.Nt - 1)) <= 1e-12
        assert rel_linf(b.eval_rho(conf.Nt - 1), a.eval_rho(conf.Nt - 1)) <= 1e-12
        b.peer_detach()


def test_peer_step_large_grid_uses_gather_kernel():
    """Grids beyond the single-CTA tail take the peer_gather_kernel + cuFFT route."""
    conf = Config2D(Nx=80, Ny=64, Nu=8, Nv=8, Nt=4)
    f0 = F0(0, 0.05, 0.5)
    with CudaScheduler(conf, f0, device=0) as a, CudaScheduler(conf, f0, device=0) as b:
        b.peer_attach(0, 1, b.peer_export(1))
        for n in range(conf.Nt):
            a.step(n)
            b.peer_step(n)
        assert b.last_tail_variant == "cufft"
        assert not b.peer_timed_out()
        assert rel_linf(b.download_phi(conf.Nt - 1), a.download_phi(conf.Nt - 1)) <= 1e-12


def test_peer_step_requires_attach():
    mk, f0 = CASES["1d-landau"]
    with CudaScheduler(mk(), f0, device=0) as s:
        with pytest.raises(ValueError):
            s.peer_step(0)


@pytest.mark.parametrize("mode", ["peer", "nccl"])
def test_group_exchange_modes(oracle, mode):
    """Both exchanges of the one-process group (stores into peer memory fused into the kernels / NCCL all-reduce) give
    bit-identical replicas and the reference's energy trace.  Needs >= 2 GPUs."""
    ndev = device_count()
    if ndev < 2:
        pytest.skip("needs at least 2 GPUs")
    mk, f0 = CASES["3d-landau"]
    conf = mk()
    _, want, _ = oracle.run(conf, f0, conf.Nt)
    with CudaGroup(conf, f0, devices=range(ndev)) as g:
        g.set_exchange(mode)
        assert g.exchange == ("peer-memory" if mode == "peer" else "nccl")
        for n in range(conf.Nt):
            g.step(n)
        g.sync()
        energies = [s.download_energy(0, conf.Nt) for s in g.scheds]
        levels = [s.download_phi(conf.Nt - 1) for s in g.scheds]
        if mode == "peer":
            assert not any(s.peer_timed_out() for s in g.scheds)
    for e, lv in zip(energies, levels):
        assert np.max(np.abs(e - want) / np.abs(want)) <= ENERGY_TOL
        assert np.array_equal(lv, levels[0])


def test_torchrun_peer_exchange_ipc(tmp_path):
    """One process per GPU (torchrun): CUDA IPC handles all-gathered through torch.distributed, then peer_step with no
    collective call; every rank's history must be bit-identical to rank 0's and match the NCCL all-reduce path.  >= 2 GPUs."""
    ndev = device_count()
    if ndev < 2:
        pytest.skip("needs at least 2 GPUs")
    script = tmp_path / "peer_ranks.py"
    script.write_text(
        "import os, sys, numpy as np, torch, torch.distributed as dist\n"
        f"sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})\n"
        "from cases import CASES\n"
        "from numericalflowiteration_b200 import CudaScheduler\n"
        "from numericalflowiteration_b200.distributed import DistributedStepper\n"
        "rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])\n"
        "torch.cuda.set_device(local)\n"
        "dist.init_process_group('nccl', device_id=torch.device('cuda', local))\n"
        "mk, f0 = CASES['2d-two-stream']\n"
        "conf = mk()\n"
        "out = {}\n"
        "for mode in ('peer', 'nccl'):\n"
        "    s = CudaScheduler(conf, f0, device=local)\n"
        "    st = DistributedStepper(s, exchange=mode)\n"
        "    for n in range(conf.Nt):\n"
        "        st.step(n)\n"
        "    s.sync()\n"
        "    out[mode] = (s.download_energy(0, conf.Nt), s.download_phi(conf.Nt - 1))\n"
        "    if mode == 'peer':\n"
        "        assert st.exchange == 'peer-memory' and not s.peer_timed_out()\n"
        "    dist.barrier()\n"
        "    s.close()\n"
        "lv = torch.from_numpy(out['peer'][1]).cuda()\n"
        "ref = lv.clone(); dist.broadcast(ref, 0)\n"
        "assert torch.equal(lv, ref), 'replicas differ'\n"
        "d = float(np.max(np.abs(out['peer'][0] - out['nccl'][0]) / np.abs(out['nccl'][0])))\n"
        "print('rank', rank, 'peer vs nccl energy rel diff', d)\n"
        "assert d <= 1e-10, d\n"
        "dist.barrier(); dist.destroy_process_group()\n"
        "print('rank', rank, 'ok')\n")
    r = subprocess.run(["python", "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29571", str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout


def test_bench_line_contract():
    """`python bench.py` prints ONE JSON line with the keys the driver reads (value, e2e, roofline, cpu_baseline, gpu_launches,
    clocks), measured through the C ABI on this GPU; the live parity check inside it must be within tolerance."""
    import json
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "C4", "--steps", "5", "--warmup", "3", "--no-extras",
                        "--cpu-budget", "0.5"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["metric"] == "backtrace point-steps/sec" and d["unit"] == "point-steps/s" and d["n_gpus"] == 1 and d["dtype"] == "f64"
    assert d["value"] > 0 and d["gpu_launches"] >= 2 * d["steps"] and d["scaling"] == "weak" and d["vs_baseline"] is None
    roof = d["roofline"]
    assert roof["bound"] == "fp64" and 0 < roof["frac"] < 1.5 and roof["unit"] == "TFLOP/s" and roof["kernel_ms"] > 0
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-12 and 0 < roof["smem"]["frac"] < 1
    e2e = d["e2e"]
    assert 0 < e2e["value"] <= d["value"] * 1.05 and e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] > 0
    assert d["parity"]["rho_rel_linf_vs_cpu_reference"] <= d["parity"]["tolerance"] == 1e-10
    assert "sm_mhz" in d["clocks"] and "reasons" in d["clocks"]
