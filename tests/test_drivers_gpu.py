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
