"""GPU: whose rounding is it?  The reference adds f over the velocity nodes of a spatial node sequentially in double
(nufi/rho.hpp:299-306): with Nu*Nv*Nw = 262 144 terms (C5-64) that sum carries ~sqrt(N) eps of relative error in dV*sum f, i.e.
~6e-14 / alpha = 6e-11 relative to the density perturbation (alpha = 1e-3) -- two thirds of the 1e-10 parity tolerance, and not
the device's doing.  The device accumulates with a compensated (two-sum) addition; against the oracle's yardstick that carries
the same sum in long double (oracle/nufi_oracle.c: orc_rho_sweep_extended) it must be an order of magnitude closer than the
reference's own double sum is."""
import math

import numpy as np
import pytest

from numericalflowiteration_b200 import Config3D, CudaScheduler, F0, stride_t

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nx", [32, 64])
def test_c5_rho_against_the_extended_precision_sum(nx, oracle):
    L = 10 * math.pi
    conf = Config3D(Nx=nx, Ny=nx, Nz=nx, Nu=nx, Nv=nx, Nw=nx, Nt=2, x_max=L, y_max=L, z_max=L, u_min=-6, u_max=6, v_min=-6, v_max=6,
                    w_min=-6, w_max=6)
    f0 = F0(0, 0.001, 0.2)
    n = 2
    st = stride_t(conf)
    with CudaScheduler(conf, f0, device=0) as s:
        for m in range(n):
            s.step(m)
        rho = s.eval_rho(n)
        hist = np.concatenate([s.download_phi(m) for m in range(n)] + [np.zeros(st)])
        variant = s.last_variant
    l_n = nx  # one x-row of nodes: nx * nx^3 * n point-steps on the CPU, twice
    ext = oracle.rho_extended(conf, f0, n, hist, 0, l_n)[:l_n]
    ref = oracle.rho(conf, f0, n, hist, 0, l_n)[:l_n]
    scale = float(np.max(np.abs(ext)))
    err_gpu = float(np.max(np.abs(rho[:l_n] - ext))) / scale
    err_ref = float(np.max(np.abs(ref - ext))) / scale
    err_gpu_ref = float(np.max(np.abs(rho[:l_n] - ref))) / scale
    print(f"C5-{nx} depth {n} [{variant}]: rho rel-Linf  device vs long-double sum {err_gpu:.2e}   reference (double, sequential) vs "
          f"long-double sum {err_ref:.2e}   device vs reference {err_gpu_ref:.2e}")
    assert err_gpu <= 1e-12
    assert err_gpu_ref <= 1e-10
