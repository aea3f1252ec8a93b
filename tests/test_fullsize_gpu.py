"""Free runs at BASELINE.json's FULL sizes on the GPU against the committed electric-energy traces of the real reference
(tests/golden/fullsize_<case>.npz, made by tests/golden/make_fullsize_traces.py from oracle/_ref in the build container).
north_star gate: electric-energy trace relative error <= 1e-8 over the run; rho relative L-inf <= 1e-10 per step is checked on
the last step's rho (free-running, so it carries the accumulated difference of the whole run -- gate 1e-8 like the energy)."""
import os

import numpy as np
import pytest

from cases import rel_linf

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ENERGY_TOL = 1e-8


def _fixture(name):
    path = os.path.join(HERE, "golden", f"fullsize_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not committed")
    return np.load(path)


@pytest.mark.parametrize("name", ["C1", "C2", "C3", "C4"])
def test_fullsize_energy_trace(name):
    import sys

    sys.path.insert(0, os.path.dirname(HERE))
    from bench import make_workload
    from numericalflowiteration_b200 import CudaScheduler

    g = _fixture(name)
    conf, f0, _, desc = make_workload(name, 1)
    assert str(g["workload"]) == desc and int(g["f0_kind"]) == f0.kind and list(g["f0_p"]) == list(f0.p)
    nt = int(g["steps"])
    want = g["energy"]
    with CudaScheduler(conf, f0, device=0) as s:
        for n in range(nt):
            s.step(n)
        got = s.download_energy(0, nt)
        level = s.download_phi(nt - 1)
        rho = s.eval_rho(nt - 1)
    err = np.abs(got - want) / np.abs(want)
    print(f"{name}: {nt} steps, energy rel err max {err.max():.3e} (at step {int(err.argmax())}), last level rel-Linf "
          f"{rel_linf(level, g['level_last']):.3e}, last rho rel-Linf {rel_linf(rho, g['rho_last']):.3e}")
    assert err.max() <= ENERGY_TOL
    assert rel_linf(level, g["level_last"]) <= ENERGY_TOL
    assert rel_linf(rho, g["rho_last"]) <= ENERGY_TOL
