"""Electric-energy traces of the REAL reference (oracle/_ref, the reference's own headers compiled in place from
/root/reference) at BASELINE.json's full sizes -- the free-running CPU loop of bin/test_nufi_cpu_{1,2,3}d.cpp.  Run in the
build container only (minutes of CPU per case; /root/reference does not exist on the GPU box); the traces are committed as
tests/golden/fullsize_<case>.npz and checked against the GPU's free run by tests/test_fullsize_gpu.py
(north_star: electric-energy trace relative error <= 1e-8 over the run).

    python tests/golden/make_fullsize_traces.py [C1 C2 C3 C4]

The `_fast` build (-O3 -march=x86-64-v3) makes the main traces: same source, FMA contraction allowed -- its rounding differs
from the canonical -O2 -ffp-contract=off build at the 1e-16 level per operation.  `--canonical` writes a second trace
(fullsize_<case>_canonical.npz) with that build: the spread between the two is the reference's OWN sensitivity to rounding
over the run (chaotic two-stream: it grows to O(1e-2); weak Landau: the signal decays 13 orders of magnitude into the rounding
floor), which bounds how tightly any other implementation can be compared at each step.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from bench import make_workload  # noqa: E402
from oracle.oracle_py import Reference  # noqa: E402

# steps of the free run per case: the whole run where the CPU can afford it, a window otherwise (C3: 16.8 M points per step)
STEPS = {"C1": 1600, "C2": 1600, "C3": 40, "C4": 50}


def main():
    args = sys.argv[1:]
    canonical = "--canonical" in args  # the reference-like -O2 -ffp-contract=off build: a second, independently rounded trace
    names = [a for a in args if not a.startswith("--")] or ["C4", "C3", "C1", "C2"]
    variant = "" if canonical else ("_fast" if Reference.available("_fast") else "")
    suffix = "_canonical" if canonical else ""
    ref = Reference(variant)
    for name in names:
        conf, f0, _, desc = make_workload(name, 1)
        nt = STEPS[name]
        t0 = time.time()
        coeffs, energy, rho = ref.run(conf, f0, nt)
        dt = time.time() - t0
        st = coeffs.size // nt
        np.savez_compressed(os.path.join(HERE, f"fullsize_{name}{suffix}.npz"), energy=energy, steps=np.int64(nt), rho_last=rho,
                            level_last=coeffs[(nt - 1) * st: nt * st], workload=np.array(desc), f0_kind=np.int64(f0.kind),
                            f0_p=np.array(list(f0.p)), build=np.array(os.path.basename(ref.path)), threads=np.int64(ref.threads()),
                            seconds=np.float64(dt))
        print(f"{name}: {nt} steps in {dt:.1f} s on {ref.threads()} threads ({os.path.basename(ref.path)}); energy[0]={energy[0]:.16e} "
              f"energy[-1]={energy[-1]:.16e}", flush=True)


if __name__ == "__main__":
    main()
