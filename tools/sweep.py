#!/usr/bin/env python
"""Tuning sweep (development tool): backtrace kernel time of one workload for a grid of shape overrides.
    python tools/sweep.py C2 --ilp 1 2 --W 15 16 28 31 --lc 0 4 8
Overrides go through the NUFI_B200_ILP / NUFI_B200_W / NUFI_B200_LC environment variables read by launch_backtrace."""
import argparse
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import FLOP_PER_POINT_STEP, GpuRunner, free_run, make_workload  # noqa: E402
from numericalflowiteration_b200 import n_quad  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("--depth", type=int, default=0)
    ap.add_argument("--ilp", type=int, nargs="*", default=[0])
    ap.add_argument("--W", type=int, nargs="*", default=[0])
    ap.add_argument("--lc", type=int, nargs="*", default=[0])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--tn", type=int, nargs="*", default=[0], help="nodes per tile (NUFI_B200_TN): 32, 16, 8, 4, 2, 1")
    ap.add_argument("--xpp", type=int, default=-1, help="level format of 2d/3d histories: 0 B-spline, 1 xpp, -1 auto")
    a = ap.parse_args()
    conf, f0, depth, desc = make_workload(a.workload, 1)
    n = a.depth or depth
    if a.xpp >= 0:
        os.environ["NUFI_B200_XPP"] = str(a.xpp)
    torch.cuda.set_device(0)
    r = GpuRunner(conf, f0, 0, 1, torch, None)
    free_run(r, n)
    r.s.set_kernel_timing(True)
    ps = float(n_quad(conf)) * n
    print(desc, "depth", n, flush=True)
    for ilp, W, lc, tn in itertools.product(a.ilp, a.W, a.lc, a.tn):
        for k, v in (("NUFI_B200_ILP", ilp), ("NUFI_B200_W", W), ("NUFI_B200_LC", lc), ("NUFI_B200_TN", tn)):
            if v:
                os.environ[k] = str(v)
            else:
                os.environ.pop(k, None)
        try:
            for _ in range(2):
                r.s.compute_rho(n, 0, r.s.n_quad)
            r.s.sync()
            r.s.backtrace_time(reset=True)
            for _ in range(a.reps):
                r.s.compute_rho(n, 0, r.s.n_quad)
            ms, cnt = r.s.backtrace_time(reset=True)
            ms /= cnt
            print(f"ilp={ilp} W={W} lc={lc} tn={tn}: {r.s.last_variant:34s} {ms:9.4f} ms  {ps / ms / 1e6:8.2f} Gps/s  "
                  f"{ps * FLOP_PER_POINT_STEP[conf.dim] / ms / 1e9:6.2f} TF", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"ilp={ilp} W={W} lc={lc}: {e}", flush=True)


if __name__ == "__main__":
    main()
