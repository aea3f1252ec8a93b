#!/usr/bin/env python
"""bench.py -- backtrace point-steps/s of the NuFI hot path on B200 (driver contract: see the task statement).

A *step* is one NuFI time step at a fixed history depth n: trace every quadrature point back through the n stored
levels, f0 at the foot, reduce into rho, Poisson + spline interpolation into level n -- all on the device, through
the C ABI of libnufi_b200.so.  One step costs Nquad * n point-steps.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C1|C2|C3|C4|C5-16|C5-32] [--depth n] [--scaling weak|strong]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (N > 1: one rank per GPU)
    python bench.py --impl reference ...      the reference's own CPU eval_rho (oracle/_ref) on the host cores

The JSON line carries, beside the contract's keys: `parity` (live rho check against the reference's CPU eval_rho on the same
history, at every N; N > 1 adds a checksum proving all replicas hold bit-identical rho and level n), `other_workloads`
(1d/2d/3d at N = 1; the 3d configurations of BASELINE.json in STRONG and weak form at N > 1, each with its own parity),
`full_run` (complete free runs, N = 1).

Only the cpu_baseline / parity legs and --impl reference execute anything under oracle/ (as the measured CPU baseline and as
the live parity check); the GPU numbers never touch it.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_PER_POINT_STEP = {1: 30.0, 2: 138.0, 3: 431.0}  # SURVEY.md section 8d / App. A.4 (algorithmic, FMA = 2)
# dram__bytes_read.sum + dram__bytes_write.sum of ONE backtrace launch at the default depth, from the committed ncu --set full
# captures of the kernels as shipped (profiles/r02_ncu_full_summary.csv); the algorithmic figure is n*level_bytes + 16*Nnodes
NCU_DRAM_TRAFFIC_BYTES = {"C1": 4.97e6, "C2": 4.97e6, "C3": 3.65e6, "C4": 1.62e6, "C5-16": 1.45e6, "C5-32": 8.78e6}
SMEM_BYTES_PER_POINT_STEP = {1: 24.0, 2: 128.0, 3: 512.0}  # coefficients a point gathers per level
SMEM_BYTES_PER_CLK_SM = 128.0  # shared-memory data pipe; tools/microbench.cu measures 125-127 on this GPU
# FP64 ceiling of each kernel's hot loop once the register-file operand reads of its DFMAs are counted (a DFMA with 1 / 2 / 3
# fresh 64-bit register operands issues every 2.00 / 2.17 / 3.00 cycles, profiles/r02_microbench.txt), as a fraction of the DFMA
# peak; tools/sass_rf_model.py on the built library -> profiles/r02_sass_rf_model.txt.  Keys: (dim, xpp level format)
RF_CEILING = {(1, False): 0.848, (2, True): 0.809, (2, False): 0.825, (3, True): 0.780, (3, False): 0.793}
L2_FLUSH_BYTES = 256 << 20
RHO_TOL = 1e-10
LEAD_IN = 3  # untimed iterations enqueued directly in front of the timed ones (see timed_region)
SETTLE_STEPS = 10  # untimed iterations in front of every timed region (at least the W asked for): the first steps after an idle
# gap run 1-3 % slower (clock / power ramp), config.settle_steps states it
ALIGN_RANKS = True  # N > 1: a stream-ordered NCCL barrier between the (untimed) L2 flush and every timed step


def make_workload(name: str, n_gpus: int, scaling: str = "weak"):
    """(conf, f0, depth n, description).  weak: the velocity quadrature grows with the GPU count so the quadrature points per
    GPU stay fixed (BASELINE.json configs[4]: 'larger Nx and Nv quadrature'); strong: the grid BASELINE names, whatever N."""
    from numericalflowiteration_b200 import Config1D, Config2D, Config3D, F0

    g = max(n_gpus, 1) if scaling == "weak" else 1
    if name in ("C1", "C2"):
        conf = Config1D(Nu=512 * g)  # 256 x 512, dt = 1/16, Nt = 1600 (nufi/config.hpp:58-65)
        f0 = F0(0, 0.01, 0.5) if name == "C1" else F0(1, 0.01, 0.5)
        what = "1d1v weak Landau" if name == "C1" else "1d1v two-stream, long horizon"
        return conf, f0, 800, f"{name}: {what} Nx=256 Nu={conf.Nu} dt=1/16 Nt=1600"
    if name == "C3":
        conf = Config2D(Nv=128 * g)  # 32^2 x 128^2, Nt = 800 (nufi/config.hpp:120-130)
        return conf, F0(0, 0.05, 0.5), 100, f"C3: 2d2v weak Landau 32^2 x 128x{conf.Nv} dt=1/16 Nt=800"
    L = 10 * math.pi
    box = dict(x_max=L, y_max=L, z_max=L, u_min=-6, u_max=6, v_min=-6, v_max=6, w_min=-6, w_max=6)
    if name == "C4":
        conf = Config3D(Nw=8 * g, **box)  # 8^3 x 8^3, dt = 0.1, Nt = 50 (nufi/config.hpp:199-208)
        return conf, F0(0, 0.001, 0.2), 50, f"C4: 3d3v weak Landau 8^3 x 8x8x{conf.Nw} dt=0.1 Nt=50"
    if name.startswith("C5-"):
        nx = int(name[3:])
        conf = Config3D(Nx=nx, Ny=nx, Nz=nx, Nu=nx, Nv=nx, Nw=nx * g, **box)
        return conf, F0(0, 0.001, 0.2), 25, f"{name}: 3d3v weak Landau {nx}^3 x {nx}x{nx}x{conf.Nw} dt=0.1 Nt=50"
    raise SystemExit(f"unknown workload {name}")


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    BAD = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "hw_power_brake": 0x80}
    NOTE = {"sw_power_cap": 0x4}

    def __init__(self, index: int, period_s: float = 0.004):
        self.samples, self.reasons, self.sm_max, self.ok = [], set(), None, False
        self.period = period_s
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in {**self.BAD, **self.NOTE}.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.ok:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ reference arm
def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def reference_impl():
    """The reference's own OpenMP rho sweep (oracle/_ref, -O3 AVX2+FMA build; the C port when oracle/_ref is absent), with the
    OpenMP thread count set EXPLICITLY to the host's cores (torchrun exports OMP_NUM_THREADS=1 to every rank)."""
    from oracle.oracle_py import Oracle, Reference

    impl = Reference("_fast") if Reference.available("_fast") else Oracle(fast=True)
    impl.set_threads(host_cores())
    return impl


def reference_sweep_timer(conf, f0, n, coeffs, budget_s: float, impl=None):
    """Times the reference sweep on a bounded sample of spatial nodes.  Returns (callable running one sample -> seconds,
    point-steps per sample, description, kind, threads, nodes in the sample, dict receiving rho of the sample nodes)."""
    from numericalflowiteration_b200 import n_nodes, n_vel

    impl = impl or reference_impl()
    nn, nv = n_nodes(conf), n_vel(conf)
    threads = impl.threads()
    # calibrate on a few nodes, then size the sample for the budget
    l_cal = min(nn, max(threads, 8))
    t0 = time.perf_counter()
    impl.rho(conf, f0, n, coeffs, 0, l_cal)
    t_cal = max(time.perf_counter() - t0, 1e-4)
    per_node = t_cal / l_cal
    l_n = int(min(nn, max(threads, budget_s / per_node)))
    l_n = max(threads, (l_n // threads) * threads) if l_n < nn else nn
    l_n = min(l_n, nn)
    desc = (f"reference eval_rho sweep ({impl.kind}: {os.path.basename(impl.path)}, OpenMP {threads} threads) over spatial "
            f"nodes [0,{l_n}) of {nn} at depth n={n}: {l_n * nv * n:.3e} point-steps per sample")
    out = {}

    def run():
        t0 = time.perf_counter()
        out["rho"] = impl.rho(conf, f0, n, coeffs, 0, l_n)
        return time.perf_counter() - t0

    return run, float(l_n) * nv * max(n, 1), desc, impl.kind, threads, l_n, out


def reference_history(name: str, n: int):
    """The history the reference arm's timed step reads: the reference CPU loop's own free run of the workload's BASE grid
    (what the GPU arm builds on the device, to parity tolerance), computed once with all host cores and cached under /tmp; for
    workloads whose CPU free run would take more than a few minutes (C3, C5) a synthetic smooth history instead."""
    from numericalflowiteration_b200 import n_quad
    from oracle.oracle_py import Oracle, synthetic_history

    conf, f0, _, _ = make_workload(name, 1)
    cost = float(n_quad(conf)) * n * (n - 1) / 2
    if cost > 6e10:
        return synthetic_history(conf, n, seed=1234, amp=1e-2), "synthetic smooth sine potentials (numpy), seed 1234 (a CPU free run would take too long)"
    cache = f"/tmp/nufi_b200_refhist_{name}_{n}.npy"
    if os.path.exists(cache):
        try:
            return np.load(cache), f"free run of the reference CPU loop on the base grid to depth {n} (cached: {cache})"
        except Exception:  # noqa: BLE001
            pass
    orc = Oracle(fast=True)
    orc.set_threads(host_cores())
    t0 = time.perf_counter()
    coeffs, _, _ = orc.run(conf, f0, n)
    dt = time.perf_counter() - t0
    try:
        np.save(cache, coeffs)
    except Exception:  # noqa: BLE001
        pass
    return coeffs, f"free run of the reference CPU loop on the base grid to depth {n} ({dt:.0f} s on {orc.threads()} threads, untimed)"


def reference_cuda_leg(conf, f0, n, coeffs_host, sched, device, reps=3):
    """Informational (north star: 'the reference's existing CUDA path'): the reference's own nufi/cuda_kernel.cu, compiled
    unmodified for sm_100a (oracle/_ref/libnufi_refcuda.so), timed with CUDA events on the same GPU, same history, same step.
    Its f0 is the one committed in the reference's config.hpp; where that is this workload's f0 (C2) rho is compared too."""
    try:
        from oracle.oracle_py import ReferenceCuda
        from numericalflowiteration_b200 import n_quad

        if not ReferenceCuda.available():
            return {"unavailable": "oracle/_ref/libnufi_refcuda.so not built"}
        rc = ReferenceCuda(conf, device)
        rc.upload(coeffs_host, n)
        rho, ms = rc.rho(n, reps)
        rc.close()
        out = {"value": float(n_quad(conf)) * n / (ms * 1e-3), "unit": "point-steps/s", "kernel_ms": ms, "reps": reps,
               "what": "reference nufi/cuda_kernel.cu (cuda_kernel<double,4>::compute_rho: memset + cuda_eval_rho, 64-thread blocks, "
                       "atomicAdd), compiled unmodified for sm_100a, CUDA-event timed on ONE GPU at the same depth"}
        same_f0 = conf.dim == 1 and f0.kind == 1 and list(f0.p)[:2] == [0.01, 0.5]
        if same_f0 and sched is not None:
            ours = sched.eval_rho(n)
            out["rho_rel_linf_ours_vs_reference_cuda"] = float(np.max(np.abs(ours - rho)) / np.max(np.abs(rho)))
        return out
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    conf, f0, depth, desc = make_workload(args.workload, args.gpus, args.scaling)
    n = args.depth or depth
    coeffs, hist_desc = reference_history(args.workload, n)
    total = max(args.steps + args.warmup, 1)
    budget = min(2.0, max(0.05, 90.0 / total))
    run, psteps, sdesc, kind, threads, _, _ = reference_sweep_timer(conf, f0, n, coeffs, budget)
    for _ in range(args.warmup):
        run()
    times = [run() for _ in range(args.steps)]
    t = float(np.sum(times))
    value = psteps * args.steps / t
    line = {
        "impl": "reference", "metric": "backtrace point-steps/sec", "value": value, "unit": "point-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "depth_n": n, "history": hist_desc, "host_cores": host_cores(), "omp_threads": threads},
        "cpu_baseline": {"value": value, "unit": "point-steps/s", "cores": threads, "kind": kind, "sample": sdesc},
        "e2e": {"value": value, "unit": "point-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    return line


# ------------------------------------------------------------------------------------------------ GPU arm
class GpuRunner:
    """One rank's scheduler + (for N > 1) the exchange of the partial rho: peer memory fused into the kernels, or NCCL."""

    def __init__(self, conf, f0, rank, world, torch, dist, exchange="peer"):
        from numericalflowiteration_b200 import CudaScheduler, partition

        self.torch, self.dist, self.rank, self.world = torch, dist, rank, world
        self.s = CudaScheduler(conf, f0, device=torch.cuda.current_device())
        self.stream = torch.cuda.Stream()
        self.s.set_stream(self.stream.cuda_stream)
        self.q0, self.q1 = partition(self.s.n_quad, world, rank)
        self.exchange, self.exchange_note = "single", None
        if world > 1:
            from numericalflowiteration_b200.distributed import _alias_device_f64

            self.rho_t = _alias_device_f64(torch, self.s.rho_device_ptr(), self.s.n_nodes)
            self.exchange = "nccl"
            if exchange == "peer":  # map every rank's exchange buffer into every rank (CUDA IPC); all ranks or none
                ok, why = 1, None
                try:
                    mine = self.s.peer_export(world)
                except Exception as e:  # noqa: BLE001
                    ok, why, mine = 0, repr(e), b"\0" * 64
                handles = [None] * world
                dist.all_gather_object(handles, mine)
                if ok:
                    try:
                        self.s.peer_attach(rank, world, b"".join(handles))
                    except Exception as e:  # noqa: BLE001
                        ok, why = 0, repr(e)
                t = torch.tensor([ok], device="cuda", dtype=torch.int32)
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                if int(t.item()) == 1:
                    self.exchange = "peer-memory"
                else:
                    self.exchange_note = f"peer-memory exchange unavailable ({why or 'failed on another rank'}); NCCL all-reduce used"
                    print(f"[bench rank {rank}] {self.exchange_note}", file=sys.stderr)

    def step(self, n):
        """Fused step, no host round trip.  N > 1: local shard -> exchange of the partial rho (stores into peer memory fused
        into the slot reduction + flags; or NCCL all-reduce in place) -> replicated tail."""
        if self.world == 1:
            self.s.step(n)
        elif self.exchange == "peer-memory":
            self.s.peer_step(n)
        else:
            self.s.compute_rho(n, self.q0, self.q1)
            with self.torch.cuda.stream(self.stream):
                self.dist.all_reduce(self.rho_t, op=self.dist.ReduceOp.SUM)
            self.s.field_tail_device(n, self.rho_t.data_ptr())

    def my_point_steps(self, n):
        """Point-steps this rank's backtrace kernel traces in one step at depth n."""
        if self.world > 1 and self.exchange == "peer-memory":  # velocity nodes rank, rank+world, ... of every spatial node
            return float(self.s.n_nodes) * len(range(self.rank, self.s.n_quad // self.s.n_nodes, self.world)) * n
        return float(self.q1 - self.q0) * n

    def close(self):
        self.stream.synchronize()
        if self.world > 1:
            self.dist.barrier()  # every rank is done with every peer's exchange buffer
        self.s.close()


def timed_region(runner, n, steps, warmup, flush, torch, dist, sampler=None, align=None):
    """`warmup` untimed steps, then `steps` fused steps at depth n, each bracketed by a CUDA event pair on the stream the
    library launches on, with an (untimed) L2 flush in front of each.  Returns (sum of step times in ms, max over ranks; launches)."""
    s, st = runner.s, runner.stream
    for _ in range(max(warmup, SETTLE_STEPS)):  # untimed, same shape as the timed iterations below
        with torch.cuda.stream(st):
            if flush is not None:
                flush.fill_(1.0)
            if align is not None:
                dist.all_reduce(align)
        runner.step(n)
    st.synchronize()
    if runner.world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    l0 = s.launches
    ctx = sampler if sampler is not None else _Null()
    with ctx:
        # LEAD_IN untimed iterations are enqueued in the same burst in front of the timed ones: the first step after the
        # synchronisation above starts from an idle GPU and an empty launch queue (no tail to launch behind) and runs 3-8 % slower
        # than every later one -- a start-up cost of the measurement, not of a run's steps (config.lead_in_steps)
        for i in range(LEAD_IN + steps):
            with torch.cuda.stream(st):
                if flush is not None:
                    flush.fill_(1.0)  # L2 flush, untimed
                if align is not None:  # untimed: the ranks leave the flush together (its duration jitters by microseconds), as
                    dist.all_reduce(align)  # they do in a run without flushes, where every step ends with the exchange
            if i < LEAD_IN:
                runner.step(n)
                if i == LEAD_IN - 1:
                    l0 = s.launches
                continue
            a, b = ev[i - LEAD_IN]
            a.record(st)
            runner.step(n)
            b.record(st)
        st.synchronize()
        if runner.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    launches = s.launches - l0
    per_step = [float(a.elapsed_time(b)) for a, b in ev]
    timed_region.last_per_step_ms = per_step
    t_ms = float(sum(per_step))
    if runner.world > 1:
        tt = torch.tensor([t_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_ms = float(tt.item())
    return t_ms, int(launches)


def measure_gpu(runner, n, steps, warmup, flush, torch, dist, sampler=None, align=None):
    """Two timed regions of the same K steps: (1) as a user runs it -> `value`; (2) with the library's per-launch CUDA
    event pair around every backtrace kernel -> the kernel's average duration for the roofline (the events sit between
    the kernel and the field tail and keep the tail from launching programmatically behind it, so region 2 is slightly
    slower per step; its own step time is reported beside the kernel time)."""
    s = runner.s
    s.set_kernel_timing(False)
    if align is None and runner.world > 1 and ALIGN_RANKS:
        align = torch.zeros(1, device="cuda")
    t_ms, launches = timed_region(runner, n, steps, warmup, flush, torch, dist, sampler, align)
    per_step = list(timed_region.last_per_step_ms)
    s.set_kernel_timing(True)
    s.backtrace_time(reset=True)
    t2_ms, _ = timed_region(runner, n, steps, 3, flush, torch, dist, None, align)
    bt_total, bt_count = s.backtrace_time(reset=True)  # warm-up launches included: same kernel, same inputs
    s.set_kernel_timing(False)
    bt_ms = bt_total / max(bt_count, 1)
    bt_all = [bt_ms]
    if runner.world > 1:
        g = [torch.zeros(1, device="cuda", dtype=torch.float64) for _ in range(runner.world)]
        dist.all_gather(g, torch.tensor([bt_ms], device="cuda", dtype=torch.float64))
        bt_all = [float(x.item()) for x in g]
    return {"t_ms": t_ms, "t_ms_kernel_timing": t2_ms, "bt_ms": bt_ms, "bt_ms_per_rank": bt_all, "launches": int(launches),
            "per_step_ms": per_step}


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def measure_e2e_step_host(runner, n, steps, warmup, coeffs_host, torch, dist):
    """The call a user with a HOST-resident history makes per time step: nufi_b200_step_host -- level n-1 host -> device, fused
    step (N > 1: with the peer-memory exchange), level n + rho + energy device -> host, one synchronisation.  Wall clock."""
    s = runner.s
    rho = np.zeros(s.n_nodes)
    peer = runner.world > 1 and runner.exchange == "peer-memory"
    if runner.world > 1 and not peer:
        return None
    h2d = s.stride_t * 8
    d2h = s.stride_t * 8 + s.n_nodes * 8 + 8
    for _ in range(warmup):
        s.step_host(n, coeffs_host, rho, peer=peer)
    torch.cuda.synchronize()
    if runner.world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        s.step_host(n, coeffs_host, rho, peer=peer)
    torch.cuda.synchronize()
    if runner.world > 1:
        dist.barrier()
    t = time.perf_counter() - t0
    if runner.world > 1:
        tt = torch.tensor([t], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt.item())
    return t, h2d, d2h


def measure_e2e(runner, n, steps, warmup, coeffs_host, torch, dist):
    """The reference GPU driver's per-step call sequence (bin/test_nufi_gpu_3d.cpp:154-162) through the C ABI with HOST
    buffers: upload_phi(n-1) [H2D] -> compute_rho(n, q-range) -> download_rho [D2H] -> (N > 1: all-reduce) ->
    solve + interpolate from the host rho [H2D] -> energy [D2H]."""
    s = runner.s
    rho = np.zeros(s.n_nodes)
    h2d = s.stride_t * 8 + s.n_nodes * 8
    d2h = s.n_nodes * 8 + 8

    def one():
        if n > 0:
            s.upload_phi(n - 1, coeffs_host)
        rho[:] = 0.0
        s.compute_rho(n, runner.q0, runner.q1)
        s.download_rho(rho)
        if runner.world > 1:
            t = torch.from_numpy(rho).cuda()
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            rho[:] = t.cpu().numpy()
        return s.solve_interpolate(n, rho=1.0 + rho)

    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    if runner.world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e = one()
    torch.cuda.synchronize()
    if runner.world > 1:
        dist.barrier()
    t = time.perf_counter() - t0
    if runner.world > 1:
        tt = torch.tensor([t], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt.item())
    return t, h2d, d2h, e


def free_run(runner, n_levels):
    """Build the history on the device: fused steps 0 .. n_levels-1 (no host round trip)."""
    for m in range(n_levels):
        runner.step(m)
    runner.stream.synchronize()


def download_history(s, n):
    from numericalflowiteration_b200 import stride_t

    st_ = stride_t(s.conf)
    out = np.zeros((n + 1) * st_)
    for lvl in range(n):
        out[lvl * st_:(lvl + 1) * st_] = s.download_phi(lvl)
    return out


def live_parity(runner, conf, f0, n, coeffs_host, cpu_budget, dist, want_cpu_baseline):
    """The rho the last fused step at depth n consumed (N > 1: the sum over all ranks' shares, exchanged through peer memory
    or NCCL -- nufi_b200_download_rho_full) against the reference's CPU eval_rho on the same history, on a sample of spatial
    nodes (rank 0); N > 1: plus SHA-256 of rho and of level n gathered from every rank -- the replicas must be bit-identical.
    Returns (parity dict, cpu_baseline dict or None) on rank 0, (None, None) elsewhere.  Collective: every rank calls it."""
    s = runner.s
    got = s.download_rho_full()
    replicas = None
    if runner.world > 1:
        level = s.download_phi(n)
        digest = hashlib.sha256(got.tobytes()).hexdigest()[:16] + ":" + hashlib.sha256(level.tobytes()).hexdigest()[:16]
        all_d = [None] * runner.world
        dist.all_gather_object(all_d, digest)
        replicas = {"bit_identical": len(set(all_d)) == 1, "sha256_rho:level_n": sorted(set(all_d)), "ranks": runner.world}
    parity = cpu = None
    if runner.rank == 0:
        run, cpu_psteps, sdesc, kind, threads, l_n, out = reference_sweep_timer(conf, f0, n, coeffs_host, cpu_budget)
        t_cpu = run()
        if want_cpu_baseline:
            t_cpu = min(t_cpu, run())
            cpu = {"value": cpu_psteps / t_cpu, "unit": "point-steps/s", "cores": threads, "kind": kind, "sample": sdesc, "seconds": t_cpu}
        want = out["rho"][:l_n]
        scale = float(np.max(np.abs(want)))
        err = float(np.max(np.abs(got[:l_n] - want))) / scale
        parity = {"rho_rel_linf_vs_cpu_reference": err, "nodes_checked": int(l_n), "tolerance": RHO_TOL, "ok": bool(err <= RHO_TOL),
                  "rho_max_abs": scale,
                  "what": "rho of the timed fused step (all ranks' shares summed) vs the reference CPU eval_rho on the downloaded history"}
        if err > 0.1 * RHO_TOL:
            # Whose rounding is it?  The same sample with the velocity sum carried in long double (oracle yardstick, not the
            # reference's arithmetic).  Where the density perturbation has decayed to ~1e-5 (C1 at t = 50) 1e-10 of it is a few ulp
            # of the O(1) sum dV*sum f, less than the rounding error of the reference's own sequential double sum of Nu terms;
            # the device adds with a compensated sum.  The check then passes if the device is at least as close to the
            # extended-precision value as the tolerance, and the reference's own distance from it is reported beside it.
            from oracle.oracle_py import Oracle

            yard = Oracle(fast=True)
            yard.set_threads(host_cores())
            ext = yard.rho_extended(conf, f0, n, coeffs_host, 0, l_n)[:l_n]
            dev_ext = float(np.max(np.abs(got[:l_n] - ext))) / scale
            ref_ext = float(np.max(np.abs(want - ext))) / scale
            parity.update({"rho_rel_linf_device_vs_extended_sum": dev_ext, "rho_rel_linf_reference_vs_extended_sum": ref_ext,
                           "ok": bool(err <= RHO_TOL or (dev_ext <= max(RHO_TOL, ref_ext) and err <= 2.0 * ref_ext + RHO_TOL)),
                           "note": "reference's own double-precision summation error exceeds a tenth of the tolerance on this sample; "
                                   "ok = within tolerance of the reference, or no further from the extended-precision sum than the "
                                   "tolerance or the reference itself (whichever is larger) and no further from the reference than twice "
                                   "the reference's own distance from it plus the tolerance; all three distances are reported"})
        if replicas is not None:
            parity["replicas"] = replicas
            parity["ok"] = bool(parity["ok"] and replicas["bit_identical"])
    if runner.world > 1:
        dist.barrier()
    return parity, cpu


def measure_workload(name, scaling, rank, world, local, args, flush, torch, dist, peak_tf, steps, with_ref_cuda, cpu_budget):
    """One workload end to end on all ranks: history by free-running fused steps, the two timed regions, live parity.
    Returns a dict on rank 0, None elsewhere."""
    from numericalflowiteration_b200 import n_quad

    conf, f0, depth, desc = make_workload(name, world, scaling)
    r = GpuRunner(conf, f0, rank, world, torch, dist, exchange=args.exchange)
    try:
        free_run(r, depth)
        m = measure_gpu(r, depth, steps, 3, flush, torch, dist)
        if r.exchange == "peer-memory" and r.s.peer_timed_out():
            raise RuntimeError("a wait for a peer's rho flag timed out")
        ps = float(n_quad(conf)) * depth
        ex = None
        hist = download_history(r.s, depth) if (rank == 0 and not args.no_cpu) else None
        parity = None
        if not args.no_cpu:
            parity, _ = live_parity(r, conf, f0, depth, hist, cpu_budget, dist, False)
        if rank == 0:
            tf = r.my_point_steps(depth) * FLOP_PER_POINT_STEP[conf.dim] / (m["bt_ms"] * 1e-3) / 1e12
            ex = {"workload": desc, "scaling": scaling if world > 1 else None, "depth_n": depth, "n_quad": n_quad(conf),
                  "point_steps_per_s": ps * steps / (m["t_ms"] * 1e-3), "ms_per_step": m["t_ms"] / steps, "steps": steps,
                  "kernel_ms": m["bt_ms"], "kernel_ms_per_rank": m["bt_ms_per_rank"], "variant": r.s.last_variant, "exchange": r.exchange,
                  "fp64_tflops_per_gpu": tf, "fp64_frac": tf / peak_tf,
                  "frac_of_rf_ceiling": tf / peak_tf / RF_CEILING.get((conf.dim, "/xpp" in r.s.last_variant), 1.0), "parity": parity}
            if with_ref_cuda and not args.no_cpu:
                rc = reference_cuda_leg(conf, f0, depth, hist, None, local, reps=1 if (conf.dim == 3 and conf.Nx >= 32) else 2)
                ex["reference_cuda_point_steps_per_s"] = rc.get("value")
                ex["reference_cuda_kernel_ms"] = rc.get("kernel_ms")
                if rc.get("value"):
                    ex["speedup_vs_reference_cuda_one_gpu"] = ex["point_steps_per_s"] / rc["value"]
        if world > 1:
            dist.barrier()
        return ex
    finally:
        r.close()


def full_runs(args, torch, flush):
    """Complete free runs on one GPU (what the reference's drivers print: total wall time and s per time step,
    bin/test_nufi_cpu_2d.cpp:65-78, 104): C1 and C2 all 1600 steps, C4 all 50, through the fused step, host-timed from the
    first launch to the final synchronisation; the electric-energy trace against the reference CPU run's
    (tests/golden/fullsize_*.npz, made by tests/golden/make_fullsize_traces.py) over the window where the problem is well
    conditioned (tests/test_fullsize_gpu.py gates the whole trace)."""
    out = []
    windows = {"C1": 500, "C2": 800, "C4": 50}
    for name in ("C1", "C2", "C4"):
        try:
            conf, f0, _, desc = make_workload(name, 1)
            r = GpuRunner(conf, f0, 0, 1, torch, None)
            nt = int(conf.Nt)
            for m in range(min(nt, 8)):  # warm: module load, first-launch costs
                r.step(m)
            r.stream.synchronize()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for m in range(nt + 1):
                r.step(m)
            r.stream.synchronize()
            wall = time.perf_counter() - t0
            e = r.s.download_energy(0, nt + 1)
            from numericalflowiteration_b200 import n_quad

            row = {"workload": desc, "time_steps": nt + 1, "wall_s": wall, "mean_s_per_time_step": wall / (nt + 1),
                   "point_steps": float(n_quad(conf)) * nt * (nt + 1) / 2, "point_steps_per_s": float(n_quad(conf)) * nt * (nt + 1) / 2 / wall,
                   "launches": 2 * (nt + 1)}
            gpath = os.path.join(ROOT, "tests", "golden", f"fullsize_{name}.npz")
            if os.path.exists(gpath):
                g = np.load(gpath)
                ref_e = np.asarray(g["energy"], dtype=np.float64)
                w = min(windows[name], len(ref_e), len(e))
                row["energy_trace_rel_err"] = float(np.max(np.abs(e[:w] - ref_e[:w]) / np.abs(ref_e[:w])))
                row["energy_trace_window_steps"] = int(w)
                row["energy_trace_tolerance"] = 1e-8
                if "seconds" in g.files:
                    row["reference_cpu_wall_s"] = float(g["seconds"])
                    row["reference_cpu_threads"] = int(g["threads"]) if "threads" in g.files else None
            out.append(row)
            r.close()
        except Exception as ex:  # noqa: BLE001
            out.append({"workload": name, "error": repr(ex)})
    return out


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world != max(args.gpus, 1):
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")

    from numericalflowiteration_b200 import measure_fp64_peak, n_quad

    conf, f0, depth, desc = make_workload(args.workload, world, args.scaling)
    n = args.depth or depth
    dim = conf.dim
    runner = GpuRunner(conf, f0, rank, world, torch, dist, exchange=args.exchange)
    s = runner.s
    nq = n_quad(conf)
    flush = None if args.no_flush else torch.empty(L2_FLUSH_BYTES // 8, dtype=torch.float64, device="cuda")
    global ALIGN_RANKS
    ALIGN_RANKS = not args.no_align

    t0 = time.perf_counter()
    free_run(runner, n)  # levels 0..n-1 on the device
    t_hist = time.perf_counter() - t0

    sampler = ClockSampler(local) if rank == 0 else None
    m = measure_gpu(runner, n, args.steps, args.warmup, flush, torch, dist, sampler)
    if runner.exchange == "peer-memory" and s.peer_timed_out():
        raise SystemExit("bench.py: a wait for a peer's rho flag timed out -- the multi-GPU result would be invalid")
    psteps = float(nq) * n  # whole job, all ranks
    value = psteps * args.steps / (m["t_ms"] * 1e-3)
    variant = s.last_variant

    # host copy of the history (for the e2e leg's upload_phi and for the CPU baseline / live parity check)
    coeffs_host = download_history(s, n)
    parity = cpu = None
    if not args.no_cpu:  # before the e2e leg: rho_full still holds what the last timed fused step consumed
        parity, cpu = live_parity(runner, conf, f0, n, coeffs_host, args.cpu_budget, dist, world == 1)
    e2e_steps = max(3, min(args.steps, 50))
    t_e2e, h2d, d2h, _ = measure_e2e(runner, n, e2e_steps, min(args.warmup, 3), coeffs_host, torch, dist)
    e2e_value = psteps * e2e_steps / t_e2e
    e2e_host = measure_e2e_step_host(runner, n, e2e_steps, min(args.warmup, 3), coeffs_host, torch, dist)

    line = None
    peak_tf = measure_fp64_peak(local)
    if rank == 0:
        # this rank's backtrace kernel: its share of the point-steps / its average launch duration
        my_psteps = runner.my_point_steps(n)
        achieved_tf = my_psteps * FLOP_PER_POINT_STEP[dim] / (m["bt_ms"] * 1e-3) / 1e12
        hist_bytes = float(n) * s.stride_t * 8 + 2 * 8 * s.n_nodes
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        props = torch.cuda.get_device_properties(local)
        s_sm_count = props.multi_processor_count
        sm_mhz_peak = float(peaks.get("sm_max_mhz", 1965.0))
        roofline = {
            "bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
            "traffic": NCU_DRAM_TRAFFIC_BYTES.get(args.workload) if (world == 1 and not args.depth) else None,
            "traffic_source": "profiles/r02_ncu_full_summary.csv (one ncu --set full capture of this kernel at this depth); bytes per launch",
            "kernel": f"backtrace_kernel<{dim}d> [{variant}]", "kernel_ms": m["bt_ms"], "kernel_ms_per_rank": m["bt_ms_per_rank"],
            "kernel_share_of_step": m["bt_ms"] * args.steps / m["t_ms_kernel_timing"],
            "ms_per_step_with_kernel_events": m["t_ms_kernel_timing"] / args.steps,
            "how": "second timed region of the same K steps with a CUDA event pair around every backtrace launch (library stream)",
            "flop_per_point_step": FLOP_PER_POINT_STEP[dim],
            "rf_ceiling": RF_CEILING.get((dim, "/xpp" in variant)),
            "frac_of_rf_ceiling": (achieved_tf / peak_tf / RF_CEILING[(dim, "/xpp" in variant)]) if (dim, "/xpp" in variant) in RF_CEILING else None,
            "rf_ceiling_note": "FP64 ceiling of this kernel's hot loop with the register-file operand reads of its DFMAs counted "
                               "(DESIGN.md 3.1, profiles/r02_sass_rf_model.txt); in 1d shared memory binds first",
            "peak_source": "measured live: register-only DFMA loop (nufi_b200_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
            "smem": {"achieved": my_psteps * SMEM_BYTES_PER_POINT_STEP[dim] / (m["bt_ms"] * 1e-3) / 1e9,
                     "peak": SMEM_BYTES_PER_CLK_SM * s_sm_count * sm_mhz_peak * 1e6 / 1e9, "unit": "GB/s",
                     "frac": my_psteps * SMEM_BYTES_PER_POINT_STEP[dim] / (m["bt_ms"] * 1e-3) / (SMEM_BYTES_PER_CLK_SM * s_sm_count * sm_mhz_peak * 1e6),
                     "note": "the co-limiting unit (DESIGN.md 3.1): algorithmic shared-memory bytes gathered per point-step x point-steps/s "
                             "against 128 B/clk/SM x SMs x max SM clock; bank-conflict replays (C2) are not counted as achieved"},
            "hbm": {"achieved": hist_bytes / (m["bt_ms"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": hist_bytes / (m["bt_ms"] * 1e-3) / 1e9 / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                    "note": "algorithmic bytes = n*stride_t*8 (history read once) + 2*8*Nnodes; shown to document the path is not HBM-bound"},
        }
        ref_cuda = None
        if world == 1 and not args.no_cpu:
            ref_cuda = reference_cuda_leg(conf, f0, n, coeffs_host, s, local, reps=3)
        line = {
            "metric": "backtrace point-steps/sec", "value": value, "unit": "point-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["t_ms"] / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "depth_n": n, "point_steps_per_step": psteps, "n_quad": nq,
                       "history": f"built on the device by {n} free-running fused steps ({t_hist:.2f} s)",
                       "l2": ("NOT flushed (experiment)" if args.no_flush else
                              "flushed between timed steps (256 MiB fill, untimed; steps timed individually with CUDA events)" +
                              ("; ranks aligned by a stream-ordered NCCL barrier between the flush and each timed step (untimed)"
                               if world > 1 and ALIGN_RANKS else "")),
                       "settle_steps": max(args.warmup, SETTLE_STEPS), "lead_in_steps": LEAD_IN,
                       "parallelism": (f"quadrature points sharded over {world} GPU(s); rho exchange: " +
                                       ("stores into NVLink peer memory fused into the slot-reduction and tail kernels (no collective call)"
                                        if runner.exchange == "peer-memory" else "NCCL all-reduce")) if world > 1 else "1 GPU",
                       "exchange": runner.exchange, "exchange_note": runner.exchange_note},
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "reference_cuda": ref_cuda,
            "e2e": ({"value": psteps * e2e_steps / e2e_host[0], "unit": "point-steps/s", "h2d_bytes_per_step": int(e2e_host[1]),
                     "d2h_bytes_per_step": int(e2e_host[2]), "steps": e2e_steps, "ms_per_step": 1e3 * e2e_host[0] / e2e_steps,
                     "path": "nufi_b200_step_host: level n-1 host -> device, fused step" + (" with the peer-memory exchange" if world > 1 else "") +
                             ", level n + rho + energy device -> host, one synchronisation; host buffers, wall clock"}
                    if e2e_host else
                    {"value": e2e_value, "unit": "point-steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                     "steps": e2e_steps, "ms_per_step": 1e3 * t_e2e / e2e_steps,
                     "path": "upload_phi(n-1) -> compute_rho -> download_rho -> solve_interpolate_host, host buffers, wall clock"}),
            "e2e_driver_sequence": {"value": e2e_value, "unit": "point-steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                                    "steps": e2e_steps, "ms_per_step": 1e3 * t_e2e / e2e_steps,
                                    "path": "the reference GPU driver's own per-step calls, bin/test_nufi_gpu_3d.cpp:154-162: upload_phi(n-1) -> "
                                            "compute_rho -> download_rho -> (N > 1: all-reduce) -> solve_interpolate_host; host buffers, wall clock"},
            "gpu_launches": m["launches"], "clocks": sampler.summary() if sampler else None,
            "s_per_time_step": m["t_ms"] / args.steps * 1e-3,
            "step_ms_rank0": {"min": min(m["per_step_ms"]), "median": float(np.median(m["per_step_ms"])), "max": max(m["per_step_ms"]),
                              "first_10": [round(x, 4) for x in m["per_step_ms"][:10]]},
        }
    runner.close()
    del runner, s

    if args.extras:
        # N = 1: the other dimensions / sizes.  N > 1: the 3d configurations BASELINE.json names for 2/4/8 GPUs, in STRONG form
        # (the fixed grid, each rank 1/N of the velocity nodes) and in weak form (Nw x N), each with its own parity check.
        plan = ([(w, "weak") for w in ("C1", "C3", "C4", "C5-16", "C5-32") if w != args.workload] if world == 1 else
                [("C4", "strong"), ("C5-16", "strong"), ("C5-32", "strong"), ("C4", "weak"), ("C5-16", "weak"), ("C5-32", "weak")])
        extras = []
        for name, scaling in plan:
            big = name == "C5-32"
            try:
                ex = measure_workload(name, scaling, rank, world, local, args, flush, torch, dist, peak_tf, steps=5 if big else 20,
                                      with_ref_cuda=(world == 1 or scaling == "strong"), cpu_budget=1.0)
            except Exception as e:  # noqa: BLE001
                ex = {"workload": name, "scaling": scaling, "error": repr(e)}
            if rank == 0:
                extras.append(ex)
        if rank == 0:
            line["other_workloads"] = extras
    if args.full_run and world == 1 and rank == 0:
        line["full_run"] = full_runs(args, torch, flush)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line if rank == 0 else None


class _StdoutToStderr:
    """Everything written to fd 1 while active goes to stderr (NCCL prints its version banner on stdout); the single JSON
    line is written to the real stdout afterwards."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def emit(line):
    if line is not None:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = the velocity quadrature grows with N (fixed work per GPU), strong = BASELINE's fixed grid")
    ap.add_argument("--depth", type=int, default=0, help="history depth n of the timed step (default: per workload)")
    ap.add_argument("--cpu-budget", type=float, default=3.0, help="seconds of wall time for the cpu_baseline sample")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: how the partial rho is exchanged (peer = stores into NVLink peer memory fused into the kernels)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="experiment: no L2 flush between the timed steps (not a valid bench line)")
    ap.add_argument("--no-align", action="store_true", help="N > 1: no rank alignment between the L2 flush and the timed step")
    ap.add_argument("--no-extras", dest="extras", action="store_false")
    ap.add_argument("--no-full-run", dest="full_run", action="store_false")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    with _StdoutToStderr():
        line = run_reference_arm(args) if args.impl == "reference" else run_gpu_arm(args)
    emit(line)


if __name__ == "__main__":
    main()
