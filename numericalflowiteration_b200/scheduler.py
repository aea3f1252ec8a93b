"""Host-side mirror of ``nufi::dim{1,2,3}::cuda_scheduler<double,4>`` (reference nufi/cuda_scheduler.hpp:33-164,
173-279, 286-392) over the C ABI of libnufi_b200.so.

Same five methods with the same argument meaning and error behaviour (``compute_rho``, ``download_rho``,
``upload_phi``, ``compute_metrics``, ``download_metrics``), plus the CPU-driver-shaped calls
(``eval_rho`` sweep, ``solve``, ``interpolate``) and the fused ``step`` that removes the host round trip of
bin/test_nufi_gpu_3d.cpp:154-162.  One scheduler drives one GPU; several GPUs = several processes, each with
its own q-range (see :mod:`numericalflowiteration_b200.distributed`).

Exceptions: :class:`CudaError` ~ ``nufi::cuda::exception`` (nufi/cuda_runtime.hpp:93-98),
:class:`RangeError` ~ ``std::range_error`` (nufi/cuda_kernel.cu:115-116), ``MemoryError`` ~ ``std::bad_alloc``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .config import F0, n_nodes, n_quad, stride_t

__all__ = ["CudaScheduler", "CudaGroup", "CudaError", "RangeError", "measure_fp64_peak", "device_count"]


class CudaError(RuntimeError):
    """CUDA/cuFFT failure or no device (the reference throws nufi::cuda::exception)."""


class RangeError(ValueError):
    """Time step / index out of range (the reference throws std::range_error)."""


def _raise(code: int, msg: str):
    if code == _lib.ERR_RANGE:
        raise RangeError(msg)
    if code == _lib.ERR_ALLOC:
        raise MemoryError(msg)
    if code == _lib.ERR_ARG:
        raise ValueError(msg)
    raise CudaError(msg)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class CudaScheduler:
    """``cuda_scheduler<double,4>`` for one GPU.

    Parameters mirror the reference constructor (``cuda_scheduler(const config_t&)``); ``f0`` replaces the
    compile-time choice of ``config_t::f0``; ``device=-1`` uses the current CUDA device.
    """

    def __init__(self, conf, f0: F0 | None = None, order: int = 4, device: int = -1):
        self._L = _lib.load()
        self._h = C.c_void_p()
        self.conf = conf
        self.dim = conf.dim
        self.order = order
        self.f0 = f0 if f0 is not None else F0.default(conf.dim)
        self.n_nodes = n_nodes(conf)
        self.n_quad = n_quad(conf)
        self.stride_t = stride_t(conf, order)
        create = getattr(self._L, f"nufi_b200_create_{conf.dim}d")
        rc = create(C.addressof(conf), order, C.addressof(self.f0), device, C.byref(self._h))
        if rc != _lib.OK:
            self._h = C.c_void_p()
            _raise(rc, self._L.nufi_b200_last_error(None).decode())

    # -- plumbing
    def _ck(self, rc: int):
        if rc != _lib.OK:
            _raise(rc, self._L.nufi_b200_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.nufi_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- the reference scheduler's methods
    def compute_rho(self, n: int, q_begin: int, q_end: int) -> None:
        """Asynchronous partial rho over flat q in [q_begin, q_end) (cuda_scheduler.hpp:88-111)."""
        self._ck(self._L.nufi_b200_compute_rho(self._h, n, q_begin, q_end))

    def download_rho(self, rho: np.ndarray) -> None:
        """Blocking; ``rho += partial`` (caller zeroes first, bin/test_nufi_gpu_3d.cpp:154-156)."""
        assert rho.dtype == np.float64 and rho.flags.c_contiguous and rho.size == self.n_nodes
        self._ck(self._L.nufi_b200_download_rho(self._h, _ptr(rho)))

    def upload_phi(self, n: int, coeffs: np.ndarray) -> None:
        """Copies level ``n`` out of the BASE array of the host history (cuda_kernel.cu:147-156)."""
        assert coeffs.dtype == np.float64 and coeffs.flags.c_contiguous
        if coeffs.size < (n + 1) * self.stride_t:
            raise RangeError("host history holds fewer than n+1 levels")
        self._ck(self._L.nufi_b200_upload_phi(self._h, n, _ptr(coeffs)))

    def compute_metrics(self, n: int, q_begin: int, q_end: int) -> None:
        self._ck(self._L.nufi_b200_compute_metrics(self._h, n, q_begin, q_end))

    def set_metrics_grid(self, conf_metrics=None) -> None:
        """dim 1: integrate the metrics over the (x,u) grid of ``conf_metrics`` while the field stays on the grid of the
        scheduler's own configuration -- the reference's ``cuda_scheduler(conf, conf_metrics)`` (cuda_scheduler.hpp:65-85,
        cuda_kernel.cu:55-70).  ``None`` restores the scheduler's own grid."""
        self._ck(self._L.nufi_b200_set_metrics_grid_1d(self._h, C.addressof(conf_metrics) if conf_metrics is not None else None))

    def download_metrics(self, metrics: np.ndarray) -> None:
        assert metrics.dtype == np.float64 and metrics.size == 4
        self._ck(self._L.nufi_b200_download_metrics(self._h, _ptr(metrics)))

    # -- CPU-driver-shaped calls
    def eval_rho(self, n: int, fetch: bool = True):
        """The drivers' OpenMP sweep ``rho[l] = eval_rho(n, l, coeffs, conf)`` for all l (CPU convention)."""
        rho = np.empty(self.n_nodes) if fetch else None
        self._ck(self._L.nufi_b200_eval_rho_all(self._h, n, _ptr(rho) if fetch else None))
        return rho

    def solve_interpolate(self, n: int, rho: np.ndarray | None = None, want_energy: bool = True):
        """``poisson.solve`` + ``interpolate`` into device level n; returns the electric energy."""
        e = C.c_double(0.0)
        if rho is not None:
            rho = np.ascontiguousarray(rho, dtype=np.float64)
            self._ck(self._L.nufi_b200_solve_interpolate_host(self._h, n, _ptr(rho), C.byref(e)))
            return e.value
        self._ck(self._L.nufi_b200_solve_interpolate(self._h, n, C.byref(e) if want_energy else None))
        return e.value if want_energy else None

    # -- fused path
    def step(self, n: int) -> None:
        """Backtrace + reduce + Poisson + interpolate + store level n, asynchronously, no host round trip."""
        self._ck(self._L.nufi_b200_step(self._h, n))

    def step_host(self, n: int, coeffs: np.ndarray, rho: np.ndarray | None = None, peer: bool = False) -> float:
        """One time step for a history kept on the HOST (the reference GPU drivers' loop body in one call): level n-1 of `coeffs`
        goes to the device, the fused step runs, level n is written into `coeffs`, rho (optional) is filled; returns the energy."""
        assert coeffs.dtype == np.float64 and coeffs.flags.c_contiguous
        if n > self.conf.Nt or coeffs.size < (n + 1) * self.stride_t:
            raise RangeError("Time-step out of range.")
        e = C.c_double(0.0)
        self._ck(self._L.nufi_b200_step_host(self._h, n, _ptr(coeffs), _ptr(rho) if rho is not None else None, C.byref(e), 1 if peer else 0))
        return e.value

    def download_rho_full(self) -> np.ndarray:
        """rho of the most recent (peer/group) step as the field tail consumed it: CPU convention, summed over all ranks."""
        rho = np.empty(self.n_nodes)
        self._ck(self._L.nufi_b200_download_rho_full(self._h, _ptr(rho)))
        return rho

    def download_energy(self, n_begin: int, n_end: int) -> np.ndarray:
        out = np.zeros(max(n_end - n_begin, 0))
        self._ck(self._L.nufi_b200_download_energy(self._h, n_begin, n_end, _ptr(out)))
        return out

    def download_phi(self, n: int) -> np.ndarray:
        out = np.empty(self.stride_t)
        self._ck(self._L.nufi_b200_download_phi(self._h, n, _ptr(out)))
        return out

    def upload_history(self, coeffs: np.ndarray, n_levels: int) -> None:
        """Levels 0..n_levels-1 from a host history (restart from a checkpoint)."""
        coeffs = np.ascontiguousarray(coeffs, dtype=np.float64).ravel()
        if coeffs.size < n_levels * self.stride_t:
            raise RangeError("host history holds fewer than n_levels levels")
        self._ck(self._L.nufi_b200_upload_history(self._h, n_levels, _ptr(coeffs)))

    def download_history(self, n_levels: int) -> np.ndarray:
        """Levels 0..n_levels-1 in the reference layout: the complete simulation state (checkpoint)."""
        out = np.empty(n_levels * self.stride_t)
        self._ck(self._L.nufi_b200_download_history(self._h, n_levels, _ptr(out)))
        return out

    def eval_f(self, n: int, points: np.ndarray, full: bool = True) -> np.ndarray:
        """f(t_n, x, v) (full=True: eval_f) or ftilda (full=False) at phase-space points [npts, 2*dim]."""
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2 * self.dim)
        out = np.empty(len(pts))
        self._ck(self._L.nufi_b200_eval_f(self._h, n, len(pts), _ptr(pts), _ptr(out), int(full)))
        return out

    def eval_phase_flow(self, n: int, points: np.ndarray) -> np.ndarray:
        """Feet (x.., v..) at t = 0 of the characteristics through phase-space points [npts, 2*dim] at t_n
        (``eval_phase_flow``, nufi/rho.hpp:98-131; like the reference, nothing is traced for n <= 1)."""
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2 * self.dim)
        out = np.empty_like(pts)
        self._ck(self._L.nufi_b200_eval_phase_flow(self._h, n, len(pts), _ptr(pts), _ptr(out)))
        return out

    def eval_field(self, n: int, points: np.ndarray, derivative_axis: int = -1) -> np.ndarray:
        """phi_n (derivative_axis=-1) or its first derivative along an axis at positions [npts, dim]."""
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, self.dim)
        out = np.empty(len(pts))
        self._ck(self._L.nufi_b200_eval_field(self._h, n, derivative_axis, len(pts), _ptr(pts), _ptr(out)))
        return out

    def sync(self) -> None:
        self._ck(self._L.nufi_b200_sync(self._h))

    # -- device plumbing (torch.distributed host layer)
    def set_stream(self, cuda_stream: int | None) -> None:
        self._ck(self._L.nufi_b200_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def rho_device_ptr(self) -> int:
        p = C.c_void_p()
        self._ck(self._L.nufi_b200_rho_device(self._h, C.byref(p)))
        return p.value

    def field_tail_device(self, n: int, d_rho_partial_sum: int) -> None:
        self._ck(self._L.nufi_b200_field_tail_device(self._h, n, C.c_void_p(d_rho_partial_sum)))

    # -- multi-GPU step with the rho exchange fused into the kernels over NVLink peer memory (csrc/peer.cu)
    PEER_HANDLE_BYTES = 64

    def peer_export(self, world: int) -> bytes:
        """Allocates this rank's exchange buffer; returns its CUDA IPC handle (all-gather these in rank order)."""
        buf = C.create_string_buffer(self.PEER_HANDLE_BYTES)
        self._ck(self._L.nufi_b200_peer_export(self._h, world, buf))
        return buf.raw

    def peer_attach(self, rank: int, world: int, handles: bytes) -> None:
        """Maps the exchange buffers of all ranks (``handles`` = world x 64 bytes in rank order)."""
        assert len(handles) == world * self.PEER_HANDLE_BYTES
        self._ck(self._L.nufi_b200_peer_attach(self._h, rank, world, C.c_char_p(handles)))

    def peer_step(self, n: int) -> None:
        """Backtrace of this rank's q-share -> push into every GPU -> tail that waits for all ranks; asynchronous."""
        self._ck(self._L.nufi_b200_peer_step(self._h, n))

    def peer_timed_out(self) -> bool:
        t = C.c_int(0)
        self._ck(self._L.nufi_b200_peer_status(self._h, C.byref(t)))
        return bool(t.value)

    def peer_detach(self) -> None:
        self._ck(self._L.nufi_b200_peer_detach(self._h))

    # -- introspection
    @property
    def launches(self) -> int:
        return int(self._L.nufi_b200_launch_count(self._h))

    def set_kernel_timing(self, on: bool) -> None:
        """Bracket every backtrace launch with a CUDA event pair (needed by last_backtrace_ms / backtrace_time; off by
        default because the events keep the field tail from launching programmatically behind the kernel)."""
        self._ck(self._L.nufi_b200_set_kernel_timing(self._h, int(on)))

    def last_backtrace_ms(self) -> float:
        ms = C.c_float(0)
        self._ck(self._L.nufi_b200_last_backtrace_ms(self._h, C.byref(ms)))
        return ms.value

    def backtrace_time(self, reset: bool = False) -> tuple[float, int]:
        """(total GPU ms, launches) of the backtrace kernel since the last reset (blocking)."""
        ms, cnt = C.c_double(0), C.c_uint64(0)
        self._ck(self._L.nufi_b200_backtrace_time(self._h, C.byref(ms), C.byref(cnt), int(reset)))
        return ms.value, int(cnt.value)

    @property
    def last_variant(self) -> str:
        return self._L.nufi_b200_last_variant(self._h).decode()

    def set_variant(self, v: int) -> None:
        self._ck(self._L.nufi_b200_set_variant(self._h, v))

    def set_tile_nodes(self, nodes_per_tile: int) -> None:
        """Lane layout of the backtrace kernel: 32 = one node per lane (default), 16..1 = 32/TN neighbouring velocities per node."""
        self._ck(self._L.nufi_b200_set_tile_nodes(self._h, nodes_per_tile))

    def set_tail_variant(self, v: int) -> None:
        """0 auto, 1 cuFFT tail, 2 fused single-CTA tail."""
        self._ck(self._L.nufi_b200_set_tail_variant(self._h, v))

    @property
    def last_tail_variant(self) -> str:
        return self._L.nufi_b200_last_tail_variant(self._h).decode()


def measure_fp64_peak(device: int = -1) -> float:
    """Register-only DFMA loop: measured FP64 peak of the device in TFLOP/s."""
    L = _lib.load()
    t = C.c_double(0)
    rc = L.nufi_b200_measure_fp64_peak(device, C.byref(t))
    if rc != _lib.OK:
        _raise(rc, L.nufi_b200_last_error(None).decode())
    return t.value


def device_count() -> int:
    """Visible CUDA devices (cuda::device_count in the reference); raises CudaError without a driver."""
    L = _lib.load()
    n = C.c_int(0)
    rc = L.nufi_b200_device_count(C.byref(n))
    if rc != _lib.OK:
        _raise(rc, L.nufi_b200_last_error(None).decode())
    return n.value


class CudaGroup:
    """Several GPUs driven by ONE process -- the shape of the reference's ``cuda_scheduler`` (one host thread, all visible
    devices, nufi/cuda_scheduler.hpp:43-63) with the host fan-in replaced by an exchange over NVLink on the devices.
    ``step(n)``: every device traces its contiguous share of the flat q range, the partial rho vectors are exchanged
    (stores into peer memory fused into the slot reduction, or an NCCL all-reduce), every device runs the (tiny,
    deterministic) field tail."""

    def __init__(self, conf, f0: F0 | None = None, devices=None, order: int = 4):
        self._L = _lib.load()
        devices = list(range(device_count())) if devices is None else list(devices)
        self.scheds = [CudaScheduler(conf, f0, order=order, device=d) for d in devices]
        arr = (C.c_void_p * len(self.scheds))(*[s._h for s in self.scheds])
        self._g = C.c_void_p()
        rc = self._L.nufi_b200_group_create(arr, len(self.scheds), C.byref(self._g))
        if rc != _lib.OK:
            self._g = C.c_void_p()
            _raise(rc, self._L.nufi_b200_group_last_error(None).decode())

    def step(self, n: int) -> None:
        rc = self._L.nufi_b200_group_step(self._g, n)
        if rc != _lib.OK:
            _raise(rc, self._L.nufi_b200_group_last_error(self._g).decode())

    @property
    def exchange(self) -> str:
        """'peer-memory' (fused into the kernels, default when the devices can map each other), 'nccl' or 'single'."""
        return self._L.nufi_b200_group_exchange(self._g).decode()

    def set_exchange(self, mode: str) -> None:
        rc = self._L.nufi_b200_group_set_exchange(self._g, {"peer": 0, "peer-memory": 0, "nccl": 1}[mode])
        if rc != _lib.OK:
            _raise(rc, self._L.nufi_b200_group_last_error(self._g).decode())

    def sync(self) -> None:
        rc = self._L.nufi_b200_group_sync(self._g)
        if rc != _lib.OK:
            _raise(rc, self._L.nufi_b200_group_last_error(self._g).decode())

    def close(self) -> None:
        if getattr(self, "_g", None) and self._g.value:
            self._L.nufi_b200_group_destroy(self._g)
            self._g = C.c_void_p()
        for s in getattr(self, "scheds", []):
            s.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
