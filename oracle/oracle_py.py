"""ctypes front end of the parity oracle.  TEST INFRASTRUCTURE ONLY.

Loads ``oracle/build/liboracle*.so`` (the C restatement, nufi_oracle.c) and, when present,
``oracle/_ref/libnufi_ref*.so`` (the real reference headers compiled in place, ref_harness.cpp).
Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_sz, _d, _i, _p = C.c_size_t, C.c_double, C.c_int, C.c_void_p
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


class OrcF0(C.Structure):
    _fields_ = [("kind", C.c_int), ("p", C.c_double * 4)]


def _f0(f) -> OrcF0:
    o = OrcF0()
    o.kind = int(f.kind)
    for i in range(4):
        o.p[i] = float(f.p[i])
    return o


def build(ref: bool = True) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference exists) with oracle/Makefile."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"] + (["ref"] if ref else []), check=True)


def _stride(conf, order=4):
    s = conf.Nx + order - 1
    if conf.dim >= 2:
        s *= conf.Ny + order - 1
    if conf.dim >= 3:
        s *= conf.Nz + order - 1
    return s


def _nodes(conf):
    n = conf.Nx
    if conf.dim >= 2:
        n *= conf.Ny
    if conf.dim >= 3:
        n *= conf.Nz
    return n


class Oracle:
    """The C restatement.  ``fast=True`` loads the -O3/AVX2+FMA build (CPU speed baseline only)."""

    kind = "port"

    def __init__(self, fast: bool = False):
        path = os.path.join(HERE, "build", "liboracle_fast.so" if fast else "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = L = C.CDLL(path)
        self.path = path
        L.orc_num_threads.restype = _i
        L.orc_bspline_basis.argtypes = [_i, _i, _d, _dp]
        L.orc_deboor.argtypes = [_i, _i, _d, _dp, _sz]
        L.orc_deboor.restype = _d
        for dim, nco in ((1, 1), (2, 2), (3, 3)):
            getattr(L, f"orc_field_{dim}d").argtypes = [_i] + [_i] * nco + [_d] * nco + [_dp, _p]
            getattr(L, f"orc_field_{dim}d").restype = _d
            getattr(L, f"orc_f0_{dim}d").argtypes = [_p] + [_d] * (2 * nco)
            getattr(L, f"orc_f0_{dim}d").restype = _d
            for nm in ("ftilda", "f"):
                fn = getattr(L, f"orc_{nm}_{dim}d")
                fn.argtypes = [_i, _sz] + [_d] * (2 * nco) + [_dp, _p, _p]
                fn.restype = _d
            getattr(L, f"orc_rho_{dim}d").argtypes = [_i, _sz, _sz, _dp, _p, _p]
            getattr(L, f"orc_rho_{dim}d").restype = _d
            getattr(L, f"orc_rho_sweep_{dim}d").argtypes = [_i, _sz, _dp, _p, _p, _sz, _sz, _dp]
            getattr(L, f"orc_rho_partial_{dim}d").argtypes = [_i, _sz, _dp, _p, _p, _sz, _sz, _dp]
            getattr(L, f"orc_metrics_{dim}d").argtypes = [_i, _sz, _dp, _p, _p, _sz, _sz, _dp]
            getattr(L, f"orc_poisson_{dim}d").argtypes = [_p, _dp]
            getattr(L, f"orc_poisson_{dim}d").restype = _d
            getattr(L, f"orc_interpolate_{dim}d").argtypes = [_i, _dp, _dp, _p]
            getattr(L, f"orc_run_{dim}d").argtypes = [_i, _p, _p, _sz, _sz, _dp, _p, _p]

    # -- helpers
    def threads(self) -> int:
        return int(self.lib.orc_num_threads())

    def set_threads(self, n: int) -> int:
        """OpenMP threads of the sweeps (torchrun exports OMP_NUM_THREADS=1: the CPU-baseline legs set the count explicitly)."""
        self.lib.orc_set_num_threads(int(n))
        return self.threads()

    def basis(self, order, der, x):
        out = np.zeros(order)
        self.lib.orc_bspline_basis(order, der, float(x), out)
        return out

    def field(self, conf, level, pos, der=None, order=4):
        d = conf.dim
        der = tuple(der) if der is not None else (0,) * d
        return getattr(self.lib, f"orc_field_{d}d")(order, *der, *[float(p) for p in pos],
                                                    np.ascontiguousarray(level), C.addressof(conf))

    def f0(self, conf, f0, *xv):
        o = _f0(f0)
        return getattr(self.lib, f"orc_f0_{conf.dim}d")(C.addressof(o), *[float(a) for a in xv])

    def ftilda(self, conf, f0, n, coeffs, xv, order=4, full=False):
        o = _f0(f0)
        nm = "f" if full else "ftilda"
        return getattr(self.lib, f"orc_{nm}_{conf.dim}d")(order, n, *[float(a) for a in xv],
                                                           np.ascontiguousarray(coeffs), C.addressof(conf), C.addressof(o))

    def phase_flow(self, conf, n, coeffs, xu, order=4):
        """Foot (x, u) of the characteristic through each row of ``xu`` at t_n (dim1, rho.hpp:98-131)."""
        out = np.array(xu, dtype=np.float64).reshape(-1, 2).copy()
        fn = self.lib.orc_phase_flow_1d
        fn.argtypes = [_i, _sz, C.POINTER(C.c_double), C.POINTER(C.c_double), _dp, _p]
        fn.restype = None
        cc = np.ascontiguousarray(coeffs)
        for row in out:
            x, u = C.c_double(row[0]), C.c_double(row[1])
            fn(order, n, C.byref(x), C.byref(u), cc, C.addressof(conf))
            row[0], row[1] = x.value, u.value
        return out

    def rho(self, conf, f0, n, coeffs, l_begin=0, l_end=None, order=4):
        """CPU-convention rho (with the leading 1) for nodes [l_begin,l_end) -- the drivers' OpenMP sweep."""
        N = _nodes(conf)
        l_end = N if l_end is None else l_end
        out = np.zeros(N)
        o = _f0(f0)
        getattr(self.lib, f"orc_rho_sweep_{conf.dim}d")(order, n, np.ascontiguousarray(coeffs), C.addressof(conf),
                                                       C.addressof(o), l_begin, l_end, out)
        return out

    def rho_extended(self, conf, f0, n, coeffs, l_begin=0, l_end=None, order=4):
        """The sweep with the velocity sum carried in long double (orc_rho_sweep_extended): tells whose rounding a difference
        between two FP64 implementations is.  Not the reference's arithmetic."""
        nn = _nodes(conf)
        l_end = nn if l_end is None else l_end
        rho = np.zeros(nn)
        o = _f0(f0)
        fn = self.lib.orc_rho_sweep_extended
        fn.argtypes = [_i, _i, _sz, _dp, _p, _p, _sz, _sz, _dp]
        fn.restype = None
        fn(conf.dim, order, n, np.ascontiguousarray(coeffs), C.addressof(conf), C.addressof(o), l_begin, l_end, rho)
        return rho  # full length like rho(): entries outside [l_begin, l_end) are zero

    def rho_partial(self, conf, f0, n, coeffs, q_begin, q_end, rho=None, order=4):
        """GPU-convention partial: rho[l] += -dV f over flat q in [q_begin,q_end)."""
        out = np.zeros(_nodes(conf)) if rho is None else rho
        o = _f0(f0)
        getattr(self.lib, f"orc_rho_partial_{conf.dim}d")(order, n, np.ascontiguousarray(coeffs), C.addressof(conf),
                                                         C.addressof(o), q_begin, q_end, out)
        return out

    def metrics(self, conf, f0, n, coeffs, q_begin, q_end, order=4):
        out = np.zeros(4)
        o = _f0(f0)
        getattr(self.lib, f"orc_metrics_{conf.dim}d")(order, n, np.ascontiguousarray(coeffs), C.addressof(conf),
                                                     C.addressof(o), q_begin, q_end, out)
        return out

    def poisson(self, conf, rho):
        """Returns (phi at nodes, electric energy)."""
        data = np.array(rho, dtype=np.float64).ravel().copy()
        e = getattr(self.lib, f"orc_poisson_{conf.dim}d")(C.addressof(conf), data)
        return data, float(e)

    def interpolate(self, conf, values, order=4):
        level = np.zeros(_stride(conf, order))
        getattr(self.lib, f"orc_interpolate_{conf.dim}d")(order, level, np.ascontiguousarray(values, dtype=np.float64).ravel(),
                                                         C.addressof(conf))
        return level

    def run(self, conf, f0, n_end, coeffs=None, n_begin=0, order=4):
        """The CPU drivers' loop.  Returns (coeffs[(n_end) levels], energy[n_end], rho of the last step)."""
        st = _stride(conf, order)
        if coeffs is None:
            coeffs = np.zeros(n_end * st)
        energy = np.zeros(n_end)
        rho = np.zeros(_nodes(conf))
        o = _f0(f0)
        getattr(self.lib, f"orc_run_{conf.dim}d")(order, C.addressof(conf), C.addressof(o), n_begin, n_end, coeffs,
                                                 energy.ctypes.data, rho.ctypes.data)
        return coeffs, energy, rho


class Reference:
    """The real reference headers (oracle/_ref).  order is fixed at 4, as in every reference driver."""

    kind = "reference"

    @staticmethod
    def available(variant: str = "") -> bool:
        return os.path.exists(os.path.join(HERE, "_ref", f"libnufi_ref{variant}.so"))

    def __init__(self, variant: str = ""):
        """variant: "" (canonical, selectable f0), "_asis" (f0 as committed), "_fast" (-O3 AVX2+FMA)."""
        path = os.path.join(HERE, "_ref", f"libnufi_ref{variant}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = L = C.CDLL(path)
        self.path = path
        L.orc_num_threads.restype = _i
        L.ref_set_f0.argtypes = [_i, _p]
        L.ref_basis4.argtypes = [_i, _d, _dp]
        for dim, nco in ((1, 1), (2, 2), (3, 3)):
            getattr(L, f"ref_default_conf{dim}d").argtypes = [_p]
            getattr(L, f"ref_f0_{dim}d").argtypes = [_d] * (2 * nco)
            getattr(L, f"ref_f0_{dim}d").restype = _d
            getattr(L, f"ref_field_{dim}d").argtypes = [_i] * nco + [_d] * nco + [_dp, _p]
            getattr(L, f"ref_field_{dim}d").restype = _d
            for nm in ("ftilda", "f"):
                fn = getattr(L, f"ref_{nm}_{dim}d")
                fn.argtypes = [_sz] + [_d] * (2 * nco) + [_dp, _p]
                fn.restype = _d
            getattr(L, f"ref_rho_sweep_{dim}d").argtypes = [_sz, _dp, _p, _sz, _sz, _dp]
            getattr(L, f"ref_interpolate_{dim}d").argtypes = [_dp, _dp, _p]
            getattr(L, f"ref_run_{dim}d").argtypes = [_p, _sz, _sz, _dp, _p, _p]
        self.has_orders = hasattr(L, "ref_basis_order")
        if self.has_orders:
            L.ref_basis_order.argtypes = [_i, _i, _d, _dp]
            for dim in (1, 2, 3):
                getattr(L, f"ref_rho_sweep_order_{dim}d").argtypes = [_i, _sz, _dp, _p, _sz, _sz, _dp]
                getattr(L, f"ref_interpolate_order_{dim}d").argtypes = [_i, _dp, _dp, _p]
        self.selectable = bool(L.ref_f0_selectable())

    def threads(self) -> int:
        return int(self.lib.orc_num_threads())

    def set_threads(self, n: int) -> int:
        self.lib.orc_set_num_threads(int(n))
        return self.threads()

    def set_f0(self, dim, f0):
        o = _f0(f0)
        if self.lib.ref_set_f0(dim, C.addressof(o)) != 0:
            raise RuntimeError("this oracle/_ref build has f0 fixed as committed in nufi/config.hpp")

    def default_conf(self, conf):
        """Overwrite ``conf`` with the reference's default-constructed config_t<double>."""
        getattr(self.lib, f"ref_default_conf{conf.dim}d")(C.addressof(conf))
        return conf

    def basis(self, der, x):
        out = np.zeros(4)
        self.lib.ref_basis4(der, float(x), out)
        return out

    def f0(self, conf, *xv):
        return getattr(self.lib, f"ref_f0_{conf.dim}d")(*[float(a) for a in xv])

    # the reference templates at spline orders 3..8 (every driver runs 4; the library is generic)
    def basis_order(self, order, der, x):
        out = np.zeros(order)
        assert self.lib.ref_basis_order(order, der, float(x), out) == 0
        return out

    def rho_order(self, conf, f0, order, n, coeffs, l_begin=0, l_end=None):
        if f0 is not None:
            self.set_f0(conf.dim, f0)
        nn = _nodes(conf)
        l_end = nn if l_end is None else l_end
        rho = np.zeros(nn)
        assert getattr(self.lib, f"ref_rho_sweep_order_{conf.dim}d")(order, n, np.ascontiguousarray(coeffs), C.addressof(conf),
                                                                     l_begin, l_end, rho) == 0
        return rho[l_begin:l_end] if (l_begin, l_end) != (0, nn) else rho

    def interpolate_order(self, conf, order, values):
        level = np.zeros(_stride(conf, order))
        assert getattr(self.lib, f"ref_interpolate_order_{conf.dim}d")(order, level, np.ascontiguousarray(values, dtype=np.float64).ravel(),
                                                                       C.addressof(conf)) == 0
        return level

    def field(self, conf, level, pos, der=None):
        d = conf.dim
        der = tuple(der) if der is not None else (0,) * d
        return getattr(self.lib, f"ref_field_{d}d")(*der, *[float(p) for p in pos], np.ascontiguousarray(level),
                                                    C.addressof(conf))

    def ftilda(self, conf, f0, n, coeffs, xv, full=False):
        if f0 is not None:
            self.set_f0(conf.dim, f0)
        nm = "f" if full else "ftilda"
        return getattr(self.lib, f"ref_{nm}_{conf.dim}d")(n, *[float(a) for a in xv], np.ascontiguousarray(coeffs),
                                                           C.addressof(conf))

    def phase_flow(self, conf, n, coeffs, xu):
        """nufi::dim1::eval_phase_flow<double,4> of the real reference on each row (x, u)."""
        out = np.array(xu, dtype=np.float64).reshape(-1, 2).copy()
        fn = self.lib.ref_phase_flow_1d
        fn.argtypes = [_sz, C.POINTER(C.c_double), C.POINTER(C.c_double), _dp, _p]
        fn.restype = None
        cc = np.ascontiguousarray(coeffs)
        for row in out:
            x, u = C.c_double(row[0]), C.c_double(row[1])
            fn(n, C.byref(x), C.byref(u), cc, C.addressof(conf))
            row[0], row[1] = x.value, u.value
        return out

    def rho(self, conf, f0, n, coeffs, l_begin=0, l_end=None):
        if f0 is not None:
            self.set_f0(conf.dim, f0)
        N = _nodes(conf)
        l_end = N if l_end is None else l_end
        out = np.zeros(N)
        getattr(self.lib, f"ref_rho_sweep_{conf.dim}d")(n, np.ascontiguousarray(coeffs), C.addressof(conf), l_begin, l_end, out)
        return out

    def interpolate(self, conf, values):
        level = np.zeros(_stride(conf))
        getattr(self.lib, f"ref_interpolate_{conf.dim}d")(level, np.ascontiguousarray(values, dtype=np.float64).ravel(),
                                                         C.addressof(conf))
        return level

    def run(self, conf, f0, n_end, coeffs=None, n_begin=0):
        if f0 is not None:
            self.set_f0(conf.dim, f0)
        st = _stride(conf)
        if coeffs is None:
            coeffs = np.zeros(n_end * st)
        energy = np.zeros(n_end)
        rho = np.zeros(_nodes(conf))
        getattr(self.lib, f"ref_run_{conf.dim}d")(C.addressof(conf), n_begin, n_end, coeffs, energy.ctypes.data,
                                                 rho.ctypes.data)
        return coeffs, energy, rho


class ReferenceCuda:
    """The reference's own CUDA path (nufi/cuda_kernel.cu compiled unmodified for sm_100a, oracle/_ref/libnufi_refcuda.so):
    the informational GPU baseline of bench.py.  f0 is the one committed in the reference's config.hpp."""

    kind = "reference-cuda"

    @staticmethod
    def available() -> bool:
        return os.path.exists(os.path.join(HERE, "_ref", "libnufi_refcuda.so"))

    def __init__(self, conf, device: int = 0):
        self.lib = L = C.CDLL(os.path.join(HERE, "_ref", "libnufi_refcuda.so"))
        self.conf, self.d = conf, conf.dim
        for nm, args, res in (("create", [_p, _i], _p), ("destroy", [_p], None), ("upload", [_p, _sz, _dp], _i),
                              ("rho", [_p, _sz, _sz, _sz, _dp, _i, C.POINTER(C.c_float)], _i)):
            f = getattr(L, f"refcuda_{nm}_{self.d}d")
            f.argtypes, f.restype = args, res
        self.h = getattr(L, f"refcuda_create_{self.d}d")(C.addressof(conf), device)
        if not self.h:
            raise RuntimeError("reference cuda_kernel could not be created")

    def upload(self, coeffs, n_levels: int):
        if getattr(self.lib, f"refcuda_upload_{self.d}d")(self.h, n_levels, np.ascontiguousarray(coeffs)) != 0:
            raise RuntimeError("reference upload_phi failed")

    def rho(self, n: int, reps: int = 3):
        """(rho in the CPU convention 1 + partial, ms per compute_rho call) over the whole quadrature range."""
        N = _nodes(self.conf)
        nq = N * int(np.prod([getattr(self.conf, k) for k in ("Nu", "Nv", "Nw")[: self.d]]))
        out = np.zeros(N)
        ms = C.c_float(0)
        if getattr(self.lib, f"refcuda_rho_{self.d}d")(self.h, n, 0, nq, out, reps, C.byref(ms)) != 0:
            raise RuntimeError("reference compute_rho failed")
        return 1.0 + out, float(ms.value)

    def close(self):
        if self.h:
            getattr(self.lib, f"refcuda_destroy_{self.d}d")(self.h)
            self.h = None


def synthetic_history(conf, n_levels: int, seed: int = 1234, amp: float = 1e-2, order: int = 4) -> np.ndarray:
    """Smooth random-amplitude sine potentials interpolated to spline levels, in the spirit of the reference's
    isolated-step harness (bin/test_nufi_cpu_3d_isolated.cpp:76-105): level m holds the interpolant of
    a_m * prod_d sin(2 pi (x_d - x_d,min) / L_d + phase_m,d), a_m ~ U(-amp, amp).  Built with numpy only
    (exact circulant solve through the FFT), so it is available without any compiled oracle."""
    rng = np.random.default_rng(seed)
    dims = [conf.Nx] + ([conf.Ny] if conf.dim >= 2 else []) + ([conf.Nz] if conf.dim >= 3 else [])
    out = np.zeros((n_levels, _stride(conf, order)))
    grids = np.meshgrid(*[np.arange(n) / n for n in reversed(dims)], indexing="ij")  # z, y, x order
    lam = []
    for n in reversed(dims):
        k = np.arange(n)
        w = np.exp(2j * np.pi * k / n)
        lam.append((1 + 4 * w + w * w) / 6)
    for m in range(n_levels):
        a = rng.uniform(-amp, amp)
        ph = rng.uniform(0, 2 * np.pi, size=len(dims))
        mode = rng.integers(1, 3, size=len(dims))
        vals = a * np.ones_like(grids[0])
        for g, p, md in zip(grids, ph, mode):
            vals = vals * np.sin(2 * np.pi * md * g + p)
        spec = np.fft.fftn(vals)
        sym = lam[0]
        for ax in range(1, len(dims)):
            sym = sym[..., None] * lam[ax]
        c = np.real(np.fft.ifftn(spec / sym))
        idx = [np.arange(n + order - 1) % n for n in reversed(dims)]
        out[m] = c[np.ix_(*idx)].ravel()
    return out.ravel()


def exact_history(conf, n_levels: int, amp: float = 1e-2, order: int = 4) -> np.ndarray:
    """A coefficient history that every IEEE-754 machine reproduces BIT FOR BIT: only +, -, *, / on doubles (each correctly
    rounded, no transcendental function, no FFT), so large teacher-forced inputs need not be stored -- the golden files
    tests/golden/large_*.npz hold only the reference's rho on a few nodes of it.  Level m: periodic coefficients
    a_m * prod_d b(frac(i_d / N_d + s_{m,d})),  b(t) = 16 t^2 (1-t)^2  (a C^1 periodic bump),  a_m and the shifts s dyadic
    rationals from an integer hash of m; then the (order-1) halo by index wrap."""
    dims = [conf.Nx] + ([conf.Ny] if conf.dim >= 2 else []) + ([conf.Nz] if conf.dim >= 3 else [])
    out = np.zeros((n_levels, _stride(conf, order)))
    for m in range(n_levels):
        h = (m * 2654435761 + 12345) % (1 << 32)
        a = amp * (((h >> 8) % 4096) / 2048.0 - 1.0)
        field = np.array(a)
        for ax, n in enumerate(reversed(dims)):  # z, y, x order: x fastest in the flattened level
            s = ((h >> (3 * ax + 1)) % 16) / 16.0
            t = np.arange(n) / float(n) + s
            t = t - np.floor(t)
            b = 16.0 * (t * t) * ((1.0 - t) * (1.0 - t))
            field = np.multiply.outer(field, b)
        idx = [np.arange(n + order - 1) % n for n in reversed(dims)]
        out[m] = field[np.ix_(*idx)].ravel()
    return out.ravel()
