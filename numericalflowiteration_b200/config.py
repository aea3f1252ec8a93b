"""Host-side mirror of ``nufi::dim{1,2,3}::config_t<double>`` (reference nufi/config.hpp:33-70, 91-138,
167-219).

The classes are ``ctypes.Structure`` s whose field order and types equal the reference structs, so the same
bytes are handed to the C ABI (``include/nufi_b200.h``: ``nufi_b200_config{1,2,3}d``) that a C++ caller would
pass by ``reinterpret_cast`` from its ``config_t<double>``.  As in the reference, all fields are public and
mutable; the default constructor sets the reference's defaults and derived quantities; after editing the
primary fields call :meth:`derive` (the reference's tests recompute the derived fields by hand,
bin/test_poisson.cpp:60-79).

Initial conditions: the reference selects ``f0`` by editing config.hpp; here it is a small POD
(:class:`F0`) of ``kind`` + parameters covering every (active or commented) expression in config.hpp.
"""
from __future__ import annotations

import ctypes as C
import math

__all__ = ["Config1D", "Config2D", "Config3D", "F0", "stride_t", "n_nodes", "n_vel", "n_quad"]

_sz = C.c_size_t
_d = C.c_double


class F0(C.Structure):
    """f0 selector.  kinds (expressions: nufi/config.hpp):

    * 1d: 0 Landau (:83), 1 two-stream (:82, the committed default); p = (alpha, k)
    * 2d: 0 Landau (:148-149, committed default with alpha=0.5, k=0.5), 1 two-stream (:151-158); p = (alpha, k, v0)
    * 3d: 0 Landau (:233-234), 1 two-stream (:237-242), 2 bump-on-tail (:244-246, committed default);
      p = (alpha, k, v0)
    """

    _fields_ = [("kind", C.c_int), ("p", _d * 4)]

    def __init__(self, kind: int = 0, *params: float):
        super().__init__()
        self.kind = int(kind)
        for i, v in enumerate(params):
            self.p[i] = float(v)

    def key(self):
        return (self.kind, tuple(self.p))

    # the reference's committed defaults
    @staticmethod
    def default(dim: int) -> "F0":
        return {1: F0(1, 0.01, 0.5), 2: F0(0, 0.5, 0.5), 3: F0(2, 0.03, 0.3)}[dim]

    @staticmethod
    def landau(dim: int, alpha: float, k: float) -> "F0":
        return F0(0, alpha, k)

    @staticmethod
    def two_stream(dim: int, alpha: float, k: float, v0: float = 2.4) -> "F0":
        return F0(1, alpha, k, v0)


class Config1D(C.Structure):
    """nufi::dim1::config_t<double> (config.hpp:33-70)."""

    dim = 1
    _fields_ = [
        ("Nx", _sz), ("Nu", _sz), ("Nt", _sz), ("dt", _d),
        ("x_min", _d), ("x_max", _d),
        ("u_min", _d), ("u_max", _d),
        ("dx", _d), ("dx_inv", _d), ("Lx", _d), ("Lx_inv", _d),
        ("du", _d),
    ]

    def __init__(self, **kw):
        super().__init__()
        self.Nx, self.Nu = 256, 512
        self.u_min, self.u_max = -10.0, 10.0
        self.x_min, self.x_max = 0.0, 4 * math.pi
        self.dt = 1.0 / 16.0
        self.Nt = int(100 / self.dt)
        for k, v in kw.items():
            setattr(self, k, v)
        self.derive()

    def derive(self) -> "Config1D":
        self.Lx = self.x_max - self.x_min
        self.Lx_inv = 1 / self.Lx
        self.dx = self.Lx / self.Nx
        self.dx_inv = 1 / self.dx
        self.du = (self.u_max - self.u_min) / self.Nu
        return self


class Config2D(C.Structure):
    """nufi::dim2::config_t<double> (config.hpp:91-138)."""

    dim = 2
    _fields_ = [
        ("Nx", _sz), ("Ny", _sz), ("Nu", _sz), ("Nv", _sz), ("Nt", _sz), ("dt", _d),
        ("x_min", _d), ("x_max", _d), ("y_min", _d), ("y_max", _d),
        ("u_min", _d), ("u_max", _d), ("v_min", _d), ("v_max", _d),
        ("dx", _d), ("dx_inv", _d), ("Lx", _d), ("Lx_inv", _d),
        ("dy", _d), ("dy_inv", _d), ("Ly", _d), ("Ly_inv", _d),
        ("du", _d), ("dv", _d),
    ]

    def __init__(self, **kw):
        super().__init__()
        self.Nx = self.Ny = 32
        self.Nu = self.Nv = 128
        self.u_min = self.v_min = -6.0
        self.u_max = self.v_max = 6.0
        self.x_min = self.y_min = 0.0
        self.x_max = self.y_max = 4.0 * math.pi
        self.dt = 1.0 / 16.0
        self.Nt = int(50.0 / self.dt)
        for k, v in kw.items():
            setattr(self, k, v)
        self.derive()

    def derive(self) -> "Config2D":
        self.Lx = self.x_max - self.x_min
        self.Lx_inv = 1 / self.Lx
        self.Ly = self.y_max - self.y_min
        self.Ly_inv = 1 / self.Ly
        self.dx = self.Lx / self.Nx
        self.dx_inv = 1 / self.dx
        self.dy = self.Ly / self.Ny
        self.dy_inv = 1 / self.dy
        self.du = (self.u_max - self.u_min) / self.Nu
        self.dv = (self.v_max - self.v_min) / self.Nv
        return self


class Config3D(C.Structure):
    """nufi::dim3::config_t<double> (config.hpp:167-219)."""

    dim = 3
    _fields_ = [
        ("Nx", _sz), ("Ny", _sz), ("Nz", _sz), ("Nu", _sz), ("Nv", _sz), ("Nw", _sz), ("Nt", _sz), ("dt", _d),
        ("x_min", _d), ("x_max", _d), ("y_min", _d), ("y_max", _d), ("z_min", _d), ("z_max", _d),
        ("u_min", _d), ("u_max", _d), ("v_min", _d), ("v_max", _d), ("w_min", _d), ("w_max", _d),
        ("dx", _d), ("dx_inv", _d), ("Lx", _d), ("Lx_inv", _d),
        ("dy", _d), ("dy_inv", _d), ("Ly", _d), ("Ly_inv", _d),
        ("dz", _d), ("dz_inv", _d), ("Lz", _d), ("Lz_inv", _d),
        ("du", _d), ("dv", _d), ("dw", _d),
    ]

    def __init__(self, **kw):
        super().__init__()
        self.Nx = self.Ny = self.Nz = 8
        self.Nu = self.Nv = self.Nw = 8
        self.u_min = self.v_min = self.w_min = -9.0
        self.u_max = self.v_max = self.w_max = 0.0
        self.x_min = self.y_min = self.z_min = 0.0
        self.x_max = self.y_max = self.z_max = 20 * math.pi / 3.0
        self.dt = 1.0 / 10.0
        self.Nt = int(5 / self.dt)
        for k, v in kw.items():
            setattr(self, k, v)
        self.derive()

    def derive(self) -> "Config3D":
        self.Lx = self.x_max - self.x_min
        self.Lx_inv = 1 / self.Lx
        self.Ly = self.y_max - self.y_min
        self.Ly_inv = 1 / self.Ly
        self.Lz = self.z_max - self.z_min
        self.Lz_inv = 1 / self.Lz
        self.dx = self.Lx / self.Nx
        self.dx_inv = 1 / self.dx
        self.dy = self.Ly / self.Ny
        self.dy_inv = 1 / self.dy
        self.dz = self.Lz / self.Nz
        self.dz_inv = 1 / self.dz
        self.du = (self.u_max - self.u_min) / self.Nu
        self.dv = (self.v_max - self.v_min) / self.Nv
        self.dw = (self.w_max - self.w_min) / self.Nw
        return self


def stride_t(conf, order: int = 4) -> int:
    """Doubles per history level, ``prod_d (N_d + order - 1)`` (rho.hpp:326-329)."""
    s = conf.Nx + order - 1
    if conf.dim >= 2:
        s *= conf.Ny + order - 1
    if conf.dim >= 3:
        s *= conf.Nz + order - 1
    return s


def n_nodes(conf) -> int:
    n = conf.Nx
    if conf.dim >= 2:
        n *= conf.Ny
    if conf.dim >= 3:
        n *= conf.Nz
    return n


def n_vel(conf) -> int:
    n = conf.Nu
    if conf.dim >= 2:
        n *= conf.Nv
    if conf.dim >= 3:
        n *= conf.Nw
    return n


def n_quad(conf) -> int:
    """Number of quadrature points of one time step (flat index q, cuda_kernel.cu:40-41,219-225,402-412)."""
    return n_nodes(conf) * n_vel(conf)
