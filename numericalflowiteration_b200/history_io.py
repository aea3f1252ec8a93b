"""Coefficient-history files (checkpoint / restart, exchange with the reference's dumps) -- Python side of
include/nufi/history_io.hpp; same formats:

* binary: 64-byte header (magic ``NUFIB200``, version, dim, order, Nx, Ny, Nz, n_levels, dt) + float64 levels in the reference
  layout; exact, restart is bit-identical.
* "isolated-step" text (bin/test_nufi_cpu_3d_isolated.cpp:64-73, 160-163, 190-211): seven header lines, one value per line.
* plain text (bin/test_nufi_gpu_1d.cpp:239, 364-366, 109-123): one value per line.
"""
from __future__ import annotations

import struct

import numpy as np

from .config import stride_t

_HDR = struct.Struct("<8sIIIIQQQQd")
assert _HDR.size == 64


def _dims(conf):
    return conf.Nx, (conf.Ny if conf.dim >= 2 else 1), (conf.Nz if conf.dim >= 3 else 1)


def write_binary(path, conf, coeffs, n_levels: int, order: int = 4) -> None:
    coeffs = np.ascontiguousarray(coeffs, dtype=np.float64).ravel()[: n_levels * stride_t(conf, order)]
    nx, ny, nz = _dims(conf)
    with open(path, "wb") as f:
        f.write(_HDR.pack(b"NUFIB200", 1, conf.dim, order, 0, nx, ny, nz, n_levels, conf.dt))
        f.write(coeffs.tobytes())


def read_binary(path):
    """Returns (header dict, coeffs)."""
    with open(path, "rb") as f:
        magic, version, dim, order, _, nx, ny, nz, n_levels, dt = _HDR.unpack(f.read(64))
        if magic != b"NUFIB200" or version != 1:
            raise ValueError(f"{path}: not a nufi-b200 history")
        lim = 1 << 20  # do not trust the header: every field bounded before it sizes anything (include/nufi/history_io.hpp)
        if not (1 <= dim <= 3 and 1 <= order <= 8 and 1 <= nx <= lim and (dim < 2 or 1 <= ny <= lim) and (dim < 3 or 1 <= nz <= lim)
                and n_levels <= lim):
            raise ValueError(f"{path}: implausible header")
        coeffs = np.frombuffer(f.read(), dtype=np.float64).copy()
    hdr = dict(dim=dim, order=order, Nx=nx, Ny=ny, Nz=nz, n_levels=n_levels, dt=dt)
    o = order - 1
    st = (nx + o) * (ny + o if dim >= 2 else 1) * (nz + o if dim >= 3 else 1)
    if coeffs.size != n_levels * st:
        raise ValueError(f"{path}: truncated history")
    return hdr, coeffs


def write_text_isolated(path, conf, coeffs, n_levels: int, order: int = 4, reference_precision: bool = False) -> None:
    nx, ny, nz = _dims(conf)
    coeffs = np.asarray(coeffs, dtype=np.float64).ravel()[: n_levels * stride_t(conf, order)]
    with open(path, "w") as f:
        f.write(f"Nt = {n_levels - 1}\ndt = {conf.dt:f}\nNx = {nx}\nNy = {ny}\nNz = {nz}\norder = {order}\n\n")
        fmt = "%.16g\n" if reference_precision else "%.17g\n"
        f.write("".join(fmt % v for v in coeffs))


def read_text(path, n_values: int, header_lines: int = 0) -> np.ndarray:
    """header_lines = 7 for the isolated-step format, 0 for the plain format."""
    out = np.loadtxt(path, skiprows=header_lines, max_rows=n_values, dtype=np.float64)
    if out.size < n_values:
        raise ValueError(f"{path}: too few coefficients")
    return out.ravel()


def write_text_plain(path, coeffs, precision: int = 17) -> None:
    np.savetxt(path, np.asarray(coeffs, dtype=np.float64).ravel(), fmt=f"%.{precision}g")
