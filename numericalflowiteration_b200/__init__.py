"""numericalflowiteration_b200 -- B200 (sm_100a) implementation of NuFI's data-parallel hot path.

Scope: per time step, every quadrature point is traced backwards through the stored
history of potential spline coefficients, f0 is evaluated at the foot and reduced into rho; the Poisson
solve and spline interpolation that close the step run on the device too.  The product is the C-ABI
library ``lib/libnufi_b200.so`` (``include/nufi_b200.h``); this package is its host-side mirror of the
reference's ``config_t`` / ``cuda_scheduler`` interface.  There is no CPU fallback.
"""
from .config import Config1D, Config2D, Config3D, F0, n_nodes, n_quad, n_vel, stride_t
from .scheduler import CudaError, CudaGroup, CudaScheduler, RangeError, device_count, measure_fp64_peak
from .distributed import partition, DistributedStepper
from . import history_io

__all__ = [
    "Config1D", "Config2D", "Config3D", "F0", "n_nodes", "n_quad", "n_vel", "stride_t",
    "CudaScheduler", "CudaGroup", "device_count", "CudaError", "RangeError", "measure_fp64_peak", "partition", "DistributedStepper", "history_io",
]
