"""Multi-GPU sharding of the quadrature points: one process per GPU (torchrun), ``torch.distributed`` for the
plumbing.  The path shards by independent quadrature points: each rank traces its contiguous
range of the flat index q; the only exchange per step is the sum of the small partial-rho vectors
(``MPI_Allreduce`` on host buffers in the reference, bin/test_nufi_gpu_3d.cpp:158; here an NCCL all-reduce on
the device buffer), after which every rank runs the tiny deterministic field tail redundantly -- exactly
what each MPI rank of the reference does (bin/test_nufi_gpu_3d.cpp:160-162) -- so no broadcast is needed.
"""
from __future__ import annotations

import numpy as np

__all__ = ["partition", "velocity_share", "DistributedStepper", "attach_peers"]


def partition(n_total: int, n_parts: int, part: int, begin: int = 0) -> tuple[int, int]:
    """Contiguous near-equal split of ``[begin, begin+n_total)``; the first ``n_total % n_parts`` parts get one
    more element -- the reference's rule (nufi/cuda_scheduler.hpp:88-111, bin/test_nufi_gpu_3d.cpp:80-105)."""
    chunk, rem = divmod(n_total, n_parts)
    lo = begin + part * chunk + min(part, rem)
    hi = lo + chunk + (1 if part < rem else 0)
    return lo, hi


def velocity_share(n_vel: int, n_parts: int, part: int) -> range:
    """Velocity nodes rank ``part`` traces in the fused multi-GPU step (``peer_step`` / ``group_step``): every
    ``n_parts``-th node starting at ``part``, for EVERY spatial node -- all GPUs integrate statistically identical samples of
    phase space (csrc/peer.cu).  Flat quadrature indices of the share: ``q = l * n_vel + j`` for ``j`` in the range."""
    return range(part, n_vel, n_parts)


class DistributedStepper:
    """Free-running NuFI loop over ``world_size`` GPUs.

    ``backend='nccl'``: partial rho stays on the device, all-reduced in place through a torch tensor that
    aliases the library's buffer.  ``backend='gloo'`` (CPU tests of the host logic): ``compute`` is a
    callable standing in for the device work, the reduction runs on host tensors.
    """

    def __init__(self, sched, rank: int | None = None, world_size: int | None = None, group=None, exchange: str = "nccl"):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.sched = sched
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world_size is None else world_size
        self.q_begin, self.q_end = partition(sched.n_quad, self.world, self.rank)
        self._rho_t = None
        self._stream = None
        if torch.cuda.is_available() and hasattr(sched, "rho_device_ptr"):
            # the library and the NCCL all-reduce share ONE explicit torch stream, so the collective is ordered after the
            # backtrace and before the tail without host syncs (torch's legacy default stream would map to the handle's own
            # non-blocking stream, which NCCL does not order against)
            self._stream = torch.cuda.Stream()
            sched.set_stream(self._stream.cuda_stream)
            self._rho_t = _alias_device_f64(torch, sched.rho_device_ptr(), sched.n_nodes)
        self.exchange = "nccl"
        if exchange == "peer" and self.world > 1:
            attach_peers(sched, self.rank, self.world, dist, group)
            self.exchange = "peer-memory"

    def compute_rho(self, n: int):
        """Local partial rho of step n on the device (asynchronous)."""
        self.sched.compute_rho(n, self.q_begin, self.q_end)

    def reduce_rho(self):
        """Sum of the partial rho vectors over all ranks, in place on the device."""
        if self.world > 1:
            if self._stream is not None:
                with self.torch.cuda.stream(self._stream):
                    self.dist.all_reduce(self._rho_t, op=self.dist.ReduceOp.SUM, group=self.group)
            else:
                self.dist.all_reduce(self._rho_t, op=self.dist.ReduceOp.SUM, group=self.group)
        return self._rho_t

    def step(self, n: int) -> None:
        """One time step: local backtrace -> exchange of rho -> replicated field tail -> level n."""
        if self.exchange == "peer-memory":
            self.sched.peer_step(n)  # exchange fused into the kernels (self-validating stores into peer memory)
            return
        self.compute_rho(n)
        self.reduce_rho()
        self.sched.field_tail_device(n, self._rho_t.data_ptr())


def attach_peers(sched, rank: int, world: int, dist, group=None) -> None:
    """Maps every rank's exchange buffer into every other rank (CUDA IPC handles all-gathered through torch.distributed),
    after which ``sched.peer_step(n)`` needs no collective call.  Raises CudaError if IPC mapping is not possible."""
    mine = sched.peer_export(world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    sched.peer_attach(rank, world, b"".join(gathered))
    dist.barrier(group=group)  # everyone has mapped everyone before the first push


class _CudaArray:
    """Minimal __cuda_array_interface__ carrier so torch can alias memory owned by libnufi_b200."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {
            "shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None,
        }


def _alias_device_f64(torch, ptr: int, n: int):
    return torch.as_tensor(_CudaArray(ptr, n), device="cuda")


def host_allreduce_partials(dist, partial: np.ndarray, group=None) -> np.ndarray:
    """gloo path used by the CPU tests of the sharding logic: sum partial rho over ranks on the host."""
    import torch

    t = torch.from_numpy(np.ascontiguousarray(partial, dtype=np.float64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy()
