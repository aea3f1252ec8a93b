# Development probe: is the staged 1d kernel limited by the L2 -> shared-memory fill?  Same per-SM work (one CTA-round of 15 warps x 2
# points through 800 levels) with 144, 72 and 36 CTAs pulling the history at once (Nu = 512, 256, 128): if the kernel time falls
# with the CTA count, the fill path (L2 slices serving the same lines to every SM) is what the consumers wait for.
import sys
sys.path.insert(0, '.')
import torch
from numericalflowiteration_b200 import Config1D, CudaScheduler, F0
for name, f0 in (("C1", F0(0, 0.01, 0.5)), ("C2", F0(1, 0.01, 0.5))):
    for nu in (512, 256, 128):
        conf = Config1D(Nu=nu)
        s = CudaScheduler(conf, f0, device=0)
        for m in range(800):
            s.step(m)
        s.sync()
        s.set_kernel_timing(True)
        for _ in range(3):
            s.compute_rho(800, 0, s.n_quad)
        s.sync(); s.backtrace_time(reset=True)
        for _ in range(20):
            s.compute_rho(800, 0, s.n_quad)
        ms, cnt = s.backtrace_time(reset=True)
        print(f"{name} Nu={nu:4d} {s.last_variant:28s} kernel {ms / cnt:.4f} ms", flush=True)
        s.close()
