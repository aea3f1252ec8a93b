#!/usr/bin/env python
"""Development tool: shared-memory bank-conflict model of the 1d backtrace kernel on the C2 history (saturated two-stream).

Traces every quadrature point of C2 (256 x 512) back through the n = 800 levels of the reference CPU loop's history (numpy,
vectorised) and counts the LDS.64 wavefronts of the kernel's lane layout: a warp = 32 x-neighbouring nodes sharing one velocity;
a 64-bit shared-memory load is served per half-warp; two lanes of a half-warp conflict iff their cells differ but are congruent
mod 16 (level format: 3 doubles per cell -> bank pair (3 c + j) mod 16).  Then evaluates what wider loads would change.

    python tools/sim/c2_bank_conflicts.py [history.npy]
"""
import sys
import numpy as np

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from bench import make_workload, reference_history  # noqa: E402


def dbasis(t):
    s = 1 - t
    return np.stack([-s * s / 2, (3 * t * t - 4 * t) / 2, -(3 * s * s - 4 * s) / 2, t * t / 2])


def trace_cells(conf, coeffs, n, nodes=None):
    Nx, Nu = conf.Nx, conf.Nu
    st = Nx + 3
    ix = np.arange(Nx) if nodes is None else nodes
    du = (conf.u_max - conf.u_min) / Nu
    u = conf.u_min + 0.5 * du + du * np.arange(Nu)
    x = np.repeat((conf.x_min + ix * conf.dx)[:, None], Nu, 1).astype(np.float64)  # [node, vel]
    v = np.repeat(u[None, :], len(ix), 0).copy()
    cells = np.empty((n, len(ix), Nu), dtype=np.int16)
    for m in range(n - 1, -1, -1):
        x -= conf.dt * v
        xs = x - conf.x_min
        xs -= conf.Lx * np.floor(xs * conf.Lx_inv)
        kf = np.floor(xs * conf.dx_inv)
        k = kf.astype(np.int64) % Nx
        t = xs * conf.dx_inv - kf
        D = dbasis(t)
        lev = coeffs[m * st:(m + 1) * st]
        dphi = sum(lev[k + a] * D[a] for a in range(4)) * conf.dx_inv
        v += (conf.dt if m > 0 else 0.5 * conf.dt) * (-dphi)
        cells[n - 1 - m] = k
    return cells  # [level (newest first), node, vel]


def halfwarp_degree(c):
    """c: [..., 16] cells of a half-warp -> max number of DISTINCT cells per residue class mod 16 (wavefronts of one LDS.64)."""
    c = np.sort(c, axis=-1)
    res = c & 15
    # count distinct cells per residue: mark first occurrence of each distinct cell
    first = np.ones(c.shape, dtype=bool)
    first[..., 1:] = c[..., 1:] != c[..., :-1]
    deg = np.zeros(c.shape[:-1], dtype=np.int64)
    for r in range(16):
        deg = np.maximum(deg, np.sum(first & (res == r), axis=-1))
    return deg


def halfwarp_degree_classes(c, cls):
    """c, cls: [..., 16] cells and their bank classes -> max number of DISTINCT cells per class."""
    order = np.argsort(c, axis=-1)
    c = np.take_along_axis(c, order, axis=-1)
    cls = np.take_along_axis(cls, order, axis=-1)
    first = np.ones(c.shape, dtype=bool)
    first[..., 1:] = c[..., 1:] != c[..., :-1]
    deg = np.zeros(c.shape[:-1], dtype=np.int64)
    for r in range(16):
        deg = np.maximum(deg, np.sum(first & (cls == r), axis=-1))
    return deg


def group_degree(c, lanes):
    """c: [..., lanes] cells of one wavefront group (16 lanes for a 64-bit load, 8 for a 128-bit load whose lane stride is 16 B):
    max number of DISTINCT cells per residue class mod `lanes` = wavefronts that group's load takes."""
    c = np.sort(c, axis=-1)
    res = c % lanes
    first = np.ones(c.shape, dtype=bool)
    first[..., 1:] = c[..., 1:] != c[..., :-1]
    deg = np.zeros(c.shape[:-1], dtype=np.int64)
    for r in range(lanes):
        deg = np.maximum(deg, np.sum(first & (res == r), axis=-1))
    return deg


def main():
    conf, f0, n, _ = make_workload("C2", 1)
    coeffs = np.load(sys.argv[1]) if len(sys.argv) > 1 else reference_history("C2", n)[0]
    cells = trace_cells(conf, coeffs, n)  # [800, 256, 512]
    L, Nx, Nu = cells.shape
    tiles = cells.reshape(L, Nx // 32, 32, Nu).transpose(0, 1, 3, 2)  # [level, tile, vel, lane]
    hw = tiles.reshape(L, Nx // 32, Nu, 2, 16)
    deg = halfwarp_degree(hw)  # [level, tile, vel, half]
    base = deg.sum()
    ideal = deg.size
    print(f"baseline: wavefront ratio {base / ideal:.3f} (ncu: 33.4 M / 19.7 M = 1.70)")
    per_vel = deg.sum(axis=(0, 1, 3)) / (L * (Nx // 32) * 2)
    print("per-velocity ratio (every 32nd):", np.round(per_vel[::32], 2))
    by_age = deg.mean(axis=(1, 2, 3))
    print("by history age (newest first, every 100 levels):", np.round(by_age[::100], 2))
    # what wider loads would buy: a 128-bit load is served per quarter-warp (8 lanes), so it takes a stretch of 1/7 instead of 1/15
    q = group_degree(tiles.reshape(L, Nx // 32, Nu, 4, 8), 8)
    r128 = q.sum() / q.size
    print(f"128-bit loads (8-lane groups): wavefront ratio {r128:.3f}")
    now = 3 * 2 * base / ideal
    print(f"wavefronts per warp-step: now (3 x LDS.64) {now:.2f};  (p0,p1) as one LDS.128 + p2 as LDS.64: {4 * r128 + 2 * base / ideal:.2f};"
          f"  two LDS.128 (padded cell): {8 * r128:.2f};  conflict-free floor 6")
    # (i) a second copy of every staged level at a different bank phase, the lane picks the copy by (cell >> 4) & 1
    #     (round-1 review's experiment): class(c) = (c + 8 * ((c >> 4) & 1)) mod 16 -- cells 16 apart no longer collide, but cells
    #     8 and 24 apart now do: a half-warp stretched over more than 16 cells has more lanes than free classes either way
    c64 = hw.astype(np.int64)
    alt = c64 + 8 * ((c64 >> 4) & 1)
    deg_i = halfwarp_degree_classes(c64, alt & 15)
    print(f"(i) two copies, bank phase by (cell>>4)&1: wavefront ratio {deg_i.sum() / deg_i.size:.3f} (baseline {base / ideal:.3f})")
    # (i') four copies at phases 0,4,8,12 by (cell>>4)&3
    alt4 = c64 + 4 * ((c64 >> 4) & 3)
    deg_i4 = halfwarp_degree_classes(c64, alt4 & 15)
    print(f"(i') four copies, phase by (cell>>4)&3:     wavefront ratio {deg_i4.sum() / deg_i4.size:.3f}")
    # (ii) re-pack the lanes of a CTA-round (one tile, the 15 warps x 2 points of a round = 30 velocities spread over the velocity
    #      range, 32 nodes each = 960 points) so that the 16 lanes of every half-warp have distinct cells mod 16: the best any
    #      re-packing can do at the level it is made for is max(ceil(P/16), largest residue class) half-warp loads for P points;
    #      one level later the points have moved by v*dt/dx cells -- a different amount for every velocity -- and the packing is
    #      as good as random.  Measured: ideal packing at the level itself, then the SAME packing 1, 2, 4 levels later.
    rng = np.random.default_rng(0)
    vel_sets = [np.arange(j, Nu, 18)[:30] for j in range(0, 18, 3)]  # interleaved velocity sets like a CTA-round's
    levels = [100, 300, 500, 700]
    stats = {0: [], 1: [], 2: [], 4: []}
    for m0 in levels:
        for t in range(Nx // 32):
            for vs in vel_sets:
                pts = cells[m0, 32 * t:32 * t + 32][:, vs].reshape(-1)
                P_ = (pts.size // 16) * 16
                nh = P_ // 16
                order = np.argsort(pts[:P_] & 15, kind="stable")  # deal the residues round-robin into the half-warps
                groups = np.empty(P_, dtype=np.int64)
                groups[order] = np.arange(P_) % nh
                slot = np.empty(P_, dtype=np.int64)
                slot[order] = np.arange(P_) // nh
                for lag in stats:
                    if m0 + lag >= L:
                        continue
                    later = cells[m0 + lag, 32 * t:32 * t + 32][:, vs].reshape(-1)[:P_].astype(np.int64)
                    g = np.zeros((nh, 16), dtype=np.int64)
                    g[groups, slot] = later
                    stats[lag].append(halfwarp_degree(g).mean())
    print("(ii) re-packed half-warps (dealt by cell mod 16): wavefront ratio at the level of the re-pack / 1 / 2 / 4 levels later: " +
          " / ".join(f"{np.mean(stats[k]):.2f}" for k in (0, 1, 2, 4)))
    span = tiles.max(axis=-1).astype(np.int64) - tiles.min(axis=-1)
    print("warp span in cells (32 lanes, 31 = rigid): percentiles 50/90/99:", np.percentile(span, [50, 90, 99]))


if __name__ == "__main__":
    main()
