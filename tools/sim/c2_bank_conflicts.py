#!/usr/bin/env python
"""Development tool: shared-memory bank-conflict model of the 1d backtrace kernel on the C2 history (saturated two-stream).

Traces every quadrature point of C2 (256 x 512) back through the n = 800 levels of the reference CPU loop's history (numpy,
vectorised) and counts the LDS.64 wavefronts of the kernel's lane layout: a warp = 32 x-neighbouring nodes sharing one velocity;
a 64-bit shared-memory load is served per half-warp; two lanes of a half-warp conflict iff their cells differ but are congruent
mod 16 (level format: 3 doubles per cell -> bank pair (3 c + j) mod 16).  Then evaluates lane re-packing strategies.

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


def repack(c):
    """c: [..., 32] cells of a warp.  Returns a permutation (slot -> lane) that puts, for every residue class mod 16, the first
    distinct... simple model: lanes ranked within their residue class (distinct cells only; duplicates of a cell ride along);
    rank 0 -> half-warp 0, rank 1 -> half-warp 1, others fill the free slots."""
    shp = c.shape[:-1]
    c2 = c.reshape(-1, 32)
    out = np.empty_like(c2)
    for w in range(c2.shape[0]):
        cw = c2[w]
        halves = [[], []]
        left = []
        seen = {}
        for lane in range(32):
            r = cw[lane] & 15
            key = (r, cw[lane])
            if key in seen:  # same cell as an earlier lane: broadcast, goes wherever that one went if room
                h = seen[key]
                if len(halves[h]) < 16:
                    halves[h].append(lane)
                    continue
            used = [k for k in seen if k[0] == r]
            rank = len(set(used))
            if rank < 2 and len(halves[rank]) < 16:
                halves[rank].append(lane)
                seen[key] = rank
            else:
                left.append(lane)
        for lane in left:
            h = 0 if len(halves[0]) < 16 else 1
            halves[h].append(lane)
        out[w] = cw[np.array(halves[0] + halves[1])]
    return out.reshape(*shp, 32)


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
    # re-pack every K levels, keep the permutation for the next K levels
    sub = tiles[:, :, ::8, :]  # every 8th velocity to keep the python loop affordable
    for K in (1, 4, 16, 64):
        tot = 0
        cnt = 0
        for l0 in range(0, L, K):
            blk = sub[l0:l0 + K]
            perm_src = repack(blk[0])  # cells after re-pack at the first level of the block
            # derive the permutation indices: recompute on lane ids
            ids = np.broadcast_to(np.arange(32), blk[0].shape)
            # permutation by matching: redo repack on (cell*64+lane) trick
            keyed = blk[0].astype(np.int64) * 64 + ids
            order = np.empty_like(keyed)
            flat_c = blk[0].reshape(-1, 32)
            flat_o = order.reshape(-1, 32)
            rp = repack(blk[0]).reshape(-1, 32)
            for w in range(flat_c.shape[0]):
                # map re-packed cells back to lanes (stable for duplicates)
                lanes = list(range(32))
                res = []
                for cval in rp[w]:
                    for j, ln in enumerate(lanes):
                        if flat_c[w, ln] == cval:
                            res.append(ln)
                            lanes.pop(j)
                            break
                flat_o[w] = res
            perm = order  # [tile, vel, 32] slot -> lane
            for l in range(blk.shape[0]):
                c = np.take_along_axis(blk[l], perm, axis=-1)
                tot += halfwarp_degree(c.reshape(*c.shape[:-1], 2, 16)).sum()
                cnt += c.size // 16
        print(f"re-pack every {K:3d} levels: wavefront ratio {tot / cnt:.3f}")


if __name__ == "__main__":
    main()
