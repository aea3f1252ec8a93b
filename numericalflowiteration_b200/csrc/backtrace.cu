// backtrace.cu -- the hot path: backward Stoermer-Verlet through the coefficient history + f0 + rho reduction.
//
// Replaces nufi/cuda_kernel.cu:31-51, 210-237, 393-426 (cuda_eval_rho) and :53-79, 239-271, 428-466
// (cuda_eval_metrics), i.e. the device flavour of nufi/rho.hpp eval_ftilda / eval_f / eval_rho with
// nufi/fields.hpp eval and nufi/splines.hpp inlined.  Written from scratch for sm_100a:
//
//  * Work layout.  A warp-unit is 32 consecutive spatial nodes (x fastest, a "tile") sharing ILP consecutive
//    velocity nodes, so the lanes of a warp drift rigidly and read neighbouring coefficients (conflict-free)
//    for the whole history -- instead of the reference's "velocity fastest" layout whose lanes fan out and all
//    hit one rho address with atomics.  A CTA-round is W warp-units of the SAME tile; persistent CTAs (one per
//    SM) take contiguous runs of CTA-rounds.  Each thread sums f of its node over its velocities in registers;
//    when the CTA's tile changes the consumer warps combine their sums through shared memory in a fixed order
//    and write ONE slot per (CTA, tile); a small kernel adds the slots of a tile in a fixed order.  No atomics:
//    results are run-to-run deterministic.
//  * History access.  Staged variant: a producer warp streams the history newest -> oldest from the HBM/L2-
//    resident ring into a shared-memory ring of stages with cp.async.bulk (TMA bulk copy, SASS UBLKCP) +
//    mbarrier full/empty pairs; a stage holds a CHUNK of several consecutive levels so barrier traffic and loop
//    bookkeeping are amortised over the chunk.  Global variant (levels too big for shared memory, large 3d):
//    read-only loads served by L1/L2.
//  * Arithmetic.  Position per dimension = (cell k, centred offset tau in [-1/2,1/2]); floor() and the
//    float->int conversion (quarter-rate pipes) are replaced by the 1.5*2^52 rounding trick on the FP64 pipe.
//    The regular full-kick step is the only thing in the inner loop: eval_f's initial half kick and the final
//    half kick on level 0 are peeled.  1d levels are stored as per-cell quadratics of dt*E (3 doubles per cell,
//    see tail.cu), so a 1d point-step is 7 FP64 instructions; 2d/3d use the cubic B-spline window (16/64
//    doubles) with value and derivative bases computed once per dimension and shared by the field components.
#include "internal.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace nufi_b200
{

namespace
{

constexpr double kMagic = 6755399441055744.0; // 1.5 * 2^52: adding it rounds to the nearest integer

// Periodic cell index after a move of dk cells.  POW2: mask.  Otherwise one conditional correction each way;
// anything further out (a point crossing more than a whole period in one step) is clamped into range and
// flagged -- such points are recomputed by the robust slow path after the trace.
template <bool POW2> __device__ __forceinline__ int wrap_cell(int k, int N, unsigned &bad)
{
    if constexpr (POW2) {
        return k & (N - 1);
    } else {
        if (k < 0) k += N;
        if (k >= N) k -= N;
        const unsigned kc = min(static_cast<unsigned>(k), static_cast<unsigned>(N - 1));
        bad |= kc ^ static_cast<unsigned>(k);
        return static_cast<int>(kc);
    }
}

// t2 = tau - drift.  New cell/offset such that k + 1/2 + tau is preserved and tau in [-1/2, 1/2].
template <bool POW2> __device__ __forceinline__ void relocate(double &tau, int &k, double t2, int N, unsigned &bad)
{
    const double y = t2 + kMagic;
    const double r = y - kMagic;
    tau = t2 - r;
    k = wrap_cell<POW2>(k + __double2loint(y), N, bad);
}

// Cubic B-spline basis on a cell, t = 1/2 + tau.  Returns 6*N_a(t) and 2*N'_a(t) (nufi/splines.hpp:39-79
// evaluates the same polynomials by the Cox-de Boor recurrence); the 1/6, 1/2 go into the kick factor.
__device__ __forceinline__ void basis4(double tau, double (&N)[4], double (&D)[4])
{
    const double t = 0.5 + tau, s = 0.5 - tau;
    const double t2 = t * t, s2 = s * s;
    N[0] = s2 * s;
    N[3] = t2 * t;
    N[1] = fma(t2, fma(3.0, t, -6.0), 4.0);
    N[2] = fma(s2, fma(3.0, s, -6.0), 4.0);
    D[0] = -s2;
    D[3] = t2;
    D[1] = t * fma(3.0, t, -4.0);
    D[2] = s * fma(-3.0, s, 4.0);
}

template <bool STAGED> __device__ __forceinline__ double ld(const double *p)
{
    if constexpr (STAGED) return *p;
    else return __ldg(p);
}

// ---------------------------------------------------------------- f0 (nufi/config.hpp:72-84, 140-159, 221-247)
__device__ __forceinline__ double f0_1d(const BtParams &P, double x, double u)
{
    const double alpha = P.f0p[0], k = P.f0p[1];
    double r = 0.39894228040143267793994 * (1. + alpha * cos(k * x)) * exp(-u * u / 2.);
    if (P.f0_kind == 1) r = r * u * u;
    return r;
}

__device__ __forceinline__ double f0_2d(const BtParams &P, double x, double y, double u, double v)
{
    const double alpha = P.f0p[0], k = P.f0p[1];
    const double pert = 1.0 + alpha * (cos(k * x) + cos(k * y));
    if (P.f0_kind == 1) {
        const double v0 = P.f0p[2];
        const double c = 1.0 / (8.0 * 3.14159265358979323846);
        const double feq = (exp(-0.5 * (v - v0) * (v - v0)) + exp(-0.5 * (v + v0) * (v + v0))) *
                           (exp(-0.5 * (u - v0) * (u - v0)) + exp(-0.5 * (u + v0) * (u + v0)));
        return c * pert * feq;
    }
    return 1.0 / (2.0 * 3.14159265358979323846) * exp(-0.5 * (u * u + v * v)) * pert;
}

__device__ __forceinline__ double f0_3d(const BtParams &P, double x, double y, double z, double u, double v, double w)
{
    const double alpha = P.f0p[0], k = P.f0p[1];
    if (P.f0_kind == 1) {
        const double c = 0.03174681796712048489288165246732, v0 = P.f0p[2];
        return c * (exp(-(v - v0) * (v - v0) / 2.0) + exp(-(v + v0) * (v + v0) / 2.0)) * exp(-(u * u + w * w) / 2) *
               (1 + alpha * (cos(k * x) + cos(k * y) + cos(k * z)));
    }
    const double c = 0.06349363593424096978576330493464;
    if (P.f0_kind == 2)
        return c * (0.9 * exp(-0.5 * u * u) + 0.2 * exp(-2 * (u - 4.5) * (u - 4.5))) * exp(-0.5 * (v * v + w * w)) *
               (1 + alpha * (cos(k * x) + cos(k * y) + cos(k * z)));
    return c * (1. + alpha * cos(k * x) + alpha * cos(k * y) + alpha * cos(k * z)) * exp(-(u * u + v * v + w * w) / 2);
}

// ---------------------------------------------------------------- one point, one history level
template <int DIM> struct Point
{
    double tau[DIM];
    double vel[DIM];
    int cell[DIM];
};

// Step kinds: FULL = drift + full kick (levels n-1..1, the inner loop); LAST = drift + half kick (level 0);
// FIRST = half kick without drift (eval_f's initial half step on level n).
enum StepKind { FULL = 0, LAST = 1, FIRST = 2 };

template <int KIND, bool STAGED, bool POW2>
__device__ __forceinline__ void step1d(Point<1> &p, const double *lev, const BtParams &P, unsigned &bad)
{
    if constexpr (KIND != FIRST) relocate<POW2>(p.tau[0], p.cell[0], fma(P.ncx, p.vel[0], p.tau[0]), P.Nx, bad);
    const double *c = lev + 3 * p.cell[0];
    const double p0 = ld<STAGED>(c), p1 = ld<STAGED>(c + 1), p2 = ld<STAGED>(c + 2);
    const double t = p.tau[0];
    const double q = fma(t, p2, p1);
    if constexpr (KIND == FULL) p.vel[0] = fma(t, q, p0 + p.vel[0]);
    else p.vel[0] = fma(0.5, fma(t, q, p0), p.vel[0]);
}

template <int KIND, bool STAGED, bool POW2>
__device__ __forceinline__ void step2d(Point<2> &p, const double *lev, const BtParams &P, unsigned &bad)
{
    if constexpr (KIND != FIRST) {
        relocate<POW2>(p.tau[0], p.cell[0], fma(P.ncx, p.vel[0], p.tau[0]), P.Nx, bad);
        relocate<POW2>(p.tau[1], p.cell[1], fma(P.ncy, p.vel[1], p.tau[1]), P.Ny, bad);
    }
    double Nx[4], Dx[4], Ny[4], Dy[4];
    basis4(p.tau[0], Nx, Dx);
    basis4(p.tau[1], Ny, Dy);
    const double *row = lev + (p.cell[1] * P.sx + p.cell[0]);
    double Sx = 0, Sy = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const double c0 = ld<STAGED>(row), c1 = ld<STAGED>(row + 1), c2 = ld<STAGED>(row + 2), c3 = ld<STAGED>(row + 3);
        const double pv = fma(c3, Nx[3], fma(c2, Nx[2], fma(c1, Nx[1], c0 * Nx[0])));
        const double qv = fma(c3, Dx[3], fma(c2, Dx[2], fma(c1, Dx[1], c0 * Dx[0])));
        if (b == 0) { Sx = Ny[0] * qv; Sy = Dy[0] * pv; }
        else { Sx = fma(Ny[b], qv, Sx); Sy = fma(Dy[b], pv, Sy); }
        row += P.sx;
    }
    const double h = KIND == FULL ? 1.0 : 0.5;
    p.vel[0] = fma(h * P.gx, Sx, p.vel[0]);
    p.vel[1] = fma(h * P.gy, Sy, p.vel[1]);
}

template <int KIND, bool STAGED, bool POW2>
__device__ __forceinline__ void step3d(Point<3> &p, const double *lev, const BtParams &P, unsigned &bad)
{
    if constexpr (KIND != FIRST) {
        relocate<POW2>(p.tau[0], p.cell[0], fma(P.ncx, p.vel[0], p.tau[0]), P.Nx, bad);
        relocate<POW2>(p.tau[1], p.cell[1], fma(P.ncy, p.vel[1], p.tau[1]), P.Ny, bad);
        relocate<POW2>(p.tau[2], p.cell[2], fma(P.ncz, p.vel[2], p.tau[2]), P.Nz, bad);
    }
    double Nx[4], Dx[4], Ny[4], Dy[4], Nz[4], Dz[4];
    basis4(p.tau[0], Nx, Dx);
    basis4(p.tau[1], Ny, Dy);
    basis4(p.tau[2], Nz, Dz);
    const double *plane = lev + (p.cell[2] * P.sxy + p.cell[1] * P.sx + p.cell[0]);
    double Sx = 0, Sy = 0, Sz = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const double *row = plane;
        double r = 0, s = 0, w = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const double c0 = ld<STAGED>(row), c1 = ld<STAGED>(row + 1), c2 = ld<STAGED>(row + 2), c3 = ld<STAGED>(row + 3);
            const double pv = fma(c3, Nx[3], fma(c2, Nx[2], fma(c1, Nx[1], c0 * Nx[0])));
            const double qv = fma(c3, Dx[3], fma(c2, Dx[2], fma(c1, Dx[1], c0 * Dx[0])));
            if (b == 0) { r = Ny[0] * qv; s = Dy[0] * pv; w = Ny[0] * pv; }
            else { r = fma(Ny[b], qv, r); s = fma(Dy[b], pv, s); w = fma(Ny[b], pv, w); }
            row += P.sx;
        }
        if (c == 0) { Sx = Nz[0] * r; Sy = Nz[0] * s; Sz = Dz[0] * w; }
        else { Sx = fma(Nz[c], r, Sx); Sy = fma(Nz[c], s, Sy); Sz = fma(Dz[c], w, Sz); }
        plane += P.sxy;
    }
    const double h = KIND == FULL ? 1.0 : 0.5;
    p.vel[0] = fma(h * P.gx, Sx, p.vel[0]);
    p.vel[1] = fma(h * P.gy, Sy, p.vel[1]);
    p.vel[2] = fma(h * P.gz, Sz, p.vel[2]);
}

// ---- "xpp" level format (tail.cu: row_poly): per (row, cell) the cubic A(tau_x) = sum_a c_a 6 N_a, stored as two
// double2 halves [a0 a1] and [a2 a3] (row stride P.sx double2, the second half Nx further).  The x-contraction of a
// window row becomes two Horner evaluations (value: 3 FMA, derivative A' = a1 + 2 tau (a2 + 1.5 tau a3): 2 FMA) instead of
// eight FMAs plus the x basis; loads are 2 x 128-bit per row, conflict-free for consecutive cells.  P.gx carries the 1/3
// of A' = 3 sum_a c_a 2 N'_a.
template <bool STAGED> __device__ __forceinline__ double2 ld2(const double2 *p)
{
    if constexpr (STAGED) return *p;
    else return __ldg(p);
}

template <int KIND, bool STAGED, bool POW2>
__device__ __forceinline__ void step2d_xpp(Point<2> &p, const double *lev, const BtParams &P, unsigned &bad)
{
    if constexpr (KIND != FIRST) {
        relocate<POW2>(p.tau[0], p.cell[0], fma(P.ncx, p.vel[0], p.tau[0]), P.Nx, bad);
        relocate<POW2>(p.tau[1], p.cell[1], fma(P.ncy, p.vel[1], p.tau[1]), P.Ny, bad);
    }
    double Ny[4], Dy[4];
    basis4(p.tau[1], Ny, Dy);
    const double tx = p.tau[0], ta = 1.5 * tx, tb = 2.0 * tx;
    const double2 *row = reinterpret_cast<const double2 *>(lev) + (p.cell[1] * P.sx + p.cell[0]);
    double Sx = 0, Sy = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const double2 q01 = ld2<STAGED>(row), q23 = ld2<STAGED>(row + P.Nx);
        const double qv = fma(tb, fma(ta, q23.y, q23.x), q01.y);
        const double pv = fma(tx, fma(tx, fma(tx, q23.y, q23.x), q01.y), q01.x);
        if (b == 0) { Sx = Ny[0] * qv; Sy = Dy[0] * pv; }
        else { Sx = fma(Ny[b], qv, Sx); Sy = fma(Dy[b], pv, Sy); }
        row += P.sx;
    }
    const double h = KIND == FULL ? 1.0 : 0.5;
    p.vel[0] = fma(h * P.gx, Sx, p.vel[0]);
    p.vel[1] = fma(h * P.gy, Sy, p.vel[1]);
}

template <int KIND, bool STAGED, bool POW2>
__device__ __forceinline__ void step3d_xpp(Point<3> &p, const double *lev, const BtParams &P, unsigned &bad)
{
    if constexpr (KIND != FIRST) {
        relocate<POW2>(p.tau[0], p.cell[0], fma(P.ncx, p.vel[0], p.tau[0]), P.Nx, bad);
        relocate<POW2>(p.tau[1], p.cell[1], fma(P.ncy, p.vel[1], p.tau[1]), P.Ny, bad);
        relocate<POW2>(p.tau[2], p.cell[2], fma(P.ncz, p.vel[2], p.tau[2]), P.Nz, bad);
    }
    double Ny[4], Dy[4], Nz[4], Dz[4];
    basis4(p.tau[1], Ny, Dy);
    basis4(p.tau[2], Nz, Dz);
    const double tx = p.tau[0], ta = 1.5 * tx, tb = 2.0 * tx;
    const double2 *plane = reinterpret_cast<const double2 *>(lev) + (p.cell[2] * P.sxy + p.cell[1] * P.sx + p.cell[0]);
    double Sx = 0, Sy = 0, Sz = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const double2 *row = plane;
        double r = 0, s = 0, w = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const double2 q01 = ld2<STAGED>(row), q23 = ld2<STAGED>(row + P.Nx);
            const double qv = fma(tb, fma(ta, q23.y, q23.x), q01.y);
            const double pv = fma(tx, fma(tx, fma(tx, q23.y, q23.x), q01.y), q01.x);
            if (b == 0) { r = Ny[0] * qv; s = Dy[0] * pv; w = Ny[0] * pv; }
            else { r = fma(Ny[b], qv, r); s = fma(Dy[b], pv, s); w = fma(Ny[b], pv, w); }
            row += P.sx;
        }
        if (c == 0) { Sx = Nz[0] * r; Sy = Nz[0] * s; Sz = Dz[0] * w; }
        else { Sx = fma(Nz[c], r, Sx); Sy = fma(Nz[c], s, Sy); Sz = fma(Dz[c], w, Sz); }
        plane += P.sxy;
    }
    const double h = KIND == FULL ? 1.0 : 0.5;
    p.vel[0] = fma(h * P.gx, Sx, p.vel[0]);
    p.vel[1] = fma(h * P.gy, Sy, p.vel[1]);
    p.vel[2] = fma(h * P.gz, Sz, p.vel[2]);
}

template <int DIM, int KIND, bool STAGED, bool POW2, bool XPP>
__device__ __forceinline__ void step(Point<DIM> &p, const double *lev, const BtParams &P, unsigned &bad)
{
    if constexpr (DIM == 1) step1d<KIND, STAGED, POW2>(p, lev, P, bad);
    else if constexpr (DIM == 2) {
        if constexpr (XPP) step2d_xpp<KIND, STAGED, POW2>(p, lev, P, bad);
        else step2d<KIND, STAGED, POW2>(p, lev, P, bad);
    } else {
        if constexpr (XPP) step3d_xpp<KIND, STAGED, POW2>(p, lev, P, bad);
        else step3d<KIND, STAGED, POW2>(p, lev, P, bad);
    }
}

// Robust (slow) trace of one point straight from the global history: used only for points whose fast trace
// flagged a multi-period jump (wrap_cell).  Same arithmetic, cell index reduced with a true modulo; a full
// kick is applied as two half kicks from the same position.
template <int DIM, bool XPP> __device__ __noinline__ void slow_trace(Point<DIM> &p, const BtParams &P)
{
    const int Ns[3] = {P.Nx, P.Ny, P.Nz};
    const double nc[3] = {P.ncx, P.ncy, P.ncz};
    for (int m = P.first_level; m >= 0; --m) {
        const double *lev = P.hist + static_cast<size_t>(m) * (P.level_bytes / 8);
        const bool first = P.metrics && m == P.first_level;
        if (!first) {
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                const double t2 = fma(nc[d], p.vel[d], p.tau[d]);
                const double r = rint(t2);
                p.tau[d] = t2 - r;
                long long k = static_cast<long long>(p.cell[d]) + static_cast<long long>(r);
                k %= Ns[d];
                if (k < 0) k += Ns[d];
                p.cell[d] = static_cast<int>(k);
            }
        }
        unsigned bad = 0;
        step<DIM, FIRST, false, false, XPP>(p, lev, P, bad);
        if (!(first || m == 0)) step<DIM, FIRST, false, false, XPP>(p, lev, P, bad);
    }
}

// ---------------------------------------------------------------- mbarrier / bulk-copy primitives (PTX)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// the same on precomputed 32-bit shared addresses (hot loop: no generic->shared conversion per use)
__device__ __forceinline__ void mbar_arrive_a(unsigned bar)
{
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// named barrier over the consumer warps only (the producer warp never joins)
__device__ __forceinline__ void consumer_sync(unsigned threads) { asm volatile("bar.sync 1, %0;" ::"r"(threads) : "memory"); }

// peer-memory primitive: the exchange flags are raised with fence.sys + relaxed system-scope store, read with ld.acquire.sys;
// after a fence.sys by the same thread a relaxed system-scope store completes the release pattern (PTX memory model)
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

constexpr int kMaxStages = 8;
constexpr unsigned kBarBytes = 2 * kMaxStages * 8; // full[8], empty[8]
constexpr unsigned kRedBytes = 32 * 32 * 8;        // consumer-warp reduction scratch [32 warps][32 lanes]
constexpr unsigned kSmemFixed = kBarBytes + kRedBytes;

#ifndef NUFI_3D_MT
#define NUFI_3D_MT 512 // thread bound of the 3d B-spline kernel, one point per thread (128 registers)
#endif
#ifndef NUFI_3D_XPP_MT
#define NUFI_3D_XPP_MT 640
#endif
template <int DIM, int ILP, bool XPP> struct Tune
{
    // thread-count upper bound handed to __launch_bounds__ (sets the register budget); the xpp steps need fewer registers
    static constexpr int max_threads =
        DIM == 1 ? 1024 : (DIM == 2 ? (ILP == 1 ? 768 : 512) : (ILP == 1 ? (XPP ? NUFI_3D_XPP_MT : NUFI_3D_MT) : 256));
};

// ---------------------------------------------------------------- the kernel
template <int DIM, int ILP, bool STAGED, bool POW2, bool XPP>
__global__ void __launch_bounds__(Tune<DIM, ILP, XPP>::max_threads, 1)
    backtrace_kernel(const __grid_constant__ BtParams P, const __grid_constant__ EpilogueParams E)
{
    __shared__ unsigned int s_ticket;
    __shared__ unsigned short s_tfirst[256]; // epilogue: first tile of every CTA (filled below, read by the last CTA only)
    pdl_trigger(); // the slot reduction / field tail behind this launch may be scheduled as soon as an SM has room
    if (E.mode) // visible to the epilogue through the CTA's barriers (the producer warp's share through the __syncthreads below)
        for (unsigned b = threadIdx.x; b < E.n_active && b < 256; b += blockDim.x)
            s_tfirst[b] = static_cast<unsigned short>((b * E.F.rpc) / E.F.rpt);
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long *full = reinterpret_cast<unsigned long long *>(smem);
    unsigned long long *empty = full + kMaxStages;
    double(*sred)[32] = reinterpret_cast<double(*)[32]>(smem + kBarBytes);
    const unsigned char *ring = smem + kSmemFixed;

    const int lane = threadIdx.x & 31;
    const unsigned warp = threadIdx.x >> 5;
    const unsigned W = P.W;
    // this CTA's run of CTA-rounds
    const unsigned g0 = blockIdx.x * P.rpc;
    if (g0 >= P.R) return; // whole CTA idle (uniform)
    const unsigned my_rounds = min(P.rpc, P.R - g0);
    const unsigned t_first = g0 / P.rpt;
    // Chunks of the history, newest first, aligned at the TOP: chunk i holds levels [first_level - (i+1) Lc + 1, first_level - i Lc],
    // so every chunk but the last (which ends at level 0) has exactly Lc levels and the per-chunk code has no ragged cases.
    const int n_levels = P.first_level + 1;
    const int n_chunks = n_levels > 0 ? (n_levels + P.Lc - 1) / P.Lc : 0;
    const int rem_levels = n_levels - (n_chunks - 1) * P.Lc; // levels in the bottom chunk, 1..Lc
    const unsigned level_doubles = P.level_bytes / 8;

    if constexpr (STAGED) {
        if (threadIdx.x == 0) {
            for (int s = 0; s < P.stages; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], W);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }

    if (STAGED && warp == W) {
        // ------------------------------------------------ producer warp: stream chunks newest -> oldest, once per round
        if (lane == 0) {
            int s = 0;
            unsigned ph = 0;
            bool primed = false;
            for (unsigned r = 0; r < my_rounds; ++r)
                for (int ci = 0; ci < n_chunks; ++ci) {
                    if (primed) mbar_wait(&empty[s], ph ^ 1u);
                    const bool bottom = ci == n_chunks - 1;
                    const int cnt = bottom ? rem_levels : P.Lc;
                    const int lv_lo = bottom ? 0 : P.first_level - (ci + 1) * P.Lc + 1;
                    const unsigned bytes = static_cast<unsigned>(cnt) * P.level_bytes;
                    mbar_expect_tx(&full[s], bytes);
                    const unsigned char *src = reinterpret_cast<const unsigned char *>(P.hist) + static_cast<size_t>(lv_lo) * P.level_bytes;
                    unsigned char *dst = const_cast<unsigned char *>(ring) + static_cast<size_t>(s) * P.stage_bytes;
                    for (unsigned off = 0; off < bytes; off += 32768u)
                        bulk_g2s(dst + off, src + off, min(32768u, bytes - off), &full[s]);
                    if (++s == P.stages) { s = 0; ph ^= 1u; primed = true; }
                }
        }
        return;
    }

    // ---------------------------------------------------- consumer warps
    double acc = 0;
    double m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    unsigned cur_tile = t_first;
    int s = 0;
    unsigned ph = 0;

    auto flush_tile = [&](unsigned tile) { // CTA-uniform: every consumer warp calls it
        sred[warp][lane] = acc;
        consumer_sync(W * 32);
        if (warp == 0) {
            double sum = 0;
            for (unsigned w = 0; w < W; ++w) sum += sred[w][lane];
            // TN < 32: the 32/TN lanes that traced the same node (lane % TN) are combined by a fixed shuffle tree
            for (unsigned off = 16; off >= P.TN; off >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, off);
            P.slots[(static_cast<size_t>(blockIdx.x) * P.Tmax + (tile - t_first)) * 32 + lane] = static_cast<unsigned>(lane) < P.TN ? sum : 0.0;
        }
        consumer_sync(W * 32);
        acc = 0;
    };

    for (unsigned r = 0; r < my_rounds; ++r) {
        const unsigned g = g0 + r;
        const unsigned tile = g / P.rpt;
        const unsigned jr = g - tile * P.rpt;
        if (tile != cur_tile) {
            if (!P.metrics) flush_tile(cur_tile);
            cur_tile = tile;
        }
        // warp-unit of this warp: interleaved (default) = the warps of a round and the points of a thread are spread evenly
        // over the velocity range, so every CTA sees the same mix of fast/trapped orbits (equal bank-conflict load)
        const unsigned jc = P.interleave ? jr + warp * P.rpt : jr * W + warp;
        if (jc >= P.upt) { // no unit for this warp in this round: keep the stage protocol going
            if constexpr (STAGED) {
                for (int ci = 0; ci < n_chunks; ++ci) {
                    mbar_wait(&full[s], ph);
                    if (lane == 0) mbar_arrive(&empty[s]);
                    if (++s == P.stages) { s = 0; ph ^= 1u; }
                }
            }
            continue;
        }

        unsigned long long l = P.l_first + static_cast<unsigned long long>(tile) * P.TN + (lane & (P.TN - 1));
        const unsigned vsub = static_cast<unsigned>(lane) >> P.TNlog2, G = 32u >> P.TNlog2; // velocity sub-index within the warp-unit
        const bool node_ok = l <= P.l_last;
        if (!node_ok) l = P.l_first;
        int ix, iy = 0, iz = 0;
        {
            unsigned long long t = l;
            ix = static_cast<int>(t % P.Nx);
            t /= P.Nx;
            if (DIM >= 2) { iy = static_cast<int>(t % P.Ny); t /= P.Ny; }
            if (DIM >= 3) iz = static_cast<int>(t);
        }

        Point<DIM> pt[ILP];
        bool ok[ILP];
        unsigned bad[ILP];
        double v0[ILP][DIM]; // starting velocities (metrics need them; the slow path restarts from them)
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            unsigned long long j = P.interleave ? static_cast<unsigned long long>(jc) + static_cast<unsigned long long>(i) * P.upt
                                                : static_cast<unsigned long long>(jc) * ILP + i;
            j = j * G + vsub; // the 32/TN lanes of a node take neighbouring velocities: early in the history they share cells
            ok[i] = node_ok && j < P.Nvel_loc;
            if (j >= P.Nvel_loc) j = 0;
            j = j * P.vstride + P.voff; // this GPU's share of the velocity nodes (multi-GPU step: every vstride-th one)
            const unsigned long long q = l * P.Nvel + j;
            ok[i] = ok[i] && q >= P.q_begin && q < P.q_end;
            const int iu = static_cast<int>(j % P.Nu);
            const int iv = DIM >= 2 ? static_cast<int>((j / P.Nu) % P.Nv) : 0;
            const int iw = DIM >= 3 ? static_cast<int>(j / (static_cast<unsigned long long>(P.Nu) * P.Nv)) : 0;
            bad[i] = 0;
            // node ix sits on the left edge of cell ix: xi = ix  ->  tau = -1/2
            pt[i].cell[0] = ix;
            pt[i].tau[0] = -0.5;
            pt[i].vel[0] = P.metrics ? P.ug0 + iu * P.dug : P.u0 + iu * P.du;
            if constexpr (DIM >= 2) {
                pt[i].cell[1] = iy;
                pt[i].tau[1] = -0.5;
                pt[i].vel[1] = P.metrics ? P.vg0 + iv * P.dvg : P.v0 + iv * P.dv;
            }
            if constexpr (DIM >= 3) {
                pt[i].cell[2] = iz;
                pt[i].tau[2] = -0.5;
                pt[i].vel[2] = P.metrics ? P.wg0 + iw * P.dwg : P.w0 + iw * P.dw;
            }
#pragma unroll
            for (int dd = 0; dd < DIM; ++dd) v0[i][dd] = pt[i].vel[dd];
        }

        // ---- the trace: chunks newest -> oldest; inside a chunk levels top -> bottom
        {
            const unsigned full0 = smem_u32(full), empty0 = smem_u32(empty);
            const int Lc = P.Lc;
            for (int ci = 0; ci < n_chunks; ++ci) {
                const bool bottom = ci == n_chunks - 1;
                int cnt = bottom ? rem_levels : Lc; // levels in this chunk
                const double *base;
                if constexpr (STAGED) {
                    mbar_wait_a(full0 + 8u * s, ph);
                    base = reinterpret_cast<const double *>(ring + static_cast<size_t>(s) * P.stage_bytes);
                } else {
                    base = P.hist + static_cast<size_t>(bottom ? 0 : P.first_level - (ci + 1) * Lc + 1) * level_doubles;
                }
                const double *lev = base + static_cast<size_t>(cnt - 1) * level_doubles;
                if (P.metrics && ci == 0) { // eval_f: half kick on level n at the starting position
#pragma unroll
                    for (int i = 0; i < ILP; ++i) step<DIM, FIRST, STAGED, POW2, XPP>(pt[i], lev, P, bad[i]);
                    --cnt;
                    lev -= level_doubles;
                }
                if (bottom) --cnt; // level 0 takes the half kick below
                for (int k = cnt >> 1; k > 0; --k) { // full-kick steps, two levels per trip
#pragma unroll
                    for (int i = 0; i < ILP; ++i) step<DIM, FULL, STAGED, POW2, XPP>(pt[i], lev, P, bad[i]);
#pragma unroll
                    for (int i = 0; i < ILP; ++i) step<DIM, FULL, STAGED, POW2, XPP>(pt[i], lev - level_doubles, P, bad[i]);
                    lev -= 2 * level_doubles;
                }
                if (cnt & 1) {
#pragma unroll
                    for (int i = 0; i < ILP; ++i) step<DIM, FULL, STAGED, POW2, XPP>(pt[i], lev, P, bad[i]);
                }
                if (bottom) { // level 0: drift + half kick
#pragma unroll
                    for (int i = 0; i < ILP; ++i) step<DIM, LAST, STAGED, POW2, XPP>(pt[i], base, P, bad[i]);
                }
                if constexpr (STAGED) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive_a(empty0 + 8u * s);
                    if (++s == P.stages) { s = 0; ph ^= 1u; }
                }
            }
        }

#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (!POW2 && bad[i]) { // a multi-period jump was clamped: redo this point on the robust path
#pragma unroll
                for (int dd = 0; dd < DIM; ++dd) { pt[i].vel[dd] = v0[i][dd]; pt[i].tau[dd] = -0.5; }
                pt[i].cell[0] = ix;
                if constexpr (DIM >= 2) pt[i].cell[1] = iy;
                if constexpr (DIM >= 3) pt[i].cell[2] = iz;
                slow_trace<DIM, XPP>(pt[i], P);
            }
            // foot of the characteristic in physical coordinates (periodic image inside the box)
            const double x = P.x_min + (pt[i].cell[0] + (0.5 + pt[i].tau[0])) * P.dx;
            double f;
            if constexpr (DIM == 1) f = f0_1d(P, x, pt[i].vel[0]);
            else if constexpr (DIM == 2) {
                const double y = P.y_min + (pt[i].cell[1] + (0.5 + pt[i].tau[1])) * P.dy;
                f = f0_2d(P, x, y, pt[i].vel[0], pt[i].vel[1]);
            } else {
                const double y = P.y_min + (pt[i].cell[1] + (0.5 + pt[i].tau[1])) * P.dy;
                const double z = P.z_min + (pt[i].cell[2] + (0.5 + pt[i].tau[2])) * P.dz;
                f = f0_3d(P, x, y, z, pt[i].vel[0], pt[i].vel[1], pt[i].vel[2]);
            }
            if (ok[i]) {
                acc += f;
                if (P.metrics) { // nufi/cuda_kernel.cu:72-78, 264-270, 459-465
                    double vsq = v0[i][0] * v0[i][0];
                    if constexpr (DIM >= 2) vsq += v0[i][1] * v0[i][1];
                    if constexpr (DIM >= 3) vsq += v0[i][2] * v0[i][2];
                    m0 += P.mweight * f;
                    m1 += P.mweight * f * f;
                    m2 += DIM == 1 ? P.mweight * (vsq * f / 2) : P.mweight * vsq * f / 2;
                    m3 += (f > 0) ? -P.mweight * f * log(f) : 0;
                }
            }
        }
    }

    if (!P.metrics) {
        flush_tile(cur_tile);
        if (E.mode) { // ---- epilogue: the last CTA to arrive reduces the slots of all tiles (and pushes them to the peers)
            // (flush_tile ended with a barrier over the consumer warps: thread 0 has observed every slot store of this CTA, so its
            //  one cumulative fence orders them all before the ticket -- the pattern of cooperative-groups grid sync.  A fence in
            //  every thread costs microseconds here, and tens of them at system scope below.)
            if (threadIdx.x == 0) {
                __threadfence();
                s_ticket = atomicAdd(E.done, 1u);
                __threadfence();
            }
            consumer_sync(W * 32);
            if (s_ticket == E.n_active - 1) {
                const FinishParams &F = E.F;
                // Batches of 4 tiles = 32 (tile, w) tasks spread over the consumer warps: task (tile, w) adds the slots of CTAs
                // b_lo+w, b_lo+w+8, ... (loads issued four at a time, so a task costs about one L2 round trip); then one warp per
                // tile adds the 8 partial sums in order -- the association of finish_rho_kernel.
                for (unsigned t0 = 0; t0 < F.n_tiles; t0 += 4) {
                    for (unsigned task = warp; task < 32; task += W) {
                        const unsigned tile = t0 + (task >> 3), w = task & 7;
                        double sum = 0;
                        if (tile < F.n_tiles) {
                            const unsigned b_lo = (tile * F.rpt) / F.rpc;
                            const unsigned b_hi = ((tile + 1) * F.rpt - 1) / F.rpc;
                            for (unsigned b = b_lo + w; b <= b_hi; b += 32) {
                                double v[4];
#pragma unroll
                                for (unsigned u = 0; u < 4; ++u) {
                                    const unsigned bb = b + 8 * u;
                                    v[u] = 0.0;
                                    if (bb <= b_hi) { // first tile of CTA bb: tabulated at kernel start (no division per load)
                                        const unsigned t_first = bb < 256 ? s_tfirst[bb] : (bb * F.rpc) / F.rpt;
                                        v[u] = __ldcg(F.slots + (static_cast<size_t>(bb) * F.Tmax + (tile - t_first)) * 32 + lane);
                                    }
                                }
                                sum = (((sum + v[0]) + v[1]) + v[2]) + v[3]; // + 0.0 is exact: same order as one-by-one
                            }
                        }
                        sred[task][lane] = sum;
                    }
                    consumer_sync(W * 32);
                    for (unsigned k = warp; k < 4; k += W) {
                        const unsigned tile = t0 + k;
                        if (tile >= F.n_tiles) continue;
                        double tot = 0;
#pragma unroll
                        for (int w = 0; w < 8; ++w) tot += sred[8 * k + w][lane];
                        const unsigned long long l = F.l_first + static_cast<unsigned long long>(tile) * F.TN + lane;
                        if (static_cast<unsigned>(lane) < F.TN && l <= F.l_last) {
                            const double val = -F.dV * tot;
                            F.rho_partial[l] = val;
                            if (F.rho_full) F.rho_full[l] = 1 - F.dV * tot;
                            if (E.mode == 2)
                                for (int p = 0; p < E.X.world; ++p) E.X.data[p][l] = val; // NVLink stores into every GPU's buffer
                        }
                    }
                    consumer_sync(W * 32);
                }
                if (E.mode == 2) { // the batch loop ended with a barrier: lanes 0..world-1 of warp 0 have observed all remote stores
                    if (warp == 0) {
                        __threadfence_system(); // ONE system-scope fence (warp 0), then the flags go out to all peers in parallel
                        if (lane < E.X.world) st_relaxed_sys(E.X.flag[lane], E.X.epoch);
                    }
                }
                if (threadIdx.x == 0) *E.done = 0; // every participant has arrived: ready for the next launch
            }
        }
    } else { // deterministic block reduction of the four metric sums
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m0 += __shfl_down_sync(0xffffffffu, m0, o);
            m1 += __shfl_down_sync(0xffffffffu, m1, o);
            m2 += __shfl_down_sync(0xffffffffu, m2, o);
            m3 += __shfl_down_sync(0xffffffffu, m3, o);
        }
        if (lane == 0) { sred[warp][0] = m0; sred[warp][1] = m1; sred[warp][2] = m2; sred[warp][3] = m3; }
        consumer_sync(W * 32);
        if (threadIdx.x < 4) {
            double sum = 0;
            for (unsigned w = 0; w < W; ++w) sum += sred[w][threadIdx.x];
            P.mpartials[blockIdx.x * 4 + threadIdx.x] = sum;
        }
    }
}

// Adds the per-(CTA, tile) slots of each tile in a fixed order.  One block (8 warps) per tile of 32 nodes.
__global__ void __launch_bounds__(256) finish_rho_kernel(const __grid_constant__ FinishParams F)
{
    __shared__ double part[8][32];
    pdl_wait();
    pdl_trigger();
    const unsigned tile = blockIdx.x;
    const int lane = threadIdx.x & 31;
    const unsigned wj = threadIdx.x >> 5;
    const unsigned b_lo = (tile * F.rpt) / F.rpc;
    const unsigned b_hi = ((tile + 1) * F.rpt - 1) / F.rpc;
    double sum = 0;
    for (unsigned b = b_lo + wj; b <= b_hi; b += 8) {
        const unsigned t_first = (b * F.rpc) / F.rpt;
        sum += F.slots[(static_cast<size_t>(b) * F.Tmax + (tile - t_first)) * 32 + lane];
    }
    part[wj][lane] = sum;
    __syncthreads();
    if (wj == 0) {
        double tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) tot += part[w][lane];
        const unsigned long long l = F.l_first + static_cast<unsigned long long>(tile) * F.TN + lane;
        if (static_cast<unsigned>(lane) < F.TN && l <= F.l_last) {
            F.rho_partial[l] = -F.dV * tot;
            if (F.rho_full) F.rho_full[l] = 1 - F.dV * tot;
        }
    }
}

// finish_rho_kernel fused with the peer exchange (multi-GPU step): the tile's rho values are stored straight into the
// exchange buffer of EVERY GPU (own one included; remote stores travel over NVLink), and the last block to finish releases
// this rank's flag on every GPU.  grid = max(n_tiles, 1) blocks: a rank without work only raises its flags.
__global__ void __launch_bounds__(256) finish_push_kernel(const __grid_constant__ FinishParams F, const __grid_constant__ PeerPush X)
{
    __shared__ double part[8][32];
    pdl_wait();
    pdl_trigger();
    const unsigned tile = blockIdx.x;
    const int lane = threadIdx.x & 31;
    const unsigned wj = threadIdx.x >> 5;
    if (tile < F.n_tiles) {
        const unsigned b_lo = (tile * F.rpt) / F.rpc;
        const unsigned b_hi = ((tile + 1) * F.rpt - 1) / F.rpc;
        double sum = 0;
        for (unsigned b = b_lo + wj; b <= b_hi; b += 8) {
            const unsigned t_first = (b * F.rpc) / F.rpt;
            sum += F.slots[(static_cast<size_t>(b) * F.Tmax + (tile - t_first)) * 32 + lane];
        }
        part[wj][lane] = sum;
        __syncthreads();
        if (wj == 0) {
            double tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) tot += part[w][lane];
            const unsigned long long l = F.l_first + static_cast<unsigned long long>(tile) * F.TN + lane;
            if (static_cast<unsigned>(lane) < F.TN && l <= F.l_last) {
                const double val = -F.dV * tot;
                F.rho_partial[l] = val;
                if (F.rho_full) F.rho_full[l] = 1 - F.dV * tot;
                for (int p = 0; p < X.world; ++p) X.data[p][l] = val;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system(); // cumulative: this block's remote stores (observed through the barrier) before its ticket
        const unsigned t = atomicAdd(X.ticket, 1u);
        if (t == gridDim.x - 1) { // every block has stored and fenced
            *X.ticket = 0;
            __threadfence_system();
            for (int p = 0; p < X.world; ++p) st_relaxed_sys(X.flag[p], X.epoch);
        }
    }
}

__global__ void finish_metrics_kernel(const double *mpartials, unsigned grid, double *metrics)
{
    if (threadIdx.x < 4) {
        double sum = 0;
        for (unsigned c = 0; c < grid; ++c) sum += mpartials[c * 4 + threadIdx.x];
        metrics[threadIdx.x] = sum;
    }
}

// ---------------------------------------------------------------- sampling at arbitrary points (plots, diagnostics)
// physical coordinate -> (cell, centred offset), the reference's wrap/locate arithmetic (nufi/fields.hpp:315-331)
__device__ __forceinline__ void locate(double x, double x_min, double L, double L_inv, double dx_inv, int N, int &k, double &tau)
{
    x -= x_min;
    x -= L * floor(x * L_inv);
    const double kf = floor(x * dx_inv);
    k = static_cast<int>(kf);
    tau = (x * dx_inv - kf) - 0.5;
    if (k >= N) { k -= N; } // x rounded up to exactly L
    if (k < 0) k = 0;
}

struct SampleParams
{
    BtParams P;
    double Lx, Ly, Lz, Lx_inv, Ly_inv, Lz_inv, dx_inv, dy_inv, dz_inv;
    const double *pts; // [npts][2*dim]: x.., v..
    double *out;
    size_t npts;
    int with_first_half_kick; // 1: eval_f (nufi/rho.hpp:63-96, 234-281, 369-426), 0: eval_ftilda
    int feet;                 // 1: write the foot (x.., v..) of the characteristic instead of f0 there (eval_phase_flow, rho.hpp:98-131)
};

// f(t_n, x, v) at arbitrary phase-space points: one thread per point, history read from global memory.
template <int DIM, bool XPP> __global__ void sample_f_kernel(const __grid_constant__ SampleParams S)
{
    const BtParams &P = S.P;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < S.npts; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const double *q = S.pts + i * 2 * DIM;
        Point<DIM> p;
        locate(q[0], P.x_min, S.Lx, S.Lx_inv, S.dx_inv, P.Nx, p.cell[0], p.tau[0]);
        if constexpr (DIM >= 2) locate(q[1], P.y_min, S.Ly, S.Ly_inv, S.dy_inv, P.Ny, p.cell[1], p.tau[1]);
        if constexpr (DIM >= 3) locate(q[2], P.z_min, S.Lz, S.Lz_inv, S.dz_inv, P.Nz, p.cell[2], p.tau[2]);
#pragma unroll
        for (int d = 0; d < DIM; ++d) p.vel[d] = q[DIM + d];
        if (P.first_level >= 0) slow_trace<DIM, XPP>(p, P); // robust path: true modulo wrap, any jump length
        const double x = P.x_min + (p.cell[0] + (0.5 + p.tau[0])) * P.dx;
        if (S.feet) { // the flow map itself; positions reduced with L*floor(x*L_inv) as the reference does (no x_min shift)
            double *o = S.out + i * 2 * DIM;
            o[0] = x - S.Lx * floor(x * S.Lx_inv);
            if constexpr (DIM >= 2) {
                const double y = P.y_min + (p.cell[1] + (0.5 + p.tau[1])) * P.dy;
                o[1] = y - S.Ly * floor(y * S.Ly_inv);
            }
            if constexpr (DIM >= 3) {
                const double z = P.z_min + (p.cell[2] + (0.5 + p.tau[2])) * P.dz;
                o[2] = z - S.Lz * floor(z * S.Lz_inv);
            }
#pragma unroll
            for (int d = 0; d < DIM; ++d) o[DIM + d] = p.vel[d];
            continue;
        }
        double f;
        if constexpr (DIM == 1) f = f0_1d(P, x, p.vel[0]);
        else if constexpr (DIM == 2) f = f0_2d(P, x, P.y_min + (p.cell[1] + (0.5 + p.tau[1])) * P.dy, p.vel[0], p.vel[1]);
        else
            f = f0_3d(P, x, P.y_min + (p.cell[1] + (0.5 + p.tau[1])) * P.dy, P.z_min + (p.cell[2] + (0.5 + p.tau[2])) * P.dz, p.vel[0],
                      p.vel[1], p.vel[2]);
        S.out[i] = f;
    }
}

struct FieldSampleParams
{
    int dim, Nx, Ny, Nz, der; // der: -1 value, 0/1/2 first derivative along x/y/z
    double x_min, y_min, z_min, Lx, Ly, Lz, Lx_inv, Ly_inv, Lz_inv, dx_inv, dy_inv, dz_inv;
    const double *level; // reference-format level: halo, row stride Nx+3
    const double *pts;   // [npts][dim]
    double *out;
    size_t npts;
};

// phi_n or one first derivative at arbitrary points (nufi/fields.hpp eval<real,order,dx,dy,dz>).
__global__ void sample_field_kernel(const __grid_constant__ FieldSampleParams S)
{
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < S.npts; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const double *q = S.pts + i * S.dim;
        int k[3] = {0, 0, 0};
        double tau[3] = {0, 0, 0};
        locate(q[0], S.x_min, S.Lx, S.Lx_inv, S.dx_inv, S.Nx, k[0], tau[0]);
        if (S.dim >= 2) locate(q[1], S.y_min, S.Ly, S.Ly_inv, S.dy_inv, S.Ny, k[1], tau[1]);
        if (S.dim >= 3) locate(q[2], S.z_min, S.Lz, S.Lz_inv, S.dz_inv, S.Nz, k[2], tau[2]);
        double W[3][4]; // per dimension: basis values (N_a) or derivatives (N'_a * dx_inv)
        const double inv[3] = {S.dx_inv, S.dy_inv, S.dz_inv};
        for (int d = 0; d < 3; ++d) {
            double N[4], D[4];
            basis4(tau[d], N, D);
            for (int a = 0; a < 4; ++a) W[d][a] = d >= S.dim ? (a == 0 ? 1.0 : 0.0) : (S.der == d ? D[a] * 0.5 * inv[d] : N[a] * (1.0 / 6.0));
        }
        const int sy = S.Nx + 3, sz = sy * (S.Ny + 3);
        const int nb = S.dim >= 2 ? 4 : 1, nc = S.dim >= 3 ? 4 : 1;
        double r = 0;
        for (int c = 0; c < nc; ++c)
            for (int b = 0; b < nb; ++b) {
                const double *row = S.level + static_cast<size_t>(k[2] + c) * sz + static_cast<size_t>(k[1] + b) * sy + k[0];
                const double rv = fma(row[3], W[0][3], fma(row[2], W[0][2], fma(row[1], W[0][1], row[0] * W[0][0])));
                r = fma(rv, W[1][b] * W[2][c], r);
            }
        S.out[i] = r;
    }
}

template <int DIM, int ILP, bool STAGED, bool POW2, bool XPP>
cudaError_t launch_variant(const BtParams &P, const EpilogueParams &E, unsigned grid, unsigned threads, size_t smem_bytes, cudaStream_t st)
{
    auto kern = backtrace_kernel<DIM, ILP, STAGED, POW2, XPP>;
    if (smem_bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes));
        if (e != cudaSuccess) return e;
    }
    kern<<<grid, threads, smem_bytes, st>>>(P, E);
    return cudaGetLastError();
}

template <int DIM, int ILP, bool XPP>
cudaError_t launch_fmt(const BtParams &P, const EpilogueParams &E, bool staged, bool pow2, unsigned grid, unsigned threads, size_t smem_bytes, cudaStream_t st)
{
    if (staged) {
        return pow2 ? launch_variant<DIM, ILP, true, true, XPP>(P, E, grid, threads, smem_bytes, st)
                    : launch_variant<DIM, ILP, true, false, XPP>(P, E, grid, threads, smem_bytes, st);
    }
    return pow2 ? launch_variant<DIM, ILP, false, true, XPP>(P, E, grid, threads, smem_bytes, st)
                : launch_variant<DIM, ILP, false, false, XPP>(P, E, grid, threads, smem_bytes, st);
}

template <int DIM, int ILP>
cudaError_t launch_ilp(const BtParams &P, const EpilogueParams &E, bool xpp, bool staged, bool pow2, unsigned grid, unsigned threads, size_t smem_bytes, cudaStream_t st)
{
    if constexpr (DIM >= 2) {
        if (xpp) return launch_fmt<DIM, ILP, true>(P, E, staged, pow2, grid, threads, smem_bytes, st);
    }
    return launch_fmt<DIM, ILP, false>(P, E, staged, pow2, grid, threads, smem_bytes, st);
}

template <int DIM>
cudaError_t launch_dim(const BtParams &P, const EpilogueParams &E, int ilp, bool xpp, bool staged, bool pow2, unsigned grid, unsigned threads, size_t smem_bytes,
                       cudaStream_t st)
{
    return ilp == 2 ? launch_ilp<DIM, 2>(P, E, xpp, staged, pow2, grid, threads, smem_bytes, st)
                    : launch_ilp<DIM, 1>(P, E, xpp, staged, pow2, grid, threads, smem_bytes, st);
}

int max_threads_for(int dim, int ilp, bool xpp)
{
    if (dim == 1) return ilp == 1 ? Tune<1, 1, false>::max_threads : Tune<1, 2, false>::max_threads;
    if (dim == 2) {
        if (xpp) return ilp == 1 ? Tune<2, 1, true>::max_threads : Tune<2, 2, true>::max_threads;
        return ilp == 1 ? Tune<2, 1, false>::max_threads : Tune<2, 2, false>::max_threads;
    }
    if (xpp) return ilp == 1 ? Tune<3, 1, true>::max_threads : Tune<3, 2, true>::max_threads;
    return ilp == 1 ? Tune<3, 1, false>::max_threads : Tune<3, 2, false>::max_threads;
}

bool is_pow2(size_t n) { return n && !(n & (n - 1)); }

int env_int(const char *name, int dflt)
{
    const char *v = std::getenv(name);
    return v && *v ? std::atoi(v) : dflt;
}

// chains (warps x points per thread) per SM beyond which a CTA-round's time grows with its width
int saturation_chains(int dim) { return dim == 1 ? 32 : (dim == 2 ? 40 : 24); }

} // namespace

// Host side: decomposition + launch.  q is the reference's flat quadrature index (cuda_kernel.cu:40-41,
// 219-225, 402-412); [q_begin,q_end) may cut through a node's velocity range (masked per point).
int launch_backtrace(Handle *h, size_t n, size_t q_begin, size_t q_end, bool metrics, bool defer_finish)
{
    const nufi_b200_config3d &c = h->c;
    BtParams P{};
    P.dim = h->dim;
    P.Nx = static_cast<int>(c.Nx); P.Ny = static_cast<int>(c.Ny); P.Nz = static_cast<int>(c.Nz);
    P.Nu = static_cast<int>(c.Nu); P.Nv = static_cast<int>(c.Nv); P.Nw = static_cast<int>(c.Nw);
    P.sx = h->sx;
    P.sxy = h->sxy;
    P.hist = h->d_hist;
    P.first_level = metrics ? (n == 0 ? -1 : static_cast<int>(n)) : static_cast<int>(n) - 1;
    P.metrics = metrics ? 1 : 0;
    P.ncx = -(c.dt * c.dx_inv); P.ncy = -(c.dt * c.dy_inv); P.ncz = -(c.dt * c.dz_inv);
    const double scale = h->dim == 2 ? 12.0 : 72.0;
    P.gx = -c.dt * c.dx_inv / scale / (h->xpp ? 3.0 : 1.0); P.gy = -c.dt * c.dy_inv / scale; P.gz = -c.dt * c.dz_inv / scale;
    P.x_min = c.x_min; P.y_min = c.y_min; P.z_min = c.z_min;
    P.dx = c.dx; P.dy = c.dy; P.dz = c.dz;
    // rho.hpp:136-137, 291-296, 441-447: du recomputed from the bounds, first node u_min + 0.5*du
    P.du = (c.u_max - c.u_min) / c.Nu; P.u0 = c.u_min + 0.5 * P.du;
    P.dv = (c.v_max - c.v_min) / c.Nv; P.v0 = c.v_min + 0.5 * P.dv;
    P.dw = (c.w_max - c.w_min) / c.Nw; P.w0 = c.w_min + 0.5 * P.dw;
    // cuda_kernel.cu:67, 259-260, 453-455: u_min + iu*du + du/2 with the stored conf.du
    P.dug = c.du; P.ug0 = c.u_min + c.du / 2;
    P.dvg = c.dv; P.vg0 = c.v_min + c.dv / 2;
    P.dwg = c.dw; P.wg0 = c.w_min + c.dw / 2;
    P.f0_kind = h->f0.kind;
    for (int i = 0; i < 4; ++i) P.f0p[i] = h->f0.p[i];
    // metric weights exactly as the reference writes them (cuda_kernel.cu:70, 262, 457)
    P.mweight = h->dim == 1 ? c.du * c.dx : (h->dim == 2 ? c.dx * c.dy * c.du * c.dv : c.du * c.dv * c.dw);

    P.Nvel = h->n_vel;
    P.vstride = h->vstride > 0 ? h->vstride : 1;
    P.voff = h->voff;
    P.Nvel_loc = P.voff < P.Nvel ? (P.Nvel - P.voff + P.vstride - 1) / P.vstride : 0;
    if (P.Nvel_loc == 0) return fail(h, NUFI_B200_ERR_ARG, "velocity share of this GPU is empty");
    P.q_begin = q_begin; P.q_end = q_end;
    P.l_first = q_begin / P.Nvel;
    P.l_last = (q_end - 1) / P.Nvel;
    const unsigned long long n_nodes_range = P.l_last - P.l_first + 1;
    // nodes per tile: 32 = every lane its own node (rigid drift, conflict-free at any depth); fewer = 32/TN lanes per node with
    // neighbouring velocities, whose window loads coincide early in the history and are served as broadcasts (short 3d histories)
    unsigned TN = 32;
    {
        const int want = env_int("NUFI_B200_TN", 0);
        if (want == 1 || want == 2 || want == 4 || want == 8 || want == 16 || want == 32) TN = static_cast<unsigned>(want);
        else if (h->tn_force) TN = static_cast<unsigned>(h->tn_force);
    }
    P.TN = TN;
    P.TNlog2 = 0;
    while ((1u << P.TNlog2) < TN) ++P.TNlog2;
    const unsigned G = 32u / TN;
    const unsigned long long n_tiles64 = (n_nodes_range + TN - 1) / TN;
    if (n_tiles64 >= (1ull << 31)) return fail(h, NUFI_B200_ERR_RANGE, "too many tiles for one launch");
    P.n_tiles = static_cast<unsigned>(n_tiles64);
    P.level_bytes = static_cast<unsigned>(h->level_stride * 8);

    // ---- variant: stage the history through shared memory when at least two levels fit
    const size_t ring_budget = h->smem_optin > kSmemFixed + 1024 ? h->smem_optin - kSmemFixed - 1024 : 0;
    bool staged = P.first_level >= 0 && 2ull * P.level_bytes <= ring_budget;
    if (h->variant_force == 1) staged = false;
    if (h->variant_force == 2 && 2ull * P.level_bytes > ring_budget)
        return fail(h, NUFI_B200_ERR_ARG, "staged variant forced but two levels do not fit in shared memory");
    const bool pow2 = is_pow2(c.Nx) && is_pow2(c.Ny) && is_pow2(c.Nz);
    const unsigned grid = static_cast<unsigned>(h->sm_count);

    // ---- chunking of the staged history
    P.Lc = 1; P.stages = 0; P.stage_bytes = P.level_bytes;
    if (staged) {
        // as many levels per chunk as a two-stage ring allows (cap 16): the per-chunk bookkeeping (barrier wait/arrive, pointer
        // set-up) is paid once per Lc levels -- measured C1 0.117 -> 0.096 ms (Lc 5 -> 16), C3 8.65 -> 7.86 ms (Lc 1 -> 3)
        int Lc = static_cast<int>(ring_budget / (2ull * P.level_bytes));
        Lc = env_int("NUFI_B200_LC", Lc);
        Lc = Lc < 1 ? 1 : (Lc > 16 ? 16 : Lc);
        while (Lc > 1 && 2ull * Lc * P.level_bytes > ring_budget) --Lc;
        int stages = static_cast<int>(ring_budget / (static_cast<size_t>(Lc) * P.level_bytes));
        if (stages > kMaxStages) stages = kMaxStages;
        const int n_chunks = P.first_level / Lc + 1;
        if (stages > n_chunks) stages = n_chunks > 2 ? n_chunks : 2;
        P.Lc = Lc; P.stages = stages; P.stage_bytes = static_cast<unsigned>(Lc) * P.level_bytes;
    } else if (P.first_level >= 0) {
        P.Lc = P.first_level + 1; // global variant: the whole history is one "chunk" read in place
    }

    // ---- shape: points per thread (ILP) and consumer warps per CTA (W); a CTA-round = W warp-units of one tile
    int best_ilp = 1;
    unsigned best_W = 1;
    {
        double best_cost = 1e300;
        const int csat = saturation_chains(h->dim);
        const int force_ilp = env_int("NUFI_B200_ILP", 0), force_w = env_int("NUFI_B200_W", 0);
        for (int ilp = 1; ilp <= 2; ++ilp) {
            if (force_ilp && ilp != force_ilp) continue;
            // 3d: one point per thread at 128 registers (16 warps) beats two points at 255 (8 warps) -- measured
            if (!force_ilp && h->dim == 3 && ilp == 2) continue;
            const unsigned wmax = max_threads_for(h->dim, ilp, h->xpp) / 32 - (staged ? 1 : 0);
            const unsigned long long upt = ((P.Nvel_loc + G - 1) / G + ilp - 1) / ilp;
            for (unsigned W = 1; W <= wmax; ++W) {
                if (force_w && static_cast<int>(W) != force_w) continue;
                const unsigned long long rpt = (upt + W - 1) / W;
                const unsigned long long R = rpt * P.n_tiles;
                const unsigned long long rpc = (R + grid - 1) / grid;
                const unsigned long long ctas = (R + rpc - 1) / rpc;
                const double chains = static_cast<double>(W) * ilp;
                // time of a CTA-round ~ max(latency floor, throughput term); ILP 2 shares the per-level bookkeeping
                double cost = static_cast<double>(rpc) * (chains > csat ? chains : csat) * (ilp == 2 ? 0.92 : 1.0);
                cost *= 1.0 + 1e-3 * (static_cast<double>(grid) - static_cast<double>(ctas)) / grid; // prefer more busy SMs
                cost *= 1.0 + 1e-4 * chains;                                                          // then narrower CTAs
                if (cost < best_cost) { best_cost = cost; best_ilp = ilp; best_W = W; }
            }
        }
        if (best_cost >= 1e300) return fail(h, NUFI_B200_ERR_ARG, "NUFI_B200_ILP / NUFI_B200_W override out of range");
    }
    const int ilp = best_ilp;
    P.W = best_W;
    P.upt = static_cast<unsigned>(((P.Nvel_loc + G - 1) / G + ilp - 1) / ilp);
    P.rpt = (P.upt + P.W - 1) / P.W;
    const unsigned long long R64 = static_cast<unsigned long long>(P.rpt) * P.n_tiles;
    if (R64 >= (1ull << 31)) return fail(h, NUFI_B200_ERR_RANGE, "quadrature range too large for one launch");
    P.R = static_cast<unsigned>(R64);
    P.rpc = (P.R + grid - 1) / grid;
    P.Tmax = (P.rpc - 1) / P.rpt + 2;
    P.interleave = env_int("NUFI_B200_INTERLEAVE", 1) ? 1 : 0;
    const unsigned threads = (P.W + (staged ? 1 : 0)) * 32;
    const size_t smem_bytes = kSmemFixed + (staged ? static_cast<size_t>(P.stages) * P.stage_bytes : 0);

    // ---- slots: one per (CTA, tile it touches); every slot the finish kernel reads is written by its CTA
    if (!metrics) {
        const size_t need = static_cast<size_t>(grid) * P.Tmax * 32;
        if (need > h->partials_cap) {
            if (h->d_partials) cudaFree(h->d_partials);
            h->d_partials = nullptr;
            h->partials_cap = 0;
            if (cudaMalloc(&h->d_partials, need * sizeof(double)) != cudaSuccess)
                return fail(h, NUFI_B200_ERR_ALLOC, "cudaMalloc of the rho partial slots failed");
            h->partials_cap = need;
        }
        P.slots = h->d_partials;
    } else {
        P.mpartials = h->d_mpartials;
        NUFI_CUDA_CHECK(h, cudaMemsetAsync(h->d_mpartials, 0, sizeof(double) * 4 * grid, h->stream));
    }

    // ---- slot reduction: by the fused tail (defer), else by the kernel's own last-CTA epilogue (optionally pushing to the peers)
    EpilogueParams E{};
    bool tail_reduces = false;
    if (!metrics) {
        FinishParams F{};
        F.slots = h->d_partials;
        F.rho_partial = h->d_rho_partial;
        const bool all_nodes = q_begin == 0 && q_end == h->n_nodes * h->n_vel; // every node's value gets written
        const bool whole = all_nodes && P.vstride == 1;                         // ... and it is the complete sum
        F.rho_full = whole ? h->d_rho_full : nullptr;
        F.dV = h->dim == 1 ? P.du : (h->dim == 2 ? P.du * P.dv : P.du * P.dv * P.dw); // rho.hpp:145, 307, 459
        F.l_first = P.l_first; F.l_last = P.l_last;
        F.rpt = P.rpt; F.rpc = P.rpc; F.Tmax = P.Tmax;
        F.n_tiles = P.n_tiles;
        F.TN = P.TN;
        if (!all_nodes) NUFI_CUDA_CHECK(h, cudaMemsetAsync(h->d_rho_partial, 0, sizeof(double) * h->n_nodes, h->stream));
        h->fin = F;
        tail_reduces = defer_finish && whole && !h->fin_push;
        h->fin_pending = tail_reduces;
        // many tiles (TN < 32 on a large grid): the one-CTA epilogue would serialise them; use the multi-block finish kernels
        const bool epilogue = P.n_tiles <= 256;
        if (!tail_reduces && !epilogue) {
            h->fin_pending = true; // launch_finish() below runs finish_rho_kernel / finish_push_kernel (h->fin_push kept)
        } else if (!tail_reduces) {
            E.mode = h->fin_push ? 2 : 1;
            E.n_active = (P.R + P.rpc - 1) / P.rpc;
            E.done = h->d_done;
            E.F = F;
            if (h->fin_push) E.X = h->px.push;
        }
        if (tail_reduces || epilogue) h->fin_push = false;
    }

    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    if (h->kernel_timing) { // off by default: an event between two kernels keeps the second from launching programmatically
        int rc = ev_acquire(h, &ev_start, &ev_stop);
        if (rc) return rc;
        NUFI_CUDA_CHECK(h, cudaEventRecord(ev_start, h->stream));
    }
    cudaError_t e;
    if (h->dim == 1) e = launch_dim<1>(P, E, ilp, false, staged, pow2, grid, threads, smem_bytes, h->stream);
    else if (h->dim == 2) e = launch_dim<2>(P, E, ilp, h->xpp, staged, pow2, grid, threads, smem_bytes, h->stream);
    else e = launch_dim<3>(P, E, ilp, h->xpp, staged, pow2, grid, threads, smem_bytes, h->stream);
    NUFI_CUDA_CHECK(h, e);
    if (h->kernel_timing) {
        NUFI_CUDA_CHECK(h, cudaEventRecord(ev_stop, h->stream));
        h->ev_pending += 1;
    }
    h->launches += 1;
    const char *fmt = h->xpp ? "/xpp" : "";
    char tn[16] = "";
    if (P.TN != 32) std::snprintf(tn, sizeof(tn), "/tn%u", P.TN);
    if (staged) std::snprintf(h->variant_buf, sizeof(h->variant_buf), "smem-tma%s/ilp%d/W%u/Lc%dx%d%s", fmt, ilp, P.W, P.Lc, P.stages, tn);
    else std::snprintf(h->variant_buf, sizeof(h->variant_buf), "global%s/ilp%d/W%u%s", fmt, ilp, P.W, tn);
    h->last_variant = h->variant_buf;

    if (!metrics) {
        if (h->fin_pending && !tail_reduces) return launch_finish(h); // many tiles: multi-block slot reduction (+ push)
        return NUFI_B200_OK; // rho_partial (and rho_full) are complete when the kernel ends, or the tail reduces the slots
    } else {
        finish_metrics_kernel<<<1, 32, 0, h->stream>>>(h->d_mpartials, grid, h->d_metrics);
        NUFI_CUDA_CHECK(h, cudaGetLastError());
    }
    h->launches += 1;
    return NUFI_B200_OK;
}

// fills the geometry / f0 part of BtParams shared by every launch
static void fill_common(const Handle *h, BtParams &P)
{
    const nufi_b200_config3d &c = h->c;
    P.dim = h->dim;
    P.Nx = static_cast<int>(c.Nx); P.Ny = static_cast<int>(c.Ny); P.Nz = static_cast<int>(c.Nz);
    P.Nu = static_cast<int>(c.Nu); P.Nv = static_cast<int>(c.Nv); P.Nw = static_cast<int>(c.Nw);
    P.sx = h->sx; P.sxy = h->sxy;
    P.hist = h->d_hist;
    P.level_bytes = static_cast<unsigned>(h->level_stride * 8);
    P.ncx = -(c.dt * c.dx_inv); P.ncy = -(c.dt * c.dy_inv); P.ncz = -(c.dt * c.dz_inv);
    const double scale = h->dim == 2 ? 12.0 : 72.0;
    P.gx = -c.dt * c.dx_inv / scale / (h->xpp ? 3.0 : 1.0); P.gy = -c.dt * c.dy_inv / scale; P.gz = -c.dt * c.dz_inv / scale;
    P.x_min = c.x_min; P.y_min = c.y_min; P.z_min = c.z_min;
    P.dx = c.dx; P.dy = c.dy; P.dz = c.dz;
    P.f0_kind = h->f0.kind;
    for (int i = 0; i < 4; ++i) P.f0p[i] = h->f0.p[i];
}

int launch_sample_f(Handle *h, size_t n, size_t npts, const double *d_pts, double *d_out, bool full, bool feet)
{
    SampleParams S{};
    S.feet = feet ? 1 : 0;
    fill_common(h, S.P);
    const nufi_b200_config3d &c = h->c;
    S.P.metrics = full ? 1 : 0;
    S.P.first_level = full ? (n == 0 ? -1 : static_cast<int>(n)) : static_cast<int>(n) - 1;
    S.Lx = c.Lx; S.Ly = c.Ly; S.Lz = c.Lz; S.Lx_inv = c.Lx_inv; S.Ly_inv = c.Ly_inv; S.Lz_inv = c.Lz_inv;
    S.dx_inv = c.dx_inv; S.dy_inv = c.dy_inv; S.dz_inv = c.dz_inv;
    S.pts = d_pts; S.out = d_out; S.npts = npts; S.with_first_half_kick = full ? 1 : 0;
    const unsigned blocks = static_cast<unsigned>(std::min<size_t>((npts + 127) / 128, 148 * 8));
    if (h->dim == 1) sample_f_kernel<1, false><<<blocks, 128, 0, h->stream>>>(S);
    else if (h->dim == 2) { if (h->xpp) sample_f_kernel<2, true><<<blocks, 128, 0, h->stream>>>(S); else sample_f_kernel<2, false><<<blocks, 128, 0, h->stream>>>(S); }
    else { if (h->xpp) sample_f_kernel<3, true><<<blocks, 128, 0, h->stream>>>(S); else sample_f_kernel<3, false><<<blocks, 128, 0, h->stream>>>(S); }
    NUFI_CUDA_CHECK(h, cudaGetLastError());
    h->launches += 1;
    return NUFI_B200_OK;
}

int launch_sample_field(Handle *h, const double *d_ref_level, int der, size_t npts, const double *d_pts, double *d_out)
{
    const nufi_b200_config3d &c = h->c;
    FieldSampleParams S{};
    S.dim = h->dim; S.Nx = static_cast<int>(c.Nx); S.Ny = static_cast<int>(c.Ny); S.Nz = static_cast<int>(c.Nz); S.der = der;
    S.x_min = c.x_min; S.y_min = c.y_min; S.z_min = c.z_min;
    S.Lx = c.Lx; S.Ly = c.Ly; S.Lz = c.Lz; S.Lx_inv = c.Lx_inv; S.Ly_inv = c.Ly_inv; S.Lz_inv = c.Lz_inv;
    S.dx_inv = c.dx_inv; S.dy_inv = c.dy_inv; S.dz_inv = c.dz_inv;
    S.level = d_ref_level; S.pts = d_pts; S.out = d_out; S.npts = npts;
    const unsigned blocks = static_cast<unsigned>(std::min<size_t>((npts + 127) / 128, 148 * 8));
    sample_field_kernel<<<blocks, 128, 0, h->stream>>>(S);
    NUFI_CUDA_CHECK(h, cudaGetLastError());
    h->launches += 1;
    return NUFI_B200_OK;
}

int launch_flag_only_push(Handle *h)
{
    FinishParams F{};
    F.n_tiles = 0;
    h->fin = F;
    h->fin_pending = true;
    h->fin_push = true;
    return launch_finish(h);
}

int launch_finish(Handle *h)
{
    if (!h->fin_pending) return NUFI_B200_OK;
    if (h->fin_push) NUFI_CUDA_CHECK(h, launch_chained(h, finish_push_kernel, dim3(h->fin.n_tiles ? h->fin.n_tiles : 1), dim3(256), 0, h->fin, h->px.push));
    else NUFI_CUDA_CHECK(h, launch_chained(h, finish_rho_kernel, dim3(h->fin.n_tiles), dim3(256), 0, h->fin));
    h->fin_push = false;
    NUFI_CUDA_CHECK(h, cudaGetLastError());
    h->fin_pending = false;
    h->launches += 1;
    return NUFI_B200_OK;
}

} // namespace nufi_b200
