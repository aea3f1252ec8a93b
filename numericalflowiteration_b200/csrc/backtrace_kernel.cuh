// backtrace_kernel.cuh -- the backtrace kernel template and its device helpers, shared by the translation units that
// instantiate it: backtrace.cu (order 4: per-cell polynomial level formats, shared-memory staging) and
// backtrace_generic.cu (orders 3, 5..8: B-spline window of order^dim coefficients straight from the reference layout).
// See backtrace.cu for the design notes.
#pragma once
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

template <bool STAGED> __device__ __forceinline__ double2 ld2(const double2 *p)
{
    if constexpr (STAGED) return *p;
    else return __ldg(p);
}

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
    // level = [Nx x (p1, p2)] [Nx x p0]: one 128-bit and one 64-bit load per point-step (a 128-bit load is served per quarter-warp:
    // eight lanes instead of sixteen have to fall on distinct banks, profiles/r02_c2_bank_conflicts.txt).  (p1, p2) feed the first
    // FMA of the kick; p0, whose address costs one more add, is only needed by the second.
    const double2 p12 = ld2<STAGED>(reinterpret_cast<const double2 *>(lev) + p.cell[0]);
    const double p1 = p12.x, p2 = p12.y, p0 = ld<STAGED>(lev + 2 * P.Nx + p.cell[0]);
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

// ---- generic spline order K (3, 5..8; nufi/splines.hpp:39-110 is generic in `order`): basis values N_i(t) and first
// derivatives D_i(t), i < K, of the K B-splines that overlap a cell, t in [0,1] the reference coordinate on the cell.  Uniform
// knots: with B_p[i] = M_p(t + p - i) (M_p the cardinal B-spline of degree p) the Cox-de Boor recurrence reads
//   B_p[i] = ((t + p - i) B_{p-1}[i-1] + (1 + i - t) B_{p-1}[i]) / p,      B_p'[i] = B_{p-1}[i-1] - B_{p-1}[i],
// evaluated in place in registers (all loops unrolled).
template <int K> __device__ __forceinline__ void basis_generic(double t, double (&N)[K], double (&D)[K])
{
    double b[K];
    b[0] = 1.0;
#pragma unroll
    for (int i = 1; i < K; ++i) b[i] = 0.0;
#pragma unroll
    for (int p = 1; p < K; ++p) {
        if (p == K - 1) { // b holds degree K-2: its differences are the derivatives of degree K-1
            D[0] = -b[0];
#pragma unroll
            for (int i = 1; i < K - 1; ++i) D[i] = b[i - 1] - b[i];
            D[K - 1] = b[K - 2];
        }
        const double inv = 1.0 / p;
        b[p] = (t * b[p - 1]) * inv;
#pragma unroll
        for (int i = p - 1; i >= 1; --i) b[i] = fma(t + (p - i), b[i - 1], ((1 + i) - t) * b[i]) * inv;
        b[0] = ((1.0 - t) * b[0]) * inv;
    }
#pragma unroll
    for (int i = 0; i < K; ++i) N[i] = b[i];
}

// One point, one level, spline order K: window of K^DIM coefficients from the reference layout (row stride P.sx = Nx+K-1,
// plane stride P.sxy), x-contraction with values and derivatives, then y, then z -- the same shared-subexpression scheme as
// the cubic steps.  P.gx/gy/gz = -dt*dx_inv (no basis scaling folded in).
template <int DIM, int K, int KIND, bool STAGED, bool POW2>
__device__ __forceinline__ void step_generic(Point<DIM> &p, const double *lev, const BtParams &P, unsigned &bad)
{
    if constexpr (KIND != FIRST) {
        relocate<POW2>(p.tau[0], p.cell[0], fma(P.ncx, p.vel[0], p.tau[0]), P.Nx, bad);
        if constexpr (DIM >= 2) relocate<POW2>(p.tau[1], p.cell[1], fma(P.ncy, p.vel[1], p.tau[1]), P.Ny, bad);
        if constexpr (DIM >= 3) relocate<POW2>(p.tau[2], p.cell[2], fma(P.ncz, p.vel[2], p.tau[2]), P.Nz, bad);
    }
    const double h = KIND == FULL ? 1.0 : 0.5;
    double Nx[K], Dx[K];
    basis_generic<K>(0.5 + p.tau[0], Nx, Dx);
    if constexpr (DIM == 1) {
        const double *c = lev + p.cell[0];
        double s = ld<STAGED>(c) * Dx[0];
#pragma unroll
        for (int a = 1; a < K; ++a) s = fma(ld<STAGED>(c + a), Dx[a], s);
        p.vel[0] = fma(h * P.gx, s, p.vel[0]);
    } else if constexpr (DIM == 2) {
        double Ny[K], Dy[K];
        basis_generic<K>(0.5 + p.tau[1], Ny, Dy);
        const double *row = lev + (p.cell[1] * P.sx + p.cell[0]);
        double Sx = 0, Sy = 0;
#pragma unroll
        for (int b = 0; b < K; ++b) {
            const double c0 = ld<STAGED>(row);
            double pv = c0 * Nx[0], qv = c0 * Dx[0];
#pragma unroll
            for (int a = 1; a < K; ++a) {
                const double ca = ld<STAGED>(row + a);
                pv = fma(ca, Nx[a], pv);
                qv = fma(ca, Dx[a], qv);
            }
            Sx = fma(Ny[b], qv, Sx);
            Sy = fma(Dy[b], pv, Sy);
            row += P.sx;
        }
        p.vel[0] = fma(h * P.gx, Sx, p.vel[0]);
        p.vel[1] = fma(h * P.gy, Sy, p.vel[1]);
    } else {
        double Ny[K], Dy[K], Nz[K], Dz[K];
        basis_generic<K>(0.5 + p.tau[1], Ny, Dy);
        basis_generic<K>(0.5 + p.tau[2], Nz, Dz);
        const double *plane = lev + (p.cell[2] * P.sxy + p.cell[1] * P.sx + p.cell[0]);
        double Sx = 0, Sy = 0, Sz = 0;
#pragma unroll
        for (int c = 0; c < K; ++c) {
            const double *row = plane;
            double r = 0, s = 0, w = 0;
#pragma unroll
            for (int b = 0; b < K; ++b) {
                const double c0 = ld<STAGED>(row);
                double pv = c0 * Nx[0], qv = c0 * Dx[0];
#pragma unroll
                for (int a = 1; a < K; ++a) {
                    const double ca = ld<STAGED>(row + a);
                    pv = fma(ca, Nx[a], pv);
                    qv = fma(ca, Dx[a], qv);
                }
                r = fma(Ny[b], qv, r);
                s = fma(Dy[b], pv, s);
                w = fma(Ny[b], pv, w);
                row += P.sx;
            }
            Sx = fma(Nz[c], r, Sx);
            Sy = fma(Nz[c], s, Sy);
            Sz = fma(Dz[c], w, Sz);
            plane += P.sxy;
        }
        p.vel[0] = fma(h * P.gx, Sx, p.vel[0]);
        p.vel[1] = fma(h * P.gy, Sy, p.vel[1]);
        p.vel[2] = fma(h * P.gz, Sz, p.vel[2]);
    }
}

// ---- "xpp" level format (tail.cu: row_poly): per (row, cell) the cubic A(tau_x) = sum_a c_a 6 N_a, stored as two
// double2 halves [a0 a1] and [a2 a3] (row stride P.sx double2, the second half Nx further).  The x-contraction of a
// window row becomes two Horner evaluations (value: 3 FMA, derivative A' = a1 + 2 tau (a2 + 1.5 tau a3): 2 FMA) instead of
// eight FMAs plus the x basis; loads are 2 x 128-bit per row, conflict-free for consecutive cells.  P.gx carries the 1/3
// of A' = 3 sum_a c_a 2 N'_a.
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

template <int DIM, int KIND, bool STAGED, bool POW2, bool XPP, int ORDER = 4>
__device__ __forceinline__ void step(Point<DIM> &p, const double *lev, const BtParams &P, unsigned &bad)
{
    if constexpr (ORDER != 4) step_generic<DIM, ORDER, KIND, STAGED, POW2>(p, lev, P, bad);
    else if constexpr (DIM == 1) step1d<KIND, STAGED, POW2>(p, lev, P, bad);
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
template <int DIM, bool XPP, int ORDER = 4> __device__ __noinline__ void slow_trace(Point<DIM> &p, const BtParams &P)
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
        step<DIM, FIRST, false, false, XPP, ORDER>(p, lev, P, bad);
        if (!(first || m == 0)) step<DIM, FIRST, false, false, XPP, ORDER>(p, lev, P, bad);
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
// ---- thread-block cluster of two CTAs sharing one stream of the history (BtParams::cluster == 2): every chunk is fetched from L2
// ONCE per pair -- each CTA issues half of it as a multicast bulk copy that lands at the same shared-memory offset in both CTAs
// and counts on both CTAs' `full` barriers -- half the L2 reads per SM.  Opt-in (NUFI_B200_CLUSTER=2): measured, it buys
// nothing on B200 (profiles/r02_fill_path.md).
__device__ __forceinline__ unsigned cluster_ctarank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() // every thread of every CTA of the cluster
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(unsigned long long *bar, unsigned rank)
{
    asm volatile("{ .reg .b32 ra; mapa.shared::cluster.u32 ra, %0, %1; mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra]; }" ::"r"(smem_u32(bar)), "r"(rank)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(unsigned long long *bar, unsigned parity) // acquires what a peer CTA released
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> the same shared-memory offset of every CTA in `mask`, completion counted on each one's mbarrier
__device__ __forceinline__ void bulk_g2s_multicast(void *dst, const void *src, unsigned bytes, unsigned long long *bar, unsigned short mask)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
// named barrier over the consumer warps only (the producer warp never joins)
__device__ __forceinline__ void consumer_sync(unsigned threads) { asm volatile("bar.sync 1, %0;" ::"r"(threads) : "memory"); }

constexpr int kMaxStages = 8;
constexpr unsigned kBarBytes = 3 * kMaxStages * 8; // full[8], empty[8], peer_free[8] (cluster pairs)
constexpr unsigned kRedBytes = 32 * 32 * 8;        // consumer-warp reduction scratch [32 warps][32 lanes]
constexpr unsigned kSmemFixed = kBarBytes + kRedBytes;

#ifndef NUFI_3D_MT
#define NUFI_3D_MT 512 // thread bound of the 3d B-spline kernel, one point per thread (128 registers)
#endif
#ifndef NUFI_3D_XPP_MT
#define NUFI_3D_XPP_MT 640
#endif
template <int DIM, int ILP, bool XPP, int ORDER = 4> struct Tune
{
    // thread-count upper bound handed to __launch_bounds__ (sets the register budget); the xpp steps need fewer registers;
    // the generic-order steps keep up to 6 basis vectors of ORDER doubles in registers
    static constexpr int max_threads =
        ORDER != 4 ? (DIM == 1 ? 512 : 256)
                   : (DIM == 1 ? 1024 : (DIM == 2 ? (ILP == 1 ? 768 : 512) : (ILP == 1 ? (XPP ? NUFI_3D_XPP_MT : NUFI_3D_MT) : 256)));
};

// ---------------------------------------------------------------- the kernel
template <int DIM, int ILP, bool STAGED, bool POW2, bool XPP, int ORDER = 4>
__global__ void __launch_bounds__(Tune<DIM, ILP, XPP, ORDER>::max_threads, 1)
    backtrace_kernel(const __grid_constant__ BtParams P, const __grid_constant__ EpilogueParams E)
{
    __shared__ unsigned int s_ticket;
    __shared__ unsigned short s_tfirst[256]; // epilogue: first tile of every CTA (filled below, read by the last CTA only)
    pdl_trigger(); // the slot reduction / field tail behind this launch may be scheduled as soon as an SM has room
    if (E.mode == 1) // visible to the epilogue through the CTA's barriers (the producer warp's share through the __syncthreads below)
        for (unsigned b = threadIdx.x; b < E.n_active && b < 256; b += blockDim.x)
            s_tfirst[b] = static_cast<unsigned short>((b * E.F.rpc) / E.F.rpt);
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long *full = reinterpret_cast<unsigned long long *>(smem);
    unsigned long long *empty = full + kMaxStages;
    unsigned long long *peer_free = empty + kMaxStages; // cluster pairs: "the other CTA's stage s may be overwritten"
    double(*sred)[32] = reinterpret_cast<double(*)[32]>(smem + kBarBytes);
    const unsigned char *ring = smem + kSmemFixed;

    const int lane = threadIdx.x & 31;
    const unsigned warp = threadIdx.x >> 5;
    const unsigned W = P.W;
    const bool paired = STAGED && P.cluster == 2; // CTAs 2k, 2k+1 share the history stream (multicast)
    // this CTA's run of CTA-rounds
    const unsigned g0 = blockIdx.x * P.rpc;
    const unsigned my_rounds = g0 < P.R ? min(P.rpc, P.R - g0) : 0;
    unsigned pair_rounds = my_rounds; // a pair walks the ring in lockstep: the CTA with fewer rounds idles through the rest
    if (paired) {
        const unsigned g0p = (blockIdx.x ^ 1u) * P.rpc;
        pair_rounds = max(my_rounds, g0p < P.R ? min(P.rpc, P.R - g0p) : 0u);
    }
    if (!paired && my_rounds == 0) return; // whole CTA idle (uniform)
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
                mbar_init(&peer_free[s], 1);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (paired) {
            cluster_sync_all(); // the partner's barriers exist before anything is sent to them
            if (pair_rounds == 0) return;
        }
    }

    // Launched programmatically behind the kernel ahead on the stream (in a run of fused steps: the previous step's field tail,
    // which is still writing the newest level): everything above ran beside it; nothing below may start before it has finished.
    pdl_wait();

    if (STAGED && warp == W) {
        // ------------------------------------------------ producer warp: stream chunks newest -> oldest, once per round
        if (lane == 0) {
            int s = 0;
            unsigned ph = 0;
            bool primed = false;
            const unsigned crank = paired ? cluster_ctarank() : 0;
            for (unsigned r = 0; r < pair_rounds; ++r)
                for (int ci = 0; ci < n_chunks; ++ci) {
                    if (primed) {
                        mbar_wait(&empty[s], ph ^ 1u);
                        if (paired) { // both CTAs of the pair must be done with stage s before either half lands in both
                            mbar_arrive_remote(&peer_free[s], crank ^ 1u);
                            mbar_wait_cluster(&peer_free[s], ph ^ 1u);
                        }
                    }
                    const bool bottom = ci == n_chunks - 1;
                    const int cnt = bottom ? rem_levels : P.Lc;
                    const int lv_lo = bottom ? 0 : P.first_level - (ci + 1) * P.Lc + 1;
                    const unsigned bytes = static_cast<unsigned>(cnt) * P.level_bytes;
                    mbar_expect_tx(&full[s], bytes);
                    const unsigned char *src = reinterpret_cast<const unsigned char *>(P.hist) + static_cast<size_t>(lv_lo) * P.level_bytes;
                    unsigned char *dst = const_cast<unsigned char *>(ring) + static_cast<size_t>(s) * P.stage_bytes;
                    if (paired) { // this CTA's half of the chunk, to both CTAs
                        const unsigned half = (bytes / 2u) & ~15u;
                        const unsigned lo = crank == 0 ? 0u : half, hi = crank == 0 ? half : bytes;
                        for (unsigned off = lo; off < hi; off += 32768u)
                            bulk_g2s_multicast(dst + off, src + off, min(32768u, hi - off), &full[s], static_cast<unsigned short>(3));
                    } else {
                        for (unsigned off = 0; off < bytes; off += 32768u)
                            bulk_g2s(dst + off, src + off, min(32768u, bytes - off), &full[s]);
                    }
                    if (++s == P.stages) { s = 0; ph ^= 1u; primed = true; }
                }
        }
        return;
    }

    // ---------------------------------------------------- consumer warps
    // sum of f over this thread's velocities, compensated (Knuth two-sum: acc carries the rounded sum, acc_lo the rounding
    // errors it dropped): the reference adds up to Nu*Nv*Nw terms sequentially (rho.hpp:299-306) and its result carries that
    // sum's rounding error; here the sum is exact to an ulp whatever its length, at 6 FP64 adds per point (not per point-step)
    double acc = 0, acc_lo = 0;
    double m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    unsigned cur_tile = t_first;
    int s = 0;
    unsigned ph = 0;

    auto flush_tile = [&](unsigned tile) { // CTA-uniform: every consumer warp calls it
        sred[warp][lane] = acc + acc_lo;
        consumer_sync(W * 32);
        if (warp == 0) {
            double sum = 0, lo = 0; // the consumer warps' sums in warp order, compensated as well
            for (unsigned w = 0; w < W; ++w) two_sum(sum, lo, sred[w][lane]);
            sum += lo;
            // TN < 32: the 32/TN lanes that traced the same node (lane % TN) are combined by a fixed shuffle tree
            for (unsigned off = 16; off >= P.TN; off >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, off);
            const size_t slot = (static_cast<size_t>(blockIdx.x) * P.Tmax + (tile - t_first)) * 32 + lane;
            const double val = static_cast<unsigned>(lane) < P.TN ? sum : 0.0;
            if (P.slots_ll) peer_store_double(P.slots_ll + slot, val, P.slot_flag); // fused step: the resident tail polls for it
            else P.slots[slot] = val;
        }
        consumer_sync(W * 32);
        acc = 0;
        acc_lo = 0;
    };

    auto idle_round = [&]() { // no work for this warp in this round: keep the stage protocol going
        if constexpr (STAGED) {
            for (int ci = 0; ci < n_chunks; ++ci) {
                mbar_wait(&full[s], ph);
                if (lane == 0) mbar_arrive(&empty[s]);
                if (++s == P.stages) { s = 0; ph ^= 1u; }
            }
        }
    };

    for (unsigned r = 0; r < pair_rounds; ++r) {
        if (r >= my_rounds) { // the partner CTA of a pair still has rounds to go
            idle_round();
            continue;
        }
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
        if (jc >= P.upt) { // no unit for this warp in this round
            idle_round();
            continue;
        }

        unsigned long long l = P.l_first + static_cast<unsigned long long>(tile) * P.TN + (lane & (P.TN - 1));
        const unsigned vsub = static_cast<unsigned>(lane) >> P.TNlog2, G = 32u >> P.TNlog2; // velocity sub-index within the warp-unit
        const bool node_ok = l <= P.l_last;
        if (!node_ok) l = P.l_first;
        int ix, iy = 0, iz = 0;
        double tx0 = -0.5; // node ix sits on the left edge of cell ix: xi = ix  ->  tau = -1/2
        if constexpr (DIM == 1) {
            ix = static_cast<int>(l);
            if (P.mgrid) // 1d metrics on their own grid (nufi/cuda_kernel.cu:55-70): node l of THAT grid, located in the field grid
                locate(P.mx_min + ix * P.mdx, P.x_min, P.Lx, P.Lx_inv, P.dx_inv, P.Nx, ix, tx0);
        } else {
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
            pt[i].cell[0] = ix;
            pt[i].tau[0] = tx0;
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
                    for (int i = 0; i < ILP; ++i) step<DIM, FIRST, STAGED, POW2, XPP, ORDER>(pt[i], lev, P, bad[i]);
                    --cnt;
                    lev -= level_doubles;
                }
                if (bottom) --cnt; // level 0 takes the half kick below
                for (int k = cnt >> 1; k > 0; --k) { // full-kick steps, two levels per trip
#pragma unroll
                    for (int i = 0; i < ILP; ++i) step<DIM, FULL, STAGED, POW2, XPP, ORDER>(pt[i], lev, P, bad[i]);
#pragma unroll
                    for (int i = 0; i < ILP; ++i) step<DIM, FULL, STAGED, POW2, XPP, ORDER>(pt[i], lev - level_doubles, P, bad[i]);
                    lev -= 2 * level_doubles;
                }
                if (cnt & 1) {
#pragma unroll
                    for (int i = 0; i < ILP; ++i) step<DIM, FULL, STAGED, POW2, XPP, ORDER>(pt[i], lev, P, bad[i]);
                }
                if (bottom) { // level 0: drift + half kick
#pragma unroll
                    for (int i = 0; i < ILP; ++i) step<DIM, LAST, STAGED, POW2, XPP, ORDER>(pt[i], base, P, bad[i]);
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
                pt[i].tau[0] = tx0;
                if constexpr (DIM >= 2) pt[i].cell[1] = iy;
                if constexpr (DIM >= 3) pt[i].cell[2] = iz;
                slow_trace<DIM, XPP, ORDER>(pt[i], P);
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
                two_sum(acc, acc_lo, f);
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

    if (my_rounds == 0) return; // idle half of a pair: no slots, no part in the epilogue
    if (!P.metrics) {
        flush_tile(cur_tile);
        if (E.mode) { // ---- epilogue: the last CTA to arrive reduces the slots of all tiles
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
                            F.rho_partial[l] = -F.dV * tot;
                            if (F.rho_full) F.rho_full[l] = fma(-F.dV, tot, 1.0);
                        }
                    }
                    consumer_sync(W * 32);
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

// ---------------------------------------------------------------- sampling at arbitrary points (plots, diagnostics)
// f(t_n, x, v) at arbitrary phase-space points: one thread per point, history read from global memory.
template <int DIM, bool XPP, int ORDER = 4> __global__ void sample_f_kernel(const __grid_constant__ SampleParams S)
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
        if (P.first_level >= 0) slow_trace<DIM, XPP, ORDER>(p, P); // robust path: true modulo wrap, any jump length
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

} // namespace

} // namespace nufi_b200
