// backtrace.cu -- the hot path: backward Stoermer-Verlet through the coefficient history + f0 + rho reduction.
//
// Replaces nufi/cuda_kernel.cu:31-51, 210-237, 393-426 (cuda_eval_rho) and :53-79, 239-271, 428-466
// (cuda_eval_metrics), i.e. the device flavour of nufi/rho.hpp eval_ftilda / eval_f / eval_rho with
// nufi/fields.hpp eval and nufi/splines.hpp inlined.  Written from scratch for sm_100a:
//
//  * Work layout.  A warp-task is 32 consecutive spatial nodes (x fastest) sharing ONE velocity node, so the
//    lanes of a warp drift rigidly and read neighbouring coefficients (conflict-free, coalesced) for the
//    whole history, instead of the reference's "velocity fastest" layout whose lanes fan out and all hit one
//    rho address with atomics.  Each thread owns a fixed node and sums f over its velocities in registers;
//    per-(CTA,warp,tile) partial sums go to a slot array that a second tiny kernel adds in a fixed order:
//    no atomics, run-to-run deterministic.
//  * History access.  Persistent CTAs (one per SM).  Staged variant: a producer warp streams level after
//    level from the HBM/L2-resident history into a shared-memory ring with cp.async.bulk (TMA bulk copy,
//    SASS UBLKCP) + mbarrier full/empty pairs; consumer warps never wait on global memory.  Global variant
//    (levels too big for shared memory, large 3d): read-only loads served by L1/L2.
//  * Arithmetic.  Position per dimension = (cell k, centred offset tau); floor() and the float->int
//    conversion (quarter-rate pipes) are replaced by the 1.5*2^52 rounding trick on the FP64 pipe.  1d levels
//    are stored as per-cell quadratics of dt*E (3 doubles/cell, see tail.cu) so a 1d point-step is 7 FP64
//    instructions; 2d/3d use the cubic B-spline window (16/64 doubles) with value and derivative bases
//    computed once per dimension and shared between the field components.
#include "internal.cuh"

namespace nufi_b200
{

namespace
{

constexpr double kMagic = 6755399441055744.0; // 1.5 * 2^52: adding it rounds to the nearest integer

__device__ __forceinline__ int wrap_cell(int k, int N)
{
    if (k < 0) k += N;
    if (k >= N) k -= N;
    if (static_cast<unsigned>(k) >= static_cast<unsigned>(N)) { // more than one period in a single step: rare
        k %= N;
        if (k < 0) k += N;
    }
    return k;
}

// t2 = tau - drift.  New cell/offset such that k + 1/2 + tau is preserved and tau in [-1/2, 1/2].
__device__ __forceinline__ void relocate(double &tau, int &k, double t2, int N)
{
    const double y = t2 + kMagic;
    const int dk = __double2loint(y);
    const double r = y - kMagic;
    tau = t2 - r;
    k = wrap_cell(k + dk, N);
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

// cd = cx*d (d = 0 for eval_f's initial half kick, which does not drift), hg = h*g (h = 1/2 for half kicks)
template <bool STAGED>
__device__ __forceinline__ void step1d(Point<1> &p, const double *lev, const BtParams &P, double d, double h)
{
    relocate(p.tau[0], p.cell[0], fma(-P.cx * d, p.vel[0], p.tau[0]), P.Nx);
    const double *c = lev + p.cell[0];
    const double p0 = ld<STAGED>(c), p1 = ld<STAGED>(c + P.sx), p2 = ld<STAGED>(c + 2 * P.sx);
    const double t = p.tau[0];
    p.vel[0] = fma(h, fma(t, fma(t, p2, p1), p0), p.vel[0]);
}

template <bool STAGED>
__device__ __forceinline__ void step2d(Point<2> &p, const double *lev, const BtParams &P, double d, double h)
{
    relocate(p.tau[0], p.cell[0], fma(-P.cx * d, p.vel[0], p.tau[0]), P.Nx);
    relocate(p.tau[1], p.cell[1], fma(-P.cy * d, p.vel[1], p.tau[1]), P.Ny);
    double Nx[4], Dx[4], Ny[4], Dy[4];
    basis4(p.tau[0], Nx, Dx);
    basis4(p.tau[1], Ny, Dy);
    const double *row = lev + p.cell[1] * P.sx + p.cell[0];
    double Sx = 0, Sy = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const double c0 = ld<STAGED>(row), c1 = ld<STAGED>(row + 1), c2 = ld<STAGED>(row + 2), c3 = ld<STAGED>(row + 3);
        const double pv = fma(c3, Nx[3], fma(c2, Nx[2], fma(c1, Nx[1], c0 * Nx[0])));
        const double qv = fma(c3, Dx[3], fma(c2, Dx[2], fma(c1, Dx[1], c0 * Dx[0])));
        Sx = fma(Ny[b], qv, Sx);
        Sy = fma(Dy[b], pv, Sy);
        row += P.sx;
    }
    p.vel[0] = fma(h * P.gx, Sx, p.vel[0]);
    p.vel[1] = fma(h * P.gy, Sy, p.vel[1]);
}

template <bool STAGED>
__device__ __forceinline__ void step3d(Point<3> &p, const double *lev, const BtParams &P, double d, double h)
{
    relocate(p.tau[0], p.cell[0], fma(-P.cx * d, p.vel[0], p.tau[0]), P.Nx);
    relocate(p.tau[1], p.cell[1], fma(-P.cy * d, p.vel[1], p.tau[1]), P.Ny);
    relocate(p.tau[2], p.cell[2], fma(-P.cz * d, p.vel[2], p.tau[2]), P.Nz);
    double Nx[4], Dx[4], Ny[4], Dy[4], Nz[4], Dz[4];
    basis4(p.tau[0], Nx, Dx);
    basis4(p.tau[1], Ny, Dy);
    basis4(p.tau[2], Nz, Dz);
    const double *plane = lev + p.cell[2] * P.sxy + p.cell[1] * P.sx + p.cell[0];
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

template <int DIM, bool STAGED>
__device__ __forceinline__ void step(Point<DIM> &p, const double *lev, const BtParams &P, double d, double h)
{
    if constexpr (DIM == 1) step1d<STAGED>(p, lev, P, d, h);
    else if constexpr (DIM == 2) step2d<STAGED>(p, lev, P, d, h);
    else step3d<STAGED>(p, lev, P, d, h);
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
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

constexpr int kMaxStages = 16;
constexpr unsigned kBarBytes = 2 * kMaxStages * 8; // full[16], empty[16]

template <int DIM, int ILP> struct Tune
{
    // consumer warps per CTA upper bound (register budget), chosen from -Xptxas -v
    static constexpr int max_threads = DIM == 1 ? 1024 : (DIM == 2 ? (ILP == 1 ? 768 : 512) : (ILP == 1 ? 512 : 256));
};

// ---------------------------------------------------------------- the kernel
template <int DIM, int ILP, bool STAGED>
__global__ void __launch_bounds__(Tune<DIM, ILP>::max_threads, 1) backtrace_kernel(const __grid_constant__ BtParams P)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ double red[32][4];

    const int lane = threadIdx.x & 31;
    const unsigned warp = threadIdx.x >> 5;
    const unsigned W = P.W;
    const unsigned long long cta_first = static_cast<unsigned long long>(blockIdx.x) * W * P.rounds;
    if (cta_first >= P.n_units) return; // whole CTA idle (uniform)
    unsigned long long left = P.n_units - cta_first;
    const unsigned rounds = static_cast<unsigned>(min(static_cast<unsigned long long>(P.rounds), (left + W - 1) / W));

    unsigned long long *full = reinterpret_cast<unsigned long long *>(smem);
    unsigned long long *empty = full + kMaxStages;
    double *ring = reinterpret_cast<double *>(smem + kBarBytes);
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

    double m0 = 0, m1 = 0, m2 = 0, m3 = 0;

    if (STAGED && warp == W) {
        // ------------------------------------------------ producer warp: stream levels newest -> oldest, once per round
        if (lane == 0) {
            int s = 0;
            unsigned ph = 0;
            bool primed = false;
            for (unsigned r = 0; r < rounds; ++r)
                for (int m = P.first_level; m >= 0; --m) {
                    if (primed) mbar_wait(&empty[s], ph ^ 1u);
                    mbar_expect_tx(&full[s], P.level_bytes);
                    const unsigned char *src = reinterpret_cast<const unsigned char *>(P.hist + static_cast<unsigned long long>(m) * P.level_stride);
                    unsigned char *dst = reinterpret_cast<unsigned char *>(ring + static_cast<size_t>(s) * level_doubles);
                    for (unsigned off = 0; off < P.level_bytes; off += 32768u)
                        bulk_g2s(dst + off, src + off, min(32768u, P.level_bytes - off), &full[s]);
                    if (++s == P.stages) { s = 0; ph ^= 1u; primed = true; }
                }
        }
    } else {
        // ------------------------------------------------ consumer warps
        double acc = 0;
        long long cur_tile = -1;
        int s = 0;
        unsigned ph = 0;
        for (unsigned r = 0; r < rounds; ++r) {
            const unsigned long long unit = cta_first + static_cast<unsigned long long>(r) * W + warp;
            const bool unit_ok = unit < P.n_units;
            const unsigned long long tile = unit_ok ? unit / P.units_per_tile : 0ull;
            const unsigned long long jc = unit_ok ? unit % P.units_per_tile : 0ull;
            if (unit_ok && static_cast<long long>(tile) != cur_tile) {
                if (cur_tile >= 0 && !P.metrics) P.partials[((blockIdx.x + cur_tile) * W + warp) * 32 + lane] = acc;
                acc = 0;
                cur_tile = static_cast<long long>(tile);
            }
            unsigned long long l = P.l_first + tile * 32 + lane;
            const bool node_ok = unit_ok && l <= P.l_last;
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
            double v0[ILP][DIM]; // starting velocities (metrics need them)
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                unsigned long long j = jc * ILP + i;
                ok[i] = node_ok && j < P.Nvel;
                const unsigned long long q = l * P.Nvel + j;
                ok[i] = ok[i] && q >= P.q_begin && q < P.q_end;
                if (j >= P.Nvel) j = 0;
                const int iu = static_cast<int>(j % P.Nu);
                const int iv = DIM >= 2 ? static_cast<int>((j / P.Nu) % P.Nv) : 0;
                const int iw = DIM >= 3 ? static_cast<int>(j / (static_cast<unsigned long long>(P.Nu) * P.Nv)) : 0;
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

            for (int m = P.first_level; m >= 0; --m) {
                const double *lev;
                if constexpr (STAGED) {
                    mbar_wait(&full[s], ph);
                    lev = ring + static_cast<size_t>(s) * level_doubles;
                } else {
                    lev = P.hist + static_cast<unsigned long long>(m) * P.level_stride;
                }
                const bool first = P.metrics && m == P.first_level; // eval_f: half kick at the start, no drift
                const double h = (m == 0 || first) ? 0.5 : 1.0;
                const double d = first ? 0.0 : 1.0;
#pragma unroll
                for (int i = 0; i < ILP; ++i) step<DIM, STAGED>(pt[i], lev, P, d, h);
                if constexpr (STAGED) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[s]);
                    if (++s == P.stages) { s = 0; ph ^= 1u; }
                }
            }

#pragma unroll
            for (int i = 0; i < ILP; ++i) {
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
        if (cur_tile >= 0 && !P.metrics) P.partials[((blockIdx.x + cur_tile) * W + warp) * 32 + lane] = acc;
    }

    if (P.metrics) { // deterministic block reduction of the four metric sums
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m0 += __shfl_down_sync(0xffffffffu, m0, o);
            m1 += __shfl_down_sync(0xffffffffu, m1, o);
            m2 += __shfl_down_sync(0xffffffffu, m2, o);
            m3 += __shfl_down_sync(0xffffffffu, m3, o);
        }
        if (lane == 0) { red[warp][0] = m0; red[warp][1] = m1; red[warp][2] = m2; red[warp][3] = m3; }
        __syncthreads();
        if (threadIdx.x < 4) {
            double sum = 0;
            for (unsigned w = 0; w < W; ++w) sum += red[w][threadIdx.x];
            P.mpartials[blockIdx.x * 4 + threadIdx.x] = sum;
        }
    }
}

// Adds the per-(CTA,warp,tile) slots in a fixed order.  One warp per tile of 32 nodes.
__global__ void finish_rho_kernel(const __grid_constant__ FinishParams F)
{
    const unsigned tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (tile >= F.n_tiles) return;
    const unsigned long long per_cta = static_cast<unsigned long long>(F.W) * F.rounds;
    const unsigned long long u_lo = static_cast<unsigned long long>(tile) * F.units_per_tile;
    const unsigned long long u_hi = u_lo + F.units_per_tile - 1;
    const unsigned c_lo = static_cast<unsigned>(u_lo / per_cta);
    unsigned c_hi = static_cast<unsigned>(u_hi / per_cta);
    if (c_hi >= F.grid) c_hi = F.grid - 1;
    double sum = 0;
    for (unsigned c = c_lo; c <= c_hi; ++c) {
        const double *slot = F.partials + (static_cast<unsigned long long>(c + tile) * F.W) * 32 + lane;
        for (unsigned w = 0; w < F.W; ++w) sum += slot[static_cast<size_t>(w) * 32];
    }
    const unsigned long long l = F.l_first + static_cast<unsigned long long>(tile) * 32 + lane;
    if (l <= F.l_last) {
        F.rho_partial[l] = -F.dV * sum;
        if (F.rho_full) F.rho_full[l] = 1 - F.dV * sum;
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

template <int DIM, int ILP, bool STAGED>
cudaError_t launch_variant(const BtParams &P, unsigned grid, unsigned threads, size_t smem_bytes, cudaStream_t st)
{
    auto kern = backtrace_kernel<DIM, ILP, STAGED>;
    if (smem_bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes));
        if (e != cudaSuccess) return e;
    }
    kern<<<grid, threads, smem_bytes, st>>>(P);
    return cudaGetLastError();
}

template <int DIM>
cudaError_t launch_dim(const BtParams &P, int ilp, bool staged, unsigned grid, unsigned threads, size_t smem_bytes, cudaStream_t st)
{
    if (ilp == 2) {
        return staged ? launch_variant<DIM, 2, true>(P, grid, threads, smem_bytes, st)
                      : launch_variant<DIM, 2, false>(P, grid, threads, smem_bytes, st);
    }
    return staged ? launch_variant<DIM, 1, true>(P, grid, threads, smem_bytes, st)
                  : launch_variant<DIM, 1, false>(P, grid, threads, smem_bytes, st);
}

int max_threads_for(int dim, int ilp)
{
    if (dim == 1) return ilp == 1 ? Tune<1, 1>::max_threads : Tune<1, 2>::max_threads;
    if (dim == 2) return ilp == 1 ? Tune<2, 1>::max_threads : Tune<2, 2>::max_threads;
    return ilp == 1 ? Tune<3, 1>::max_threads : Tune<3, 2>::max_threads;
}

} // namespace

// Host side: decomposition + launch.  q is the reference's flat quadrature index (cuda_kernel.cu:40-41,
// 219-225, 402-412); [q_begin,q_end) may cut through a node's velocity range (masked per point).
int launch_backtrace(Handle *h, size_t n, size_t q_begin, size_t q_end, bool metrics)
{
    const nufi_b200_config3d &c = h->c;
    BtParams P{};
    P.dim = h->dim;
    P.Nx = static_cast<int>(c.Nx); P.Ny = static_cast<int>(c.Ny); P.Nz = static_cast<int>(c.Nz);
    P.Nu = static_cast<int>(c.Nu); P.Nv = static_cast<int>(c.Nv); P.Nw = static_cast<int>(c.Nw);
    P.sx = h->dim == 1 ? h->Nxp : h->sx;
    P.sxy = h->sxy;
    P.level_stride = h->level_stride;
    P.hist = h->d_hist;
    P.first_level = metrics ? (n == 0 ? -1 : static_cast<int>(n)) : static_cast<int>(n) - 1;
    P.metrics = metrics ? 1 : 0;
    P.cx = c.dt * c.dx_inv; P.cy = c.dt * c.dy_inv; P.cz = c.dt * c.dz_inv;
    const double scale = h->dim == 2 ? 12.0 : 72.0;
    P.gx = -c.dt * c.dx_inv / scale; P.gy = -c.dt * c.dy_inv / scale; P.gz = -c.dt * c.dz_inv / scale;
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
    P.q_begin = q_begin; P.q_end = q_end;
    P.l_first = q_begin / P.Nvel;
    P.l_last = (q_end - 1) / P.Nvel;
    const unsigned long long n_nodes_range = P.l_last - P.l_first + 1;
    P.n_tiles = static_cast<unsigned>((n_nodes_range + 31) / 32);
    P.level_bytes = static_cast<unsigned>(h->level_stride * 8);

    // ---- variant and shape
    const size_t ring_budget = h->smem_optin > kBarBytes + 1024 ? h->smem_optin - kBarBytes - 1024 : 0;
    bool staged = P.first_level >= 0 && 2ull * P.level_bytes <= ring_budget;
    if (h->variant_force == 1) staged = false;
    if (h->variant_force == 2 && 2ull * P.level_bytes > ring_budget)
        return fail(h, NUFI_B200_ERR_ARG, "staged variant forced but two levels do not fit in shared memory");
    const unsigned grid = static_cast<unsigned>(h->sm_count);
    const unsigned long long warp_tasks = static_cast<unsigned long long>(P.n_tiles) * P.Nvel;
    // two points per thread once there is more than one resident warp-task per consumer warp
    int ilp = 1;
    {
        const unsigned wmax1 = max_threads_for(h->dim, 1) / 32 - (staged ? 1 : 0);
        if (warp_tasks > static_cast<unsigned long long>(grid) * wmax1) ilp = 2;
    }
    const unsigned wmax = max_threads_for(h->dim, ilp) / 32 - (staged ? 1 : 0);
    P.units_per_tile = (P.Nvel + ilp - 1) / ilp;
    P.n_units = P.units_per_tile * P.n_tiles;
    unsigned long long w_need = (P.n_units + grid - 1) / grid;
    P.W = static_cast<unsigned>(w_need < wmax ? (w_need < 1 ? 1 : w_need) : wmax);
    P.rounds = static_cast<unsigned>((P.n_units + static_cast<unsigned long long>(grid) * P.W - 1) / (static_cast<unsigned long long>(grid) * P.W));
    if (P.rounds == 0) P.rounds = 1;
    // rebalance: spread the same number of rounds over as few warps as needed
    P.W = static_cast<unsigned>((P.n_units + static_cast<unsigned long long>(grid) * P.rounds - 1) / (static_cast<unsigned long long>(grid) * P.rounds));
    if (P.W < 1) P.W = 1;
    const unsigned threads = (P.W + (staged ? 1 : 0)) * 32;
    size_t smem_bytes = 0;
    if (staged) {
        int stages = static_cast<int>(ring_budget / P.level_bytes);
        if (stages > kMaxStages) stages = kMaxStages;
        if (stages > P.first_level + 1) stages = P.first_level + 1 > 2 ? P.first_level + 1 : 2;
        P.stages = stages;
        smem_bytes = kBarBytes + static_cast<size_t>(stages) * P.level_bytes;
    }

    // ---- partial slots
    if (!metrics) {
        const size_t need = (static_cast<size_t>(grid) + P.n_tiles) * P.W * 32;
        if (need > h->partials_cap) {
            if (h->d_partials) cudaFree(h->d_partials);
            h->d_partials = nullptr;
            h->partials_cap = 0;
            if (cudaMalloc(&h->d_partials, need * sizeof(double)) != cudaSuccess)
                return fail(h, NUFI_B200_ERR_ALLOC, "cudaMalloc of the rho partial slots failed");
            h->partials_cap = need;
        }
        NUFI_CUDA_CHECK(h, cudaMemsetAsync(h->d_partials, 0, need * sizeof(double), h->stream));
        P.partials = h->d_partials;
    } else {
        P.mpartials = h->d_mpartials;
        NUFI_CUDA_CHECK(h, cudaMemsetAsync(h->d_mpartials, 0, sizeof(double) * 4 * grid, h->stream));
    }

    cudaEvent_t ev_start, ev_stop;
    {
        int rc = ev_acquire(h, &ev_start, &ev_stop);
        if (rc) return rc;
    }
    NUFI_CUDA_CHECK(h, cudaEventRecord(ev_start, h->stream));
    cudaError_t e;
    if (h->dim == 1) e = launch_dim<1>(P, ilp, staged, grid, threads, smem_bytes, h->stream);
    else if (h->dim == 2) e = launch_dim<2>(P, ilp, staged, grid, threads, smem_bytes, h->stream);
    else e = launch_dim<3>(P, ilp, staged, grid, threads, smem_bytes, h->stream);
    NUFI_CUDA_CHECK(h, e);
    NUFI_CUDA_CHECK(h, cudaEventRecord(ev_stop, h->stream));
    h->ev_pending += 1;
    h->launches += 1;
    h->last_variant = staged ? (ilp == 2 ? "smem-tma/ilp2" : "smem-tma/ilp1") : (ilp == 2 ? "global/ilp2" : "global/ilp1");

    if (!metrics) {
        FinishParams F{};
        F.partials = h->d_partials;
        F.rho_partial = h->d_rho_partial;
        const bool whole = q_begin == 0 && q_end == h->n_nodes * h->n_vel;
        F.rho_full = whole ? h->d_rho_full : nullptr;
        F.dV = h->dim == 1 ? P.du : (h->dim == 2 ? P.du * P.dv : P.du * P.dv * P.dw); // rho.hpp:145, 307, 459
        F.l_first = P.l_first; F.l_last = P.l_last; F.n_nodes_total = h->n_nodes;
        F.units_per_tile = P.units_per_tile;
        F.n_tiles = P.n_tiles; F.rounds = P.rounds; F.W = P.W; F.grid = grid;
        if (!whole) NUFI_CUDA_CHECK(h, cudaMemsetAsync(h->d_rho_partial, 0, sizeof(double) * h->n_nodes, h->stream));
        const unsigned fb = 128, tiles_per_block = fb / 32;
        finish_rho_kernel<<<(P.n_tiles + tiles_per_block - 1) / tiles_per_block, fb, 0, h->stream>>>(F);
        NUFI_CUDA_CHECK(h, cudaGetLastError());
    } else {
        finish_metrics_kernel<<<1, 32, 0, h->stream>>>(h->d_mpartials, grid, h->d_metrics);
        NUFI_CUDA_CHECK(h, cudaGetLastError());
    }
    h->launches += 1;
    return NUFI_B200_OK;
}

} // namespace nufi_b200
