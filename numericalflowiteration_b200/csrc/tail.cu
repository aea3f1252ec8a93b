// tail.cu -- the per-step field tail on the device: periodic Poisson solve + cubic-spline interpolation,
// and the conversion between the reference's level layout and the device level format.
//
// Replaces the host stages of the reference loop (bin/test_nufi_gpu_3d.cpp:156-162):
//   poisson<double>::solve  nufi/poisson.cpp:66-89, 190-219, 328-362   (FFTW DHT pair)
//   interpolate<double,4>   nufi/fields.hpp:63-142, 186-300, 352-490    (LSMR on the collocation system)
// Both are diagonal in Fourier space: with rho^ the DFT of rho,
//   c^_k = rho^_k / (N |kappa(k)|^2 prod_d lambda_d(k_d)),  lambda_d(k) = sum_i N_i(0) exp(+2 pi i i k / N_d),
//   kappa_d(k) = 2 pi min(k, N_d-k) / L_d,  c^_0 = 0,
//   energy = (V/2) sum_{k != 0} |kappa|^2 |rho^_k / (N |kappa|^2)|^2     (poisson.cpp:74-87, 204-217, 344-360),
// so one D2Z FFT, one pointwise kernel and one Z2D FFT replace two DHTs plus an iterative solve.
#include "internal.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace nufi_b200
{

namespace
{

struct TailParams
{
    int dim, Nx, Ny, Nz, Nxh;       // Nxh = Nx/2 + 1
    const double *kap2x, *kap2y, *kap2z; // per-dimension ii*ii*fac*fac
    const double2 *ilx, *ily, *ilz;      // per-dimension 1/lambda
    double fac_N;                   // 1/(Nx Ny Nz)
    double vol_half;                // Lx Ly Lz / 2
    int mode;                       // bit 0: Poisson symbol 1/|kappa|^2 (zero mean mode), bit 1: collocation symbol 1/lambda
};

// spec <- spec * symbol;  per-block partial of sum_k w_k |kappa|^2 |phi^_k|^2  (w = 1 on the self-conjugate
// x planes, 2 otherwise: the half spectrum stands for the full one).
__global__ void symbol_kernel(cufftDoubleComplex *spec, TailParams T, double *epart)
{
    __shared__ double red[32];
    const size_t n_spec = static_cast<size_t>(T.Nxh) * T.Ny * T.Nz;
    double e = 0;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < n_spec;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int kx = static_cast<int>(idx % T.Nxh);
        const size_t rest = idx / T.Nxh;
        const int ky = static_cast<int>(rest % T.Ny);
        const int kz = static_cast<int>(rest / T.Ny);
        cufftDoubleComplex F = spec[idx];
        double pr, pi;
        if (T.mode & 1) {
            if (kx == 0 && ky == 0 && kz == 0) {
                spec[idx] = make_cuDoubleComplex(0.0, 0.0); // data[0] = 0 (poisson.cpp:83, 214, 357)
                continue;
            }
            const double kap2 = T.kap2x[kx] + T.kap2y[ky] + T.kap2z[kz];
            const double fac = T.fac_N / kap2;
            pr = F.x * fac; pi = F.y * fac; // phi^_k / N-normalised
            const double w = (kx == 0 || 2 * kx == T.Nx) ? 1.0 : 2.0;
            e += w * kap2 * (pr * pr + pi * pi);
        } else {
            pr = F.x * T.fac_N; pi = F.y * T.fac_N;
        }
        if (!(T.mode & 2)) {
            spec[idx] = make_cuDoubleComplex(pr, pi);
            continue;
        }
        // divide by the collocation symbol, dimension by dimension
        double2 a = T.ilx[kx];
        double2 b = T.ily[ky];
        double2 c = T.ilz[kz];
        double2 ab = make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
        double2 s = make_double2(ab.x * c.x - ab.y * c.y, ab.x * c.y + ab.y * c.x);
        spec[idx] = make_cuDoubleComplex(pr * s.x - pi * s.y, pr * s.y + pi * s.x);
    }
    // deterministic block reduction
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_down_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (unsigned w = 0; w < (blockDim.x >> 5); ++w) s += red[w];
        epart[blockIdx.x] = s;
    }
}

// 1d pp-form of one cell: dt*E(tau) = p0 + p1 tau + p2 tau^2 on cell k, from the spline coefficients
// c[k..k+3] (derivative basis of nufi/splines.hpp re-centred at the cell midpoint, tau = t - 1/2).
__device__ __forceinline__ void cell_poly_1d(const double *c, double g, double &p0, double &p1, double &p2)
{
    const double c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3];
    p0 = g * (0.125 * (c3 - c0) + 0.625 * (c2 - c1));
    p1 = g * (0.5 * ((c0 - c1) + (c3 - c2)));
    p2 = g * (0.5 * ((c3 - c0) + 3.0 * (c1 - c2)));
}

// x-direction pp-form of one row-cell (2d/3d "xpp" level format): A(tau) = sum_a c_a 6 N_a(1/2 + tau)
//   = a0 + a1 tau + a2 tau^2 + a3 tau^3  (6 N_0 = (1/2 - tau)^3, 6 N_1 = 23/8 - 15/4 tau - 3/2 tau^2 + 3 tau^3, N_2(tau) = N_1(-tau),
//   N_3(tau) = N_0(-tau)); the x-derivative contraction sum_a c_a 2 N'_a is A'(tau)/3.
__device__ __forceinline__ void row_poly(double c0, double c1, double c2, double c3, double2 &a01, double2 &a23)
{
    const double s03 = c0 + c3, s12 = c1 + c2, d30 = c3 - c0, d21 = c2 - c1;
    a01.x = fma(2.875, s12, 0.125 * s03);
    a01.y = fma(3.75, d21, 0.75 * d30);
    a23.x = 1.5 * (s03 - s12);
    a23.y = fma(-3.0, d21, d30);
}

struct ExpandParams
{
    int dim, Nx, Ny, Nz, sx, sxy;
    size_t level_doubles; // device level size
    double g1;            // 1d: -dt*dx_inv
    int xpp;              // 2d/3d: level = per (row, cell) cubic in the x offset, [Q01: Nx x (a0,a1)][Q23: Nx x (a2,a3)] per row
    int halo;             // order - 1 periodic halo nodes per dimension in the reference layout (3 for the cubic formats above)
    int pp1d;             // 1d, order 4: level = per-cell quadratics of dt*E, raw level kept beside it; else level = raw layout
};

// xpp level + raw reference-format level from a coefficient source.  PERIODIC: src holds Nx*Ny*Nz periodic coefficients;
// otherwise src is a reference-format level with halo (row stride Nx+3).  One thread per (row, cell) / raw element.
template <bool PERIODIC, typename Src>
__device__ __forceinline__ void write_xpp(const Src &src, double *level, double *raw, const ExpandParams &E, size_t tid, size_t nthreads)
{
    const int rx = E.Nx + 3, ry = E.Ny + 3, rz = E.dim == 3 ? E.Nz + 3 : 1;
    auto coef = [&](int k, int j, int i) -> double {
        if (PERIODIC) { // every index is below 2 N (halo of 3, N >= 4): one conditional subtraction wraps it
            const int kw = k >= E.Nz ? k - E.Nz : k, jw = j >= E.Ny ? j - E.Ny : j, iw = i >= E.Nx ? i - E.Nx : i;
            return src((static_cast<size_t>(kw) * E.Ny + jw) * E.Nx + iw);
        }
        return src((static_cast<size_t>(k) * ry + j) * rx + i);
    };
    const size_t n_raw = static_cast<size_t>(rx) * ry * rz;
    for (size_t idx = tid; idx < n_raw; idx += nthreads) {
        const int i = static_cast<int>(idx % rx);
        const size_t rest = idx / rx;
        raw[idx] = coef(static_cast<int>(rest / ry), static_cast<int>(rest % ry), i);
    }
    const size_t n_cells = static_cast<size_t>(E.Nx) * ry * rz;
    double2 *lv = reinterpret_cast<double2 *>(level);
    for (size_t idx = tid; idx < n_cells; idx += nthreads) {
        const int i = static_cast<int>(idx % E.Nx);
        const size_t r = idx / E.Nx;
        const int j = static_cast<int>(r % ry), k = static_cast<int>(r / ry);
        double2 a01, a23;
        row_poly(coef(k, j, i), coef(k, j, i + 1), coef(k, j, i + 2), coef(k, j, i + 3), a01, a23);
        lv[r * 2 * E.Nx + i] = a01;
        lv[r * 2 * E.Nx + E.Nx + i] = a23;
    }
}

__global__ void expand_xpp_kernel(const double *src, double *level, double *raw, ExpandParams E, int periodic, const double *epart,
                                  unsigned n_epart, double vol_half, double *energy_out)
{
    const size_t tid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x, nth = static_cast<size_t>(gridDim.x) * blockDim.x;
    auto get = [&](size_t i) { return src[i]; };
    if (periodic) write_xpp<true>(get, level, raw, E, tid, nth);
    else write_xpp<false>(get, level, raw, E, tid, nth);
    if (energy_out && blockIdx.x == 0 && threadIdx.x == 0) {
        double s = 0;
        for (unsigned b = 0; b < n_epart; ++b) s += epart[b];
        *energy_out = s * vol_half;
    }
}

// Periodic coefficients (Nx*Ny*Nz, x fastest) -> device level with (order-1) halo
// (the copy loops of nufi/fields.hpp:140-141, 294-299, 482-489), 2d/3d.
__global__ void expand_kernel(const double *src, double *level, ExpandParams E, double final_scale_unused,
                              const double *epart, unsigned n_epart, double vol_half, double *energy_out)
{
    (void)final_scale_unused;
    const size_t total = E.level_doubles;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int i = static_cast<int>(idx % E.sx);
        const size_t rest = idx / E.sx;
        const int rows = E.dim >= 2 ? E.Ny + E.halo : 1;
        const int j = static_cast<int>(rest % rows);
        const int k = static_cast<int>(rest / rows);
        double v = 0;
        if (i < E.Nx + E.halo && (E.dim < 3 ? k == 0 : k < E.Nz + E.halo))
            v = src[(static_cast<size_t>(k % E.Nz) * E.Ny + (j % E.Ny)) * E.Nx + (i % E.Nx)];
        level[idx] = v;
    }
    if (energy_out && blockIdx.x == 0 && threadIdx.x == 0) {
        double s = 0;
        for (unsigned b = 0; b < n_epart; ++b) s += epart[b];
        *energy_out = s * vol_half; // energy *= Lx*Ly*Lz/2 (poisson.cpp:87, 217, 360)
    }
}

// 1d: raw level with halo + per-cell quadratics p0 + p1 tau + p2 tau^2, stored as [Nx x (p1, p2)] [Nx x p0]: one 128-bit and one
// 64-bit shared-memory load per point-step, both conflict-free for a warp's consecutive cells.
__global__ void expand1d_kernel(const double *src, double *raw, double *pp, ExpandParams E, const double *epart,
                                unsigned n_epart, double vol_half, double *energy_out)
{
    const int Nx = E.Nx;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Nx + 3; i += gridDim.x * blockDim.x) raw[i] = src[i % Nx];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < Nx; k += gridDim.x * blockDim.x) {
        double p0, p1, p2;
        double c[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) c[a] = src[(k + a) % Nx];
        cell_poly_1d(c, E.g1, p0, p1, p2);
        pp[2 * k] = p1; // level = [Nx x (p1, p2)] [Nx x p0]
        pp[2 * k + 1] = p2;
        pp[2 * Nx + k] = p0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && (3 * Nx) % 2) pp[3 * Nx] = 0; // padding double
    if (energy_out && blockIdx.x == 0 && threadIdx.x == 0) {
        double s = 0;
        for (unsigned b = 0; b < n_epart; ++b) s += epart[b];
        *energy_out = s * vol_half;
    }
}

// reference-format level (with halo, row stride Nx+3) -> device format
__global__ void ref_to_device_kernel(const double *ref, double *level, double *raw1d, ExpandParams E)
{
    if (E.pp1d) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < E.Nx + 3; i += gridDim.x * blockDim.x) raw1d[i] = ref[i];
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < E.Nx; k += gridDim.x * blockDim.x) {
            double p0, p1, p2;
            cell_poly_1d(ref + k, E.g1, p0, p1, p2);
            level[2 * k] = p1;
            level[2 * k + 1] = p2;
            level[2 * E.Nx + k] = p0;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0 && (3 * E.Nx) % 2) level[3 * E.Nx] = 0;
        return;
    }
    const size_t total = E.level_doubles;
    const int rx = E.Nx + E.halo, ry = E.dim >= 2 ? E.Ny + E.halo : 1;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int i = static_cast<int>(idx % E.sx);
        const size_t rest = idx / E.sx;
        const int j = static_cast<int>(rest % ry);
        const int k = static_cast<int>(rest / ry);
        double v = 0;
        if (i < rx && (E.dim < 3 ? k == 0 : k < E.Nz + E.halo)) v = ref[(static_cast<size_t>(k) * ry + j) * rx + i];
        level[idx] = v;
    }
}

__global__ void device_to_ref_kernel(const double *level, const double *raw1d, double *ref, ExpandParams E)
{
    if (E.pp1d || E.xpp) { // the raw reference-format level is kept beside the pp-form
        const size_t total = static_cast<size_t>(E.Nx + 3) * (E.dim >= 2 ? E.Ny + 3 : 1) * (E.dim >= 3 ? E.Nz + 3 : 1);
        for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x)
            ref[i] = raw1d[i];
        return;
    }
    const int rx = E.Nx + E.halo, ry = E.dim >= 2 ? E.Ny + E.halo : 1, rz = E.dim == 3 ? E.Nz + E.halo : 1;
    const size_t total = static_cast<size_t>(rx) * ry * rz;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int i = static_cast<int>(idx % rx);
        const size_t rest = idx / rx;
        const int j = static_cast<int>(rest % ry);
        const int k = static_cast<int>(rest / ry);
        ref[idx] = level[(static_cast<size_t>(k) * ry + j) * E.sx + i];
    }
}

__global__ void full_rho_kernel(const double *partial_sum, double *full, size_t n)
{
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        full[i] = 1 + partial_sum[i];
}


// ---------------------------------------------------------------------------------------------------------
// Fused small-grid tail: ONE CTA does  [slot reduction ->] rho -> DFT -> Poisson x spline symbol -> inverse DFT ->
// level n with halo (1d: raw level + per-cell quadratics) + energy.  For the grids the reference's drivers run
// (256 / 32^2 / 8^3 ... 16^3 nodes) the four library launches of the cuFFT path cost more than the arithmetic; here
// the transforms are direct separable sums over shared memory with exact twiddle tables (O(N * N_d), N <= 4096).
// ---------------------------------------------------------------------------------------------------------
constexpr int kSmallMaxNodes = 4096;
constexpr int kSmallMaxDim = 256;
constexpr int kSmallThreads = 1024;
constexpr int kSmallPart = 4096; // double2 scratch: DFT partial sums (<= kSmallThreads) / slot partials (8 x N doubles, N < 1024)

struct SmallTailParams
{
    TailParams T;
    ExpandParams E;
    const double2 *twx, *twy, *twz; // (cos, sin)(2 pi m / N_d)
    const double *rho;              // CPU-convention rho on the device, or nullptr: reduce the backtrace slots (F)
    FinishParams F;
    PeerRecv X;                     // X.world > 0 (multi-GPU step): rho = the per-node sums of ALL ranks, from the peer exchange buffer
    PeerPush XP;                    // XP.world > 0: this rank's sums (from its slots, F) go to the peers first; 0: the words are all there
    double *level, *raw1d, *energy_out;
};

// one separable pass along a dimension of length Nd (stride `stride`):  B[o] = sum_j A[base + j*stride] * w^(+-j*k).
// When the block has S = blockDim/N >= 2 threads per output the j-range is cut into S parts whose partial sums are
// combined in a fixed order through `part` (S*N entries); two accumulator pairs break the FMA dependency chain.
// (The transforms are NOT inlined and take the direction as a run-time sign: the one-CTA tail executes its code once per launch,
// from a cold instruction cache, and six inlined copies of them (3 dimensions x 2 directions) made the kernel 10.6 k instructions;
// one copy: 6.5 k, the inverse transform runs on code the forward one has just fetched.  Measured: no slower anywhere, C4's step
// 0.7 % faster.  sg = -1 forward, +1 inverse; multiplying by +-1 is exact, so the results are bit-identical to templated versions.)
__device__ __noinline__ void dft_pass(const double2 *A, double2 *B, double2 *part, int N, int Nd, int stride, const double2 *tw, double sg)
{
    int S = 1;
    while (2 * S * N <= static_cast<int>(blockDim.x) && 2 * S <= 8 && Nd % (2 * S) == 0) S *= 2;
    const int len = Nd / S;
    for (int t = threadIdx.x; t < S * N; t += blockDim.x) {
        const int o = t % N, p = t / N;
        const int k = (o / stride) % Nd;
        const int base = o - k * stride;
        const int j0 = p * len;
        int idx = static_cast<int>((static_cast<long long>(j0) * k) % Nd);
        double sr0 = 0, si0 = 0, sr1 = 0, si1 = 0;
        auto term = [&](int j, double &sr, double &si) {
            const double2 a = A[base + j * stride];
            const double2 w = tw[idx];
            const double t = -sg * w.y; // forward: a * (c - i s); inverse: a * (c + i s)
            sr = fma(a.x, w.x, fma(a.y, t, sr));
            si = fma(a.y, w.x, fma(-a.x, t, si));
            idx += k;
            if (idx >= Nd) idx -= Nd;
        };
        int j = j0;
#pragma unroll 2
        for (; j + 1 < j0 + len; j += 2) {
            term(j, sr0, si0);
            term(j + 1, sr1, si1);
        }
        if (j < j0 + len) term(j, sr0, si0);
        const double2 r = make_double2(sr0 + sr1, si0 + si1);
        if (S == 1) B[o] = r;
        else part[p * N + o] = r;
    }
    if (S > 1) {
        __syncthreads();
        for (int o = threadIdx.x; o < N; o += blockDim.x) {
            double2 r = part[o];
            for (int p = 1; p < S; ++p) { r.x += part[p * N + o].x; r.y += part[p * N + o].y; }
            B[o] = r;
        }
    }
}

// u * exp(sg * i * angle), w = (cos, sin)(angle)
__device__ __forceinline__ double2 cmul_tw(double2 u, double2 w, double sg)
{
    const double t = -sg * w.y;
    return make_double2(fma(u.x, w.x, u.y * t), fma(u.y, w.x, -(u.x * t)));
}

// Stockham FFT along a power-of-two dimension, all lines of the grid at once, ping-pong between the two buffers, twiddles from
// the exact table: one radix-2 stage if log2(Nd) is odd, then radix-4 stages (N/4 butterflies each) -- half the stages and
// block-wide barriers of a pure radix-2 scheme; the tail is latency-bound, not work-bound.  Returns with the result in `cur`.
__device__ __noinline__ void fft_pass(double2 *&cur, double2 *&oth, int N, int Nd, int stride, const double2 *tw, double sg)
{
    const int lg = 31 - __clz(Nd);
    int Ns = 1, ls = 0; // current sub-transform length and its log2
    if (lg & 1) {       // radix-2 stage with Ns = 1: all twiddles are 1
        const int half = Nd >> 1, lh = lg - 1;
        for (int t = threadIdx.x; t < (N >> 1); t += blockDim.x) {
            int lo = 0, q = t;
            if (stride != 1) { q = t / stride; lo = t - q * stride; }
            const int j = q & (half - 1);
            const int base = (q >> lh) * stride * Nd + lo;
            const double2 v0 = cur[base + j * stride];
            const double2 v1 = cur[base + (j + half) * stride];
            oth[base + (2 * j) * stride] = make_double2(v0.x + v1.x, v0.y + v1.y);
            oth[base + (2 * j + 1) * stride] = make_double2(v0.x - v1.x, v0.y - v1.y);
        }
        __syncthreads();
        double2 *tmp = cur; cur = oth; oth = tmp;
        Ns = 2; ls = 1;
    }
    const int quarter = Nd >> 2, lq = lg - 2;
    for (; Ns < Nd; Ns <<= 2, ls += 2) {
        const int tshift = lq - ls; // twiddle index step = Nd / (4 Ns)
        for (int t = threadIdx.x; t < (N >> 2); t += blockDim.x) {
            int lo = 0, q = t;
            if (stride != 1) { q = t / stride; lo = t - q * stride; }
            const int j = q & (quarter - 1);
            const int base = (q >> lq) * stride * Nd + lo;
            const int k = j & (Ns - 1);
            const int m = k << tshift;
            const double2 v0 = cur[base + j * stride];
            const double2 v1 = cmul_tw(cur[base + (j + quarter) * stride], tw[m], sg);
            const double2 v2 = cmul_tw(cur[base + (j + 2 * quarter) * stride], tw[2 * m], sg);
            const double2 v3 = cmul_tw(cur[base + (j + 3 * quarter) * stride], tw[3 * m], sg);
            const double2 a = make_double2(v0.x + v2.x, v0.y + v2.y), b = make_double2(v0.x - v2.x, v0.y - v2.y);
            const double2 c = make_double2(v1.x + v3.x, v1.y + v3.y), d = make_double2(v1.x - v3.x, v1.y - v3.y);
            // i * d = (-d.y, d.x); forward (sg = -1): y1 = b - i d, y3 = b + i d; inverse: the other way round
            const double2 id = make_double2(-sg * d.y, sg * d.x);
            const int o = (((j >> ls) << (ls + 2)) + k) * stride + base;
            oth[o] = make_double2(a.x + c.x, a.y + c.y);
            oth[o + Ns * stride] = make_double2(b.x + id.x, b.y + id.y);
            oth[o + 2 * Ns * stride] = make_double2(a.x - c.x, a.y - c.y);
            oth[o + 3 * Ns * stride] = make_double2(b.x - id.x, b.y - id.y);
        }
        __syncthreads();
        double2 *tmp = cur; cur = oth; oth = tmp;
    }
}

// separable transform, dimension by dimension: FFT when a length is a power of two (>= 8), direct sums otherwise
__device__ __noinline__ void transform_all(double2 *&cur, double2 *&oth, double2 *part, int dim, int Nx, int Ny, int Nz, const double2 *twx,
                                           const double2 *twy, const double2 *twz, double sg)
{
    const int N = Nx * Ny * Nz;
    for (int d = 0; d < dim; ++d) {
        const int Nd = d == 0 ? Nx : (d == 1 ? Ny : Nz);
        const int stride = d == 0 ? 1 : (d == 1 ? Nx : Nx * Ny);
        const double2 *tw = d == 0 ? twx : (d == 1 ? twy : twz);
        if (Nd >= 8 && (Nd & (Nd - 1)) == 0) {
            fft_pass(cur, oth, N, Nd, stride, tw, sg);
        } else {
            dft_pass(cur, oth, part, N, Nd, stride, tw, sg);
            __syncthreads();
            double2 *tmp = cur; cur = oth; oth = tmp;
        }
    }
}

#ifdef NUFI_TAIL_TIMING
#define TAIL_MARK(i) do { __syncthreads(); if (threadIdx.x == 0) tmark[i] = clock64(); } while (0)
#else
#define TAIL_MARK(i) do { } while (0)
#endif

__global__ void __launch_bounds__(kSmallThreads, 1) tail_small_kernel(const __grid_constant__ SmallTailParams S)
{
#ifdef NUFI_TAIL_TIMING
    __shared__ long long tmark[24];
#endif
    TAIL_MARK(0);
    extern __shared__ __align__(16) unsigned char sm_raw[];
    __shared__ double red[32];
    const TailParams &T = S.T;
    const int Nx = T.Nx, Ny = T.Ny, Nz = T.Nz;
    const int N = Nx * Ny * Nz;
    double2 *A = reinterpret_cast<double2 *>(sm_raw);
    double2 *B = A + N;
    double2 *twx = B + N, *twy = twx + Nx, *twz = twy + Ny;
    double2 *ilx = twz + Nz, *ily = ilx + Nx, *ilz = ily + Ny;
    double *kap2x = reinterpret_cast<double *>(ilz + Nz), *kap2y = kap2x + Nx, *kap2z = kap2y + Ny;
    double2 *part = reinterpret_cast<double2 *>(kap2x + ((Nx + Ny + Nz + 1) & ~1)); // kSmallPart entries

    // ---- tables into shared memory first (their global-memory latency overlaps the slot reduction)
    for (int i = threadIdx.x; i < Nx + Ny + Nz; i += blockDim.x) {
        twx[i] = S.twx[i]; // the three tables of a kind are contiguous
        ilx[i] = T.ilx[i];
        kap2x[i] = T.kap2x[i];
    }
    // slot bookkeeping without integer divisions in the load loops: first tile of every backtrace CTA and the CTA range of every
    // tile, tabulated here (one division per thread, hidden behind the backtrace kernel by the programmatic launch)
    __shared__ unsigned short s_tfirst[256], s_blo[128], s_bhi[128];
    const bool from_slots = !S.rho && S.F.n_tiles > 0;
    const bool tabulated = from_slots && S.F.n_tiles <= 128 && S.F.rpc > 0 && (S.F.rpt * S.F.n_tiles + S.F.rpc - 1) / S.F.rpc <= 256;
    if (tabulated) {
        const FinishParams &F = S.F;
        const unsigned n_ctas = (F.rpt * F.n_tiles + F.rpc - 1) / F.rpc;
        for (unsigned b = threadIdx.x; b < n_ctas; b += blockDim.x) s_tfirst[b] = static_cast<unsigned short>((b * F.rpc) / F.rpt);
        for (unsigned t = threadIdx.x; t < F.n_tiles; t += blockDim.x) {
            s_blo[t] = static_cast<unsigned short>((t * F.rpt) / F.rpc);
            s_bhi[t] = static_cast<unsigned short>(((t + 1) * F.rpt - 1) / F.rpc);
        }
    }
    TAIL_MARK(8);
    // Everything below reads what the kernels ahead of this one on the stream wrote -- unless all of it comes as self-validating
    // words (slots of a fused step, the peers' sums of a multi-GPU step; internal.cuh): then the tail starts adding as soon as
    // the words land instead of waiting for the backtrace grid to drain its stores and retire.  (Level n, rho and the energy
    // slot written below are touched by no kernel still in flight.)
    const bool polled = !S.rho && (!from_slots || S.F.slots_ll != nullptr);
    if (!polled) pdl_wait();
    __syncthreads(); // the slot tables above are complete
    TAIL_MARK(9);
    // ---- rho (either given, or the fixed-order sum of the backtrace kernel's per-(CTA, tile) slots: per node 8 strided
    //      partial sums over the CTAs that touched its tile, then added in order -- the association of finish_rho_kernel)
    if (S.rho) {
        for (int l = threadIdx.x; l < N; l += blockDim.x) A[l] = make_double2(S.rho[l], 0.0);
    } else if (!from_slots) { // multi-GPU step of a rank without a share of its own: every rank's sums are in the exchange buffer
        for (int l = threadIdx.x; l < N; l += blockDim.x) {
            const double r = fma(-S.X.dV, peer_rank_sum(S.X.rho + l, S.X.n_nodes, S.X.world, S.X.flag, S.X.status, -1, 0.0), 1.0);
            S.X.rho_full[l] = r;
            A[l] = make_double2(r, 0.0);
        }
        TAIL_MARK(10);
    } else {
        const FinishParams &F = S.F;
        auto slot_sum = [&](unsigned tile, unsigned lane, unsigned b_lo, unsigned b_hi, int w) {
            double sum = 0, lo = 0; // compensated (two-sum): the partial sum is exact to one rounding
            for (unsigned b = b_lo + w; b <= b_hi; b += 32) { // four loads in flight per trip; the order is kept
                double v[4];
                size_t at[4]; // slot index, or ~0: beyond the tile's last CTA
#pragma unroll
                for (unsigned u = 0; u < 4; ++u) {
                    const unsigned bb = b + 8 * u;
                    at[u] = ~static_cast<size_t>(0);
                    if (bb <= b_hi) {
                        const unsigned t_first = tabulated ? s_tfirst[bb] : (bb * F.rpc) / F.rpt;
                        at[u] = (static_cast<size_t>(bb) * F.Tmax + (tile - t_first)) * 32 + lane;
                    }
                }
                if (F.slots_ll) { // self-validating slots: poll until they carry this launch's epoch
                    const uint4 *src[4];
#pragma unroll
                    for (unsigned u = 0; u < 4; ++u) src[u] = at[u] != ~static_cast<size_t>(0) ? F.slots_ll + at[u] : nullptr;
                    ll_load4(src, F.slot_flag, F.status, v);
                } else {
#pragma unroll
                    for (unsigned u = 0; u < 4; ++u) v[u] = at[u] != ~static_cast<size_t>(0) ? F.slots[at[u]] : 0.0;
                }
#pragma unroll
                for (unsigned u = 0; u < 4; ++u) two_sum(sum, lo, v[u]);
            }
            return sum + lo;
        };
        auto store_rho = [&](int l, double tot) {
            F.rho_partial[l] = -F.dV * tot;
            if (S.X.world) { // multi-GPU step: the all-reduce happens here.  This rank's sum goes straight into the other GPUs'
                // exchange buffers (NVLink stores); then all ranks' sums are added in rank order (own from the register, the
                // others polled until they carry this step's epoch) -- the same association, hence the same bits, on every GPU
                for (int p = 0; p < S.XP.world; ++p)
                    if (p != S.XP.rank) peer_store_double(S.XP.rho[p] + l, tot, S.XP.flag);
                tot = peer_rank_sum(S.X.rho + l, S.X.n_nodes, S.X.world, S.X.flag, S.X.status, S.XP.rank, tot);
                const double r = fma(-S.X.dV, tot, 1.0);
                S.X.rho_full[l] = r;
                A[l] = make_double2(r, 0.0);
                return;
            }
            // rho = 1 - dV * sum with ONE rounding, of the small result (the reference rounds the O(1) product first: where the
            // density perturbation has decayed to 1e-6 that rounding alone is 1e-10 of it)
            const double r = fma(-F.dV, tot, 1.0);
            if (F.rho_full) F.rho_full[l] = r;
            A[l] = make_double2(r, 0.0);
        };
        if (8 * N <= 2 * kSmallPart) { // small grids: spread the 8 partial sums of a node over up to 8 threads
            int G = 1;                 // threads per node
            while (2 * G * N <= static_cast<int>(blockDim.x) && G < 8) G *= 2;
            double *ps = reinterpret_cast<double *>(part); // [8][N]
            const unsigned tn_log2 = 31 - __clz(F.TN); // TN is a power of two
            for (int it = threadIdx.x; it < N * G; it += blockDim.x) {
                int l = it, g = 0;
                while (l >= N) { l -= N; ++g; } // g < G <= 8
                const unsigned tile = static_cast<unsigned>(l) >> tn_log2, lane = static_cast<unsigned>(l) & (F.TN - 1);
                const unsigned b_lo = tabulated ? s_blo[tile] : (tile * F.rpt) / F.rpc;
                const unsigned b_hi = tabulated ? s_bhi[tile] : ((tile + 1) * F.rpt - 1) / F.rpc;
                for (int w = g; w < 8; w += G) ps[w * N + l] = slot_sum(tile, lane, b_lo, b_hi, w);
            }
            __syncthreads();
            TAIL_MARK(10);
            for (int l = threadIdx.x; l < N; l += blockDim.x) {
                double tot = 0, lo = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w) two_sum(tot, lo, ps[w * N + l]);
                store_rho(l, tot + lo);
            }
        } else {
            const unsigned tn_log2 = 31 - __clz(F.TN);
            for (int l = threadIdx.x; l < N; l += blockDim.x) {
                const unsigned tile = static_cast<unsigned>(l) >> tn_log2, lane = static_cast<unsigned>(l) & (F.TN - 1);
                const unsigned b_lo = tabulated ? s_blo[tile] : (tile * F.rpt) / F.rpc;
                const unsigned b_hi = tabulated ? s_bhi[tile] : ((tile + 1) * F.rpt - 1) / F.rpc;
                double tot = 0, lo = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w) two_sum(tot, lo, slot_sum(tile, lane, b_lo, b_hi, w));
                store_rho(l, tot + lo);
            }
        }
    }
    __syncthreads();
    TAIL_MARK(1);

    // ---- forward transform, dimension by dimension
    double2 *cur = A, *oth = B;
    transform_all(cur, oth, part, T.dim, Nx, Ny, Nz, twx, twy, twz, -1.0);

    TAIL_MARK(2);
    // ---- Poisson x collocation symbol, energy in Fourier space (full spectrum: every mode counted once)
    double e = 0;
    for (int idx = threadIdx.x; idx < N; idx += blockDim.x) {
        const int kx = idx % Nx;
        const int rest = idx / Nx;
        const int ky = rest % Ny;
        const int kz = rest / Ny;
        if (idx == 0) {
            cur[0] = make_double2(0.0, 0.0);
            continue;
        }
        const double2 Fk = cur[idx];
        const double kap2 = kap2x[kx] + kap2y[ky] + kap2z[kz];
        const double fac = T.fac_N / kap2;
        const double pr = Fk.x * fac, pi = Fk.y * fac;
        e += kap2 * (pr * pr + pi * pi);
        const double2 a = ilx[kx], b = ily[ky], c = ilz[kz];
        const double2 ab = make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
        const double2 s = make_double2(ab.x * c.x - ab.y * c.y, ab.x * c.y + ab.y * c.x);
        cur[idx] = make_double2(pr * s.x - pi * s.y, pr * s.y + pi * s.x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_down_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0 && S.energy_out) {
        double tot = 0;
        for (unsigned w = 0; w < (blockDim.x >> 5); ++w) tot += red[w];
        *S.energy_out = tot * T.vol_half;
    }

    // ---- inverse transform
    __syncthreads();
    TAIL_MARK(3);
    transform_all(cur, oth, part, T.dim, Nx, Ny, Nz, twx, twy, twz, 1.0);

    TAIL_MARK(4);
    // ---- level n: periodic coefficients (real part) -> device level format
    const ExpandParams &E = S.E;
    if (E.pp1d) {
        for (int i = threadIdx.x; i < Nx + 3; i += blockDim.x) S.raw1d[i] = cur[i % Nx].x;
        for (int k = threadIdx.x; k < Nx; k += blockDim.x) {
            double c[4], p0, p1, p2;
#pragma unroll
            for (int a = 0; a < 4; ++a) c[a] = cur[(k + a) % Nx].x;
            cell_poly_1d(c, E.g1, p0, p1, p2);
            S.level[2 * k] = p1;
            S.level[2 * k + 1] = p2;
            S.level[2 * Nx + k] = p0;
        }
        if (threadIdx.x == 0 && (3 * Nx) % 2) S.level[3 * Nx] = 0;
    } else if (E.xpp) {
        auto get = [&](size_t i) { return cur[i].x; };
        write_xpp<true>(get, S.level, S.raw1d, E, threadIdx.x, blockDim.x);
    } else {
        const int rows = E.dim >= 2 ? Ny + E.halo : 1;
        for (size_t idx = threadIdx.x; idx < E.level_doubles; idx += blockDim.x) {
            const int i = static_cast<int>(idx % E.sx);
            const size_t rest = idx / E.sx;
            const int j = static_cast<int>(rest % rows);
            const int k = static_cast<int>(rest / rows);
            double v = 0;
            if (i < Nx + E.halo && (E.dim < 3 ? k == 0 : k < Nz + E.halo)) v = cur[((k % Nz) * Ny + (j % Ny)) * Nx + (i % Nx)].x;
            S.level[idx] = v;
        }
    }
#ifdef NUFI_TAIL_TIMING
    TAIL_MARK(5);
    if (threadIdx.x == 0)
        printf("tail phases (cycles): rho+tw %lld [tables %lld, pdl_wait %lld, slot loads (peer step: polling loads of all ranks' sums) %lld, combine %lld]  fwd %lld  symbol %lld  inv %lld  expand %lld  total %lld\n",
               tmark[1] - tmark[0], tmark[8] - tmark[0], tmark[9] - tmark[8], tmark[10] - tmark[9], tmark[1] - tmark[10],
               tmark[2] - tmark[1], tmark[3] - tmark[2], tmark[4] - tmark[3], tmark[5] - tmark[4], tmark[5] - tmark[0]);
#endif
}

ExpandParams expand_params(const Handle *h)
{
    ExpandParams E{};
    E.dim = h->dim;
    E.Nx = static_cast<int>(h->c.Nx); E.Ny = static_cast<int>(h->c.Ny); E.Nz = static_cast<int>(h->c.Nz);
    E.sx = h->sx; E.sxy = h->sxy;
    E.xpp = h->xpp ? 1 : 0;
    E.halo = h->order - 1;
    E.pp1d = (h->dim == 1 && h->order == 4) ? 1 : 0;
    E.level_doubles = h->level_stride;
    E.g1 = -h->c.dt * h->c.dx_inv;
    return E;
}

unsigned blocks_for(size_t n, unsigned threads, unsigned cap)
{
    size_t b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    return static_cast<unsigned>(b > cap ? cap : b);
}

constexpr unsigned kSymbolBlocks = 128;

} // namespace

int tail_init(Handle *h)
{
    const nufi_b200_config3d &c = h->c;
    const int Nx = static_cast<int>(c.Nx), Ny = static_cast<int>(c.Ny), Nz = static_cast<int>(c.Nz);
    const int Nxh = Nx / 2 + 1;
    h->n_spec = static_cast<size_t>(Nxh) * Ny * Nz;
    cufftResult r;
    if (h->dim == 1) {
        r = cufftPlan1d(&h->plan_fwd, Nx, CUFFT_D2Z, 1);
        if (r == CUFFT_SUCCESS) r = cufftPlan1d(&h->plan_inv, Nx, CUFFT_Z2D, 1);
    } else if (h->dim == 2) {
        r = cufftPlan2d(&h->plan_fwd, Ny, Nx, CUFFT_D2Z);
        if (r == CUFFT_SUCCESS) r = cufftPlan2d(&h->plan_inv, Ny, Nx, CUFFT_Z2D);
    } else {
        r = cufftPlan3d(&h->plan_fwd, Nz, Ny, Nx, CUFFT_D2Z);
        if (r == CUFFT_SUCCESS) r = cufftPlan3d(&h->plan_inv, Nz, Ny, Nx, CUFFT_Z2D);
    }
    if (r != CUFFT_SUCCESS) return fail(h, NUFI_B200_ERR_CUDA, "cufftPlan failed with code " + std::to_string(static_cast<int>(r)));
    h->plans = true;

    // per-dimension tables, computed on the host with the reference's expressions
    const size_t nt = (static_cast<size_t>(Nx) + Ny + Nz + 1) & ~size_t(1); // keeps the double2 tables 16-byte aligned
    std::vector<double> tab(nt * 3);
    double *kap = tab.data();
    double *il = tab.data() + nt;
    const int Ns[3] = {Nx, Ny, Nz};
    const double Linv[3] = {c.Lx_inv, c.Ly_inv, c.Lz_inv};
    // B-spline basis at a node, N_i(0): the uniform-knot recurrence B_p[i] = ((t+p-i) B_{p-1}[i-1] + (1+i-t) B_{p-1}[i]) / p at t = 0
    double nodal[8] = {1, 0, 0, 0, 0, 0, 0, 0};
    for (int p = 1; p < h->order; ++p) {
        nodal[p] = 0.0;
        for (int i = p - 1; i >= 1; --i) nodal[i] = ((p - i) * nodal[i - 1] + (1 + i) * nodal[i]) / p;
        nodal[0] = nodal[0] / p;
    }
    size_t off = 0;
    for (int d = 0; d < 3; ++d) {
        const int N = Ns[d];
        const double fac = 2 * M_PI * Linv[d]; // poisson.cpp:71, 195-196, 333-335
        for (int k = 0; k < N; ++k) {
            double ii = (2 * k < N) ? k : N - k; // folded wavenumber (poisson.cpp:76, 207-208, 345-347)
            kap[off + k] = (d < h->dim) ? ii * ii * fac * fac : 0.0;
            // lambda(k) = sum_i N_i(0) w^i, w = exp(+2 pi i k/N): the collocation stencil of fields.hpp:76-79 (cubic:
            // N_i(0) = (1/6, 4/6, 1/6, 0), lambda = (1 + 4 w + w^2)/6)
            double lr = 1.0, li = 0.0;
            if (d < h->dim) {
                const double th = 2 * M_PI * static_cast<double>(k) / N;
                lr = 0.0;
                for (int i = 0; i < h->order; ++i) {
                    lr += nodal[i] * std::cos(i * th);
                    li += nodal[i] * std::sin(i * th);
                }
            }
            // odd orders on an even grid: lambda vanishes at the Nyquist mode (the stencil is symmetric about a half-integer) and
            // the collocation system is singular; the reference's LSMR (started from zero) returns the minimum-norm least-squares
            // solution, i.e. the pseudo-inverse: that mode of the coefficients is zero
            const double m2 = lr * lr + li * li;
            const bool null_mode = m2 < 1e-20;
            il[2 * (off + k)] = null_mode ? 0.0 : lr / m2;
            il[2 * (off + k) + 1] = null_mode ? 0.0 : -li / m2;
        }
        off += N;
    }
    NUFI_CUDA_CHECK(h, cudaMalloc(&h->d_symbol, tab.size() * sizeof(double)));
    NUFI_CUDA_CHECK(h, cudaMemcpy(h->d_symbol, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    {   // exact twiddles for the fused small-grid tail: (cos, sin)(2 pi m / N_d), the three tables back to back
        std::vector<double> tw(2 * (static_cast<size_t>(Nx) + Ny + Nz));
        size_t o = 0;
        for (int d = 0; d < 3; ++d)
            for (int m = 0; m < Ns[d]; ++m, ++o) {
                const double th = 2 * M_PI * static_cast<double>(m) / Ns[d];
                tw[2 * o] = std::cos(th);
                tw[2 * o + 1] = std::sin(th);
            }
        NUFI_CUDA_CHECK(h, cudaMalloc(&h->d_twiddle, tw.size() * sizeof(double)));
        NUFI_CUDA_CHECK(h, cudaMemcpy(h->d_twiddle, tw.data(), tw.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    NUFI_CUDA_CHECK(h, cudaMalloc(&h->d_spec, h->n_spec * sizeof(cufftDoubleComplex)));
    NUFI_CUDA_CHECK(h, cudaMalloc(&h->d_field, h->n_nodes * sizeof(double)));
    NUFI_CUDA_CHECK(h, cudaMalloc(&h->d_epart, (kSymbolBlocks + 1) * sizeof(double)));
    return NUFI_B200_OK;
}

void tail_destroy(Handle *h)
{
    if (h->plans) {
        cufftDestroy(h->plan_fwd);
        cufftDestroy(h->plan_inv);
        h->plans = false;
    }
    cudaFree(h->d_symbol);
    cudaFree(h->d_twiddle);
    h->d_twiddle = nullptr;
    cudaFree(h->d_spec);
    cudaFree(h->d_field);
    cudaFree(h->d_epart);
    h->d_symbol = nullptr; h->d_spec = nullptr; h->d_field = nullptr; h->d_epart = nullptr;
}

static bool small_tail_ok(const Handle *h)
{
    return h->n_nodes <= static_cast<size_t>(kSmallMaxNodes) && h->c.Nx <= kSmallMaxDim && h->c.Ny <= kSmallMaxDim &&
           h->c.Nz <= kSmallMaxDim;
}

bool tail_is_small(const Handle *h) { return small_tail_ok(h) && h->tail_force != 1; }

// rho (CPU convention, device) -> level n in the device history + energy[n].
// d_rho_full == nullptr: rho comes from the pending slot reduction of the last backtrace launch (fused step).
int tail_run(Handle *h, size_t n, const double *d_rho_full, bool from_peer)
{
    const nufi_b200_config3d &c = h->c;
    TailParams T{};
    T.dim = h->dim;
    T.Nx = static_cast<int>(c.Nx); T.Ny = static_cast<int>(c.Ny); T.Nz = static_cast<int>(c.Nz);
    T.Nxh = T.Nx / 2 + 1;
    const size_t nt = (c.Nx + c.Ny + c.Nz + 1) & ~size_t(1);
    T.kap2x = h->d_symbol; T.kap2y = T.kap2x + c.Nx; T.kap2z = T.kap2y + c.Ny;
    T.ilx = reinterpret_cast<const double2 *>(h->d_symbol + nt); T.ily = T.ilx + c.Nx; T.ilz = T.ily + c.Ny;
    T.fac_N = 1.0 / static_cast<double>(c.Nx * c.Ny * c.Nz);
    double vol_half = c.Lx;
    if (h->dim >= 2) vol_half = c.Lx * c.Ly;
    if (h->dim >= 3) vol_half = c.Lx * c.Ly * c.Lz;
    vol_half = vol_half / 2;
    T.vol_half = vol_half;
    T.mode = 3;
    ExpandParams E = expand_params(h);
    double *level = h->d_hist + n * h->level_stride;

    const bool small = h->tail_force == 2 || (h->tail_force == 0 && small_tail_ok(h));
    if (small) {
        if (!small_tail_ok(h)) return fail(h, NUFI_B200_ERR_ARG, "fused single-CTA tail forced but the grid is too large for it");
        SmallTailParams S{};
        S.T = T; S.E = E;
        S.twx = reinterpret_cast<const double2 *>(h->d_twiddle); S.twy = S.twx + c.Nx; S.twz = S.twy + c.Ny;
        S.level = level;
        S.raw1d = h->d_raw ? h->d_raw + n * h->raw_stride : nullptr;
        S.energy_out = h->d_energy + n;
        if (from_peer) { // multi-GPU step: this rank's slots (if it had a share) -> its sums -> every GPU; rho from all ranks' sums
            S.rho = nullptr;
            S.X = h->px.recv;
            if (h->fin_pending) {
                S.F = h->fin;
                S.XP = h->px.push;
                h->fin_pending = false;
            }
        } else if (d_rho_full) {
            S.rho = d_rho_full;
        } else if (h->fin_pending && h->fin.slots_ll) { // fused step: the tail adds the slots itself
            S.rho = nullptr;
            S.F = h->fin;
            h->fin_pending = false;
        } else { // the backtrace kernel's own epilogue, or finish_rho_kernel, leaves rho in memory
            int rc = launch_finish(h);
            if (rc) return rc;
            S.rho = h->d_rho_full;
        }
        const size_t smem = (2 * h->n_nodes + 2 * (c.Nx + c.Ny + c.Nz) + kSmallPart) * sizeof(double2) + ((c.Nx + c.Ny + c.Nz + 1) & ~size_t(1)) * sizeof(double);
        if (smem > 48 * 1024)
            NUFI_CUDA_CHECK(h, cudaFuncSetAttribute(tail_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        unsigned threads = kSmallThreads;
        if (const char *e = std::getenv("NUFI_B200_TAIL_THREADS")) { // experiments: every loop of the kernel strides by blockDim
            const int v = std::atoi(e);
            if (v >= 32 && v <= kSmallThreads && v % 32 == 0) threads = static_cast<unsigned>(v);
        }
        NUFI_CUDA_CHECK(h, launch_chained(h, tail_small_kernel, dim3(1), dim3(threads), smem, S));
        h->launches += 1;
        h->last_tail = "fused-1cta";
        h->level_valid[n] = 1;
        return NUFI_B200_OK;
    }

    if (from_peer) { // large grids: gather kernel (polls the ranks' self-validating sums, adds in rank order) -> d_rho_full
        int rc = launch_peer_gather(h);
        if (rc) return rc;
        d_rho_full = h->d_rho_full;
    }
    if (!d_rho_full) { // the cuFFT path reads rho from memory: run the slot reduction first
        int rc = launch_finish(h);
        if (rc) return rc;
        d_rho_full = h->d_rho_full;
    }
    if (cufftSetStream(h->plan_fwd, h->stream) != CUFFT_SUCCESS || cufftSetStream(h->plan_inv, h->stream) != CUFFT_SUCCESS)
        return fail(h, NUFI_B200_ERR_CUDA, "cufftSetStream failed");
    // cuFFT's D2Z may overwrite nothing of its input (out-of-place), Z2D may overwrite its input (d_spec: fine)
    if (cufftExecD2Z(h->plan_fwd, const_cast<double *>(d_rho_full), h->d_spec) != CUFFT_SUCCESS)
        return fail(h, NUFI_B200_ERR_CUDA, "cufftExecD2Z failed");
    const unsigned sblocks = blocks_for(h->n_spec, 256, kSymbolBlocks);
    symbol_kernel<<<sblocks, 256, 0, h->stream>>>(h->d_spec, T, h->d_epart);
    NUFI_CUDA_CHECK(h, cudaGetLastError());
    if (cufftExecZ2D(h->plan_inv, h->d_spec, h->d_field) != CUFFT_SUCCESS)
        return fail(h, NUFI_B200_ERR_CUDA, "cufftExecZ2D failed");
    if (E.pp1d) {
        expand1d_kernel<<<blocks_for(c.Nx + 3, 256, 64), 256, 0, h->stream>>>(h->d_field, h->d_raw + n * h->raw_stride, level, E,
                                                                          h->d_epart, sblocks, vol_half, h->d_energy + n);
    } else if (h->xpp) {
        expand_xpp_kernel<<<blocks_for(h->stride_t, 256, 1184), 256, 0, h->stream>>>(h->d_field, level, h->d_raw + n * h->raw_stride, E, 1,
                                                                                     h->d_epart, sblocks, vol_half, h->d_energy + n);
    } else {
        expand_kernel<<<blocks_for(h->level_stride, 256, 1184), 256, 0, h->stream>>>(h->d_field, level, E, 0.0, h->d_epart, sblocks,
                                                                                     vol_half, h->d_energy + n);
    }
    NUFI_CUDA_CHECK(h, cudaGetLastError());
    h->launches += 4; // D2Z, symbol, Z2D, expand (cuFFT may use more than one kernel per transform)
    h->last_tail = "cufft";
    h->level_valid[n] = 1;
    return NUFI_B200_OK;
}

// Reference-format level (halo, row stride Nx+3) from periodic coefficients; 1 thread per output element.
__global__ void expand_ref_kernel(const double *src, double *ref, int dim, int Nx, int Ny, int Nz, int halo)
{
    const int rx = Nx + halo, ry = dim >= 2 ? Ny + halo : 1, rz = dim >= 3 ? Nz + halo : 1;
    const size_t total = static_cast<size_t>(rx) * ry * rz;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int i = static_cast<int>(idx % rx);
        const size_t rest = idx / rx;
        const int j = static_cast<int>(rest % ry);
        const int k = static_cast<int>(rest / ry);
        ref[idx] = src[(static_cast<size_t>(k % Nz) * Ny + (j % Ny)) * Nx + (i % Nx)];
    }
}

__global__ void sum_epart_kernel(const double *epart, unsigned n, double vol_half, double *out)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double s = 0;
        for (unsigned b = 0; b < n; ++b) s += epart[b];
        *out = s * vol_half;
    }
}

// Spectral filter on a device vector of nodal values: mode 1 = poisson::solve (phi at the nodes + energy),
// mode 2 = interpolate (periodic coefficients).  Result in h->d_field; energy (mode 1) in h->d_epart[kSymbolBlocks].
int tail_filter(Handle *h, const double *d_values, int mode)
{
    const nufi_b200_config3d &c = h->c;
    TailParams T{};
    T.dim = h->dim;
    T.Nx = static_cast<int>(c.Nx); T.Ny = static_cast<int>(c.Ny); T.Nz = static_cast<int>(c.Nz);
    T.Nxh = T.Nx / 2 + 1;
    const size_t nt = (c.Nx + c.Ny + c.Nz + 1) & ~size_t(1);
    T.kap2x = h->d_symbol; T.kap2y = T.kap2x + c.Nx; T.kap2z = T.kap2y + c.Ny;
    T.ilx = reinterpret_cast<const double2 *>(h->d_symbol + nt); T.ily = T.ilx + c.Nx; T.ilz = T.ily + c.Ny;
    T.fac_N = 1.0 / static_cast<double>(c.Nx * c.Ny * c.Nz);
    double vol_half = c.Lx;
    if (h->dim >= 2) vol_half = c.Lx * c.Ly;
    if (h->dim >= 3) vol_half = c.Lx * c.Ly * c.Lz;
    T.vol_half = vol_half / 2;
    T.mode = mode;
    if (cufftSetStream(h->plan_fwd, h->stream) != CUFFT_SUCCESS || cufftSetStream(h->plan_inv, h->stream) != CUFFT_SUCCESS)
        return fail(h, NUFI_B200_ERR_CUDA, "cufftSetStream failed");
    if (cufftExecD2Z(h->plan_fwd, const_cast<double *>(d_values), h->d_spec) != CUFFT_SUCCESS)
        return fail(h, NUFI_B200_ERR_CUDA, "cufftExecD2Z failed");
    const unsigned sblocks = blocks_for(h->n_spec, 256, kSymbolBlocks);
    symbol_kernel<<<sblocks, 256, 0, h->stream>>>(h->d_spec, T, h->d_epart);
    NUFI_CUDA_CHECK(h, cudaGetLastError());
    if (cufftExecZ2D(h->plan_inv, h->d_spec, h->d_field) != CUFFT_SUCCESS)
        return fail(h, NUFI_B200_ERR_CUDA, "cufftExecZ2D failed");
    sum_epart_kernel<<<1, 32, 0, h->stream>>>(h->d_epart, sblocks, T.vol_half, h->d_epart + kSymbolBlocks);
    NUFI_CUDA_CHECK(h, cudaGetLastError());
    h->launches += 4;
    return NUFI_B200_OK;
}

// periodic coefficients in h->d_field -> reference-format level in h->d_stage
int expand_field_to_stage(Handle *h)
{
    expand_ref_kernel<<<blocks_for(h->stride_t, 256, 1184), 256, 0, h->stream>>>(h->d_field, h->d_stage, h->dim, static_cast<int>(h->c.Nx),
                                                                               static_cast<int>(h->c.Ny), static_cast<int>(h->c.Nz), h->order - 1);
    NUFI_CUDA_CHECK(h, cudaGetLastError());
    h->launches += 1;
    return NUFI_B200_OK;
}

double *tail_energy_scratch(Handle *h) { return h->d_epart + kSymbolBlocks; }

int convert_level_to_device(Handle *h, size_t n, const double *d_ref_level)
{
    ExpandParams E = expand_params(h);
    if (h->xpp)
        expand_xpp_kernel<<<blocks_for(h->stride_t, 256, 1184), 256, 0, h->stream>>>(d_ref_level, h->d_hist + n * h->level_stride,
                                                                                     h->d_raw + n * h->raw_stride, E, 0, nullptr, 0, 0.0, nullptr);
    else
    ref_to_device_kernel<<<blocks_for(h->level_stride, 256, 1184), 256, 0, h->stream>>>(
        d_ref_level, h->d_hist + n * h->level_stride, E.pp1d ? h->d_raw + n * h->raw_stride : nullptr, E);
    NUFI_CUDA_CHECK(h, cudaGetLastError());
    h->launches += 1;
    return NUFI_B200_OK;
}

int convert_level_from_device(Handle *h, size_t n, double *d_ref_level)
{
    ExpandParams E = expand_params(h);
    device_to_ref_kernel<<<blocks_for(h->stride_t, 256, 1184), 256, 0, h->stream>>>(
        h->d_hist + n * h->level_stride, h->d_raw ? h->d_raw + n * h->raw_stride : nullptr, d_ref_level, E);
    NUFI_CUDA_CHECK(h, cudaGetLastError());
    h->launches += 1;
    return NUFI_B200_OK;
}

int make_full_rho(Handle *h, const double *d_partial_sum, double *d_full)
{
    full_rho_kernel<<<blocks_for(h->n_nodes, 256, 1184), 256, 0, h->stream>>>(d_partial_sum, d_full, h->n_nodes);
    NUFI_CUDA_CHECK(h, cudaGetLastError());
    h->launches += 1;
    return NUFI_B200_OK;
}

} // namespace nufi_b200
