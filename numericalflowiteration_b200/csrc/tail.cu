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

namespace nufi_b200
{

namespace
{

struct TailParams
{
    int Nx, Ny, Nz, Nxh;            // Nxh = Nx/2 + 1
    const double *kap2x, *kap2y, *kap2z; // per-dimension ii*ii*fac*fac
    const double2 *ilx, *ily, *ilz;      // per-dimension 1/lambda
    double fac_N;                   // 1/(Nx Ny Nz)
    double vol_half;                // Lx Ly Lz / 2
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
        if (kx == 0 && ky == 0 && kz == 0) {
            spec[idx] = make_cuDoubleComplex(0.0, 0.0); // data[0] = 0 (poisson.cpp:83, 214, 357)
            continue;
        }
        const double kap2 = T.kap2x[kx] + T.kap2y[ky] + T.kap2z[kz];
        const double fac = T.fac_N / kap2;
        const double pr = F.x * fac, pi = F.y * fac; // phi^_k / N-normalised
        const double w = (kx == 0 || 2 * kx == T.Nx) ? 1.0 : 2.0;
        e += w * kap2 * (pr * pr + pi * pi);
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

struct ExpandParams
{
    int dim, Nx, Ny, Nz, sx, sxy, Nxp;
    size_t level_doubles; // device level size
    double g1;            // 1d: -dt*dx_inv
    int shift;            // 0: src indexed by periodic coefficient index
};

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
        const int rows = E.Ny + 3;
        const int j = static_cast<int>(rest % rows);
        const int k = static_cast<int>(rest / rows);
        double v = 0;
        if (i < E.Nx + 3 && (E.dim < 3 ? k == 0 : k < E.Nz + 3))
            v = src[(static_cast<size_t>(k % E.Nz) * E.Ny + (j % E.Ny)) * E.Nx + (i % E.Nx)];
        level[idx] = v;
    }
    if (energy_out && blockIdx.x == 0 && threadIdx.x == 0) {
        double s = 0;
        for (unsigned b = 0; b < n_epart; ++b) s += epart[b];
        *energy_out = s * vol_half; // energy *= Lx*Ly*Lz/2 (poisson.cpp:87, 217, 360)
    }
}

// 1d: raw level with halo + per-cell quadratics [p0 | p1 | p2], each Nxp long.
__global__ void expand1d_kernel(const double *src, double *raw, double *pp, ExpandParams E, const double *epart,
                                unsigned n_epart, double vol_half, double *energy_out)
{
    const int Nx = E.Nx;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Nx + 3; i += gridDim.x * blockDim.x) raw[i] = src[i % Nx];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < E.Nxp; k += gridDim.x * blockDim.x) {
        double p0 = 0, p1 = 0, p2 = 0;
        if (k < Nx) {
            double c[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) c[a] = src[(k + a) % Nx];
            cell_poly_1d(c, E.g1, p0, p1, p2);
        }
        pp[k] = p0;
        pp[E.Nxp + k] = p1;
        pp[2 * E.Nxp + k] = p2;
    }
    if (energy_out && blockIdx.x == 0 && threadIdx.x == 0) {
        double s = 0;
        for (unsigned b = 0; b < n_epart; ++b) s += epart[b];
        *energy_out = s * vol_half;
    }
}

// reference-format level (with halo, row stride Nx+3) -> device format
__global__ void ref_to_device_kernel(const double *ref, double *level, double *raw1d, ExpandParams E)
{
    if (E.dim == 1) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < E.Nx + 3; i += gridDim.x * blockDim.x) raw1d[i] = ref[i];
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < E.Nxp; k += gridDim.x * blockDim.x) {
            double p0 = 0, p1 = 0, p2 = 0;
            if (k < E.Nx) cell_poly_1d(ref + k, E.g1, p0, p1, p2);
            level[k] = p0;
            level[E.Nxp + k] = p1;
            level[2 * E.Nxp + k] = p2;
        }
        return;
    }
    const size_t total = E.level_doubles;
    const int rx = E.Nx + 3, ry = E.Ny + 3;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int i = static_cast<int>(idx % E.sx);
        const size_t rest = idx / E.sx;
        const int j = static_cast<int>(rest % ry);
        const int k = static_cast<int>(rest / ry);
        double v = 0;
        if (i < rx && (E.dim < 3 ? k == 0 : k < E.Nz + 3)) v = ref[(static_cast<size_t>(k) * ry + j) * rx + i];
        level[idx] = v;
    }
}

__global__ void device_to_ref_kernel(const double *level, const double *raw1d, double *ref, ExpandParams E)
{
    if (E.dim == 1) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < E.Nx + 3; i += gridDim.x * blockDim.x) ref[i] = raw1d[i];
        return;
    }
    const int rx = E.Nx + 3, ry = E.Ny + 3, rz = E.dim == 3 ? E.Nz + 3 : 1;
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

ExpandParams expand_params(const Handle *h)
{
    ExpandParams E{};
    E.dim = h->dim;
    E.Nx = static_cast<int>(h->c.Nx); E.Ny = static_cast<int>(h->c.Ny); E.Nz = static_cast<int>(h->c.Nz);
    E.sx = h->sx; E.sxy = h->sxy; E.Nxp = h->Nxp;
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
    size_t off = 0;
    for (int d = 0; d < 3; ++d) {
        const int N = Ns[d];
        const double fac = 2 * M_PI * Linv[d]; // poisson.cpp:71, 195-196, 333-335
        for (int k = 0; k < N; ++k) {
            double ii = (2 * k < N) ? k : N - k; // folded wavenumber (poisson.cpp:76, 207-208, 345-347)
            kap[off + k] = (d < h->dim) ? ii * ii * fac * fac : 0.0;
            // lambda(k) = (1 + 4 w + w^2)/6, w = exp(+2 pi i k/N): N_i(0) = (1/6, 4/6, 1/6, 0) (fields.hpp:76-79)
            double lr = 1.0, li = 0.0;
            if (d < h->dim) {
                const double th = 2 * M_PI * static_cast<double>(k) / N;
                lr = (1.0 + 4.0 * std::cos(th) + std::cos(2 * th)) / 6.0;
                li = (4.0 * std::sin(th) + std::sin(2 * th)) / 6.0;
            }
            const double m2 = lr * lr + li * li;
            il[2 * (off + k)] = lr / m2;
            il[2 * (off + k) + 1] = -li / m2;
        }
        off += N;
    }
    NUFI_CUDA_CHECK(h, cudaMalloc(&h->d_symbol, tab.size() * sizeof(double)));
    NUFI_CUDA_CHECK(h, cudaMemcpy(h->d_symbol, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    NUFI_CUDA_CHECK(h, cudaMalloc(&h->d_spec, h->n_spec * sizeof(cufftDoubleComplex)));
    NUFI_CUDA_CHECK(h, cudaMalloc(&h->d_field, h->n_nodes * sizeof(double)));
    NUFI_CUDA_CHECK(h, cudaMalloc(&h->d_epart, kSymbolBlocks * sizeof(double)));
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
    cudaFree(h->d_spec);
    cudaFree(h->d_field);
    cudaFree(h->d_epart);
    h->d_symbol = nullptr; h->d_spec = nullptr; h->d_field = nullptr; h->d_epart = nullptr;
}

// rho (CPU convention, device) -> level n in the device history + energy[n]
int tail_run(Handle *h, size_t n, const double *d_rho_full)
{
    const nufi_b200_config3d &c = h->c;
    if (cufftSetStream(h->plan_fwd, h->stream) != CUFFT_SUCCESS || cufftSetStream(h->plan_inv, h->stream) != CUFFT_SUCCESS)
        return fail(h, NUFI_B200_ERR_CUDA, "cufftSetStream failed");
    // cuFFT's D2Z may overwrite nothing of its input (out-of-place), Z2D may overwrite its input (d_spec: fine)
    if (cufftExecD2Z(h->plan_fwd, const_cast<double *>(d_rho_full), h->d_spec) != CUFFT_SUCCESS)
        return fail(h, NUFI_B200_ERR_CUDA, "cufftExecD2Z failed");
    TailParams T{};
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
    const unsigned sblocks = blocks_for(h->n_spec, 256, kSymbolBlocks);
    symbol_kernel<<<sblocks, 256, 0, h->stream>>>(h->d_spec, T, h->d_epart);
    NUFI_CUDA_CHECK(h, cudaGetLastError());
    if (cufftExecZ2D(h->plan_inv, h->d_spec, h->d_field) != CUFFT_SUCCESS)
        return fail(h, NUFI_B200_ERR_CUDA, "cufftExecZ2D failed");
    ExpandParams E = expand_params(h);
    double *level = h->d_hist + n * h->level_stride;
    if (h->dim == 1) {
        expand1d_kernel<<<blocks_for(h->Nxp, 256, 64), 256, 0, h->stream>>>(h->d_field, h->d_raw + n * h->raw_stride, level, E,
                                                                          h->d_epart, sblocks, vol_half, h->d_energy + n);
    } else {
        expand_kernel<<<blocks_for(h->level_stride, 256, 1184), 256, 0, h->stream>>>(h->d_field, level, E, 0.0, h->d_epart, sblocks,
                                                                                     vol_half, h->d_energy + n);
    }
    NUFI_CUDA_CHECK(h, cudaGetLastError());
    h->launches += 4; // D2Z, symbol, Z2D, expand (cuFFT may use more than one kernel per transform)
    h->level_valid[n] = 1;
    return NUFI_B200_OK;
}

int convert_level_to_device(Handle *h, size_t n, const double *d_ref_level)
{
    ExpandParams E = expand_params(h);
    ref_to_device_kernel<<<blocks_for(h->level_stride, 256, 1184), 256, 0, h->stream>>>(
        d_ref_level, h->d_hist + n * h->level_stride, h->dim == 1 ? h->d_raw + n * h->raw_stride : nullptr, E);
    NUFI_CUDA_CHECK(h, cudaGetLastError());
    h->launches += 1;
    return NUFI_B200_OK;
}

int convert_level_from_device(Handle *h, size_t n, double *d_ref_level)
{
    ExpandParams E = expand_params(h);
    device_to_ref_kernel<<<blocks_for(h->stride_t, 256, 1184), 256, 0, h->stream>>>(
        h->d_hist + n * h->level_stride, h->dim == 1 ? h->d_raw + n * h->raw_stride : nullptr, d_ref_level, E);
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
