// api.cu -- the C ABI of libnufi_b200.so (include/nufi_b200.h).  Owns all device state of one GPU:
// the (Nt+1)-level coefficient history in the device level format, rho, metrics, energies, FFT plans.
// Mirrors the surface of nufi::dim{1,2,3}::cuda_kernel (nufi/cuda_kernel.cu:81-189, 273-371, 468-573).
#include "internal.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

namespace nufi_b200
{

static thread_local std::string g_create_error;

int fail(Handle *h, int code, const std::string &msg)
{
    if (h) h->err = msg;
    else g_create_error = msg;
    return code;
}

// ---- event ring: one (start, stop) pair per backtrace launch, read back lazily so timing never stalls the stream
static int ev_fold_oldest(Handle *h)
{
    const size_t oldest = (h->ev_head + kEvRingPairs - h->ev_pending) % kEvRingPairs;
    NUFI_CUDA_CHECK(h, cudaEventSynchronize(h->ev_ring[2 * oldest + 1]));
    float ms = 0;
    NUFI_CUDA_CHECK(h, cudaEventElapsedTime(&ms, h->ev_ring[2 * oldest], h->ev_ring[2 * oldest + 1]));
    h->bt_ms_total += ms;
    h->bt_ms_last = ms;
    h->bt_count += 1;
    h->ev_pending -= 1;
    return NUFI_B200_OK;
}

int ev_acquire(Handle *h, cudaEvent_t *start, cudaEvent_t *stop)
{
    if (h->ev_ring.empty()) h->ev_ring.assign(2 * kEvRingPairs, nullptr);
    if (h->ev_pending == kEvRingPairs) { // full: fold the oldest pair first
        int rc = ev_fold_oldest(h);
        if (rc) return rc;
    }
    for (int i = 0; i < 2; ++i)
        if (!h->ev_ring[2 * h->ev_head + i]) NUFI_CUDA_CHECK(h, cudaEventCreate(&h->ev_ring[2 * h->ev_head + i]));
    *start = h->ev_ring[2 * h->ev_head];
    *stop = h->ev_ring[2 * h->ev_head + 1];
    h->ev_head = (h->ev_head + 1) % kEvRingPairs;
    return NUFI_B200_OK;
}

int ev_drain(Handle *h)
{
    while (h->ev_pending > 0) {
        int rc = ev_fold_oldest(h);
        if (rc) return rc;
    }
    return NUFI_B200_OK;
}

namespace
{

void free_all(Handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    tail_destroy(h);
    peer_free(h);
    cudaFree(h->d_hist); cudaFree(h->d_raw);
    cudaFree(h->d_rho_partial); cudaFree(h->d_rho_full); cudaFree(h->d_partials); cudaFree(h->d_partials_ll); cudaFree(h->d_ll_status);
    cudaFree(h->d_metrics); cudaFree(h->d_mpartials); cudaFree(h->d_energy); cudaFree(h->d_stage); cudaFree(h->d_done);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    if (h->h_up) cudaFreeHost(h->h_up);
    if (h->ev_up) cudaEventDestroy(h->ev_up);
    for (cudaEvent_t e : h->ev_ring)
        if (e) cudaEventDestroy(e);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

int create_common(const nufi_b200_config3d &c, int dim, int order, const nufi_b200_f0 *f0, int device, nufi_b200_handle **out)
{
    if (!out) return fail(nullptr, NUFI_B200_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (order < 3 || order > 8) // the orders the reference instantiates (nufi/cuda_kernel.cu:191-203, 373-385, 575-587)
        return fail(nullptr, NUFI_B200_ERR_ARG, "spline order must be 3..8 (4 = cubic, what every reference driver uses, is the specialised fast path)");
    const size_t o = static_cast<size_t>(order), halo = o - 1;
    const size_t nmin = o < 4 ? 4 : o;
    if (c.Nx < nmin || (dim >= 2 && c.Ny < nmin) || (dim >= 3 && c.Nz < nmin) || c.Nu == 0 || (dim >= 2 && c.Nv == 0) || (dim >= 3 && c.Nw == 0))
        return fail(nullptr, NUFI_B200_ERR_ARG, "grid too small: need N >= max(4, order) nodes per spatial dimension and >= 1 velocity node");
    if (c.Nx > (1u << 20) || c.Ny > (1u << 20) || c.Nz > (1u << 20))
        return fail(nullptr, NUFI_B200_ERR_ARG, "grid too large");
    if (!f0 || f0->kind < 0 || f0->kind > (dim == 3 ? 2 : 1)) return fail(nullptr, NUFI_B200_ERR_ARG, "unknown f0 kind");
    {   // f0 is evaluated at the periodic image of the foot inside the box (the reference evaluates it at the unwrapped position):
        // the two agree iff f0 has the period of the box, i.e. k*L is a multiple of 2 pi in every dimension.  Refuse anything else.
        const double k = f0->p[1];
        const double Ls[3] = {c.Lx, c.Ly, c.Lz};
        for (int d = 0; d < dim; ++d) {
            const double periods = k * Ls[d] / (2.0 * 3.14159265358979323846);
            if (!(std::fabs(periods - std::nearbyint(periods)) <= 1e-9 * (1.0 + std::fabs(periods))))
                return fail(nullptr, NUFI_B200_ERR_ARG,
                            "f0 is not periodic in the box: its wavenumber k = p[1] must satisfy k*L = 2*pi*m in every dimension "
                            "(the device evaluates f0 at the periodic image of the foot of the characteristic)");
        }
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, NUFI_B200_ERR_CUDA,
                    std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                        " (libnufi_b200 has no CPU fallback)");
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= ndev) return fail(nullptr, NUFI_B200_ERR_ARG, "device index out of range");

    Handle *h = new (std::nothrow) Handle;
    if (!h) return fail(nullptr, NUFI_B200_ERR_ALLOC, "out of host memory");
    h->dim = dim; h->order = order; h->device = device; h->c = c; h->f0 = *f0; h->Nt = c.Nt;
    h->n_nodes = c.Nx * c.Ny * c.Nz;
    h->n_vel = c.Nu * c.Nv * c.Nw;
    h->stride_t = (c.Nx + halo) * (dim >= 2 ? c.Ny + halo : 1) * (dim >= 3 ? c.Nz + halo : 1);
    // device level format
    if (order != 4) { // generic orders: the reference layout itself (halo of order-1), padded to a 16-byte multiple
        h->sx = static_cast<int>(c.Nx + halo);
        h->sxy = dim >= 2 ? h->sx * static_cast<int>(c.Ny + halo) : 0;
        h->level_stride = (h->stride_t + 1) & ~size_t(1);
    } else if (dim == 1) {
        h->level_stride = (3 * c.Nx + 1) & ~size_t(1); // per-cell quadratics [Nx x (p1, p2)] [Nx x p0] (tail.cu), 16-byte multiple
        h->raw_stride = (c.Nx + 3 + 1) & ~size_t(1);
        h->sx = 0; h->sxy = 0;
    } else {
        h->sx = static_cast<int>(c.Nx + 3);
        h->sxy = h->sx * static_cast<int>(c.Ny + 3);
        size_t ls = static_cast<size_t>(h->sxy) * (dim == 3 ? c.Nz + 3 : 1);
        h->level_stride = (ls + 1) & ~size_t(1); // multiple of 16 bytes for the bulk copies
        // x-direction pp-form ("xpp"): 4 doubles per (row, cell), cuts the FP64 work of a point-step by a quarter to a
        // third at 4 Nx/(Nx+3) times the bytes.  Used when two such levels still fit the shared-memory ring.
        const size_t rows = (c.Ny + 3) * (dim == 3 ? c.Nz + 3 : 1);
        const size_t xpp_stride = rows * 4 * c.Nx;
        int want = 2 * xpp_stride * sizeof(double) + 16 * 1024 <= 227 * 1024 ? 1 : 0;
        if (const char *e = std::getenv("NUFI_B200_XPP")) want = std::atoi(e);
        if (want) {
            h->xpp = true;
            h->level_stride = xpp_stride;
            h->raw_stride = (h->stride_t + 1) & ~size_t(1);
            h->sx = static_cast<int>(2 * c.Nx);            // row stride in double2 units; the (a2,a3) half starts Nx further
            h->sxy = h->sx * static_cast<int>(c.Ny + 3);
        }
    }
    h->level_valid.assign(c.Nt + 1, 0);

#define CREATE_CHECK(expr)                                                                                      \
    do {                                                                                                        \
        cudaError_t e__ = (expr);                                                                               \
        if (e__ != cudaSuccess) {                                                                               \
            std::string m = std::string(cudaGetErrorName(e__)) + ": " + cudaGetErrorString(e__) + " [" #expr "]"; \
            int code = (e__ == cudaErrorMemoryAllocation) ? NUFI_B200_ERR_ALLOC : NUFI_B200_ERR_CUDA;           \
            free_all(h);                                                                                        \
            cudaGetLastError();                                                                                 \
            return fail(nullptr, code, m);                                                                      \
        }                                                                                                       \
    } while (0)

    CREATE_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CREATE_CHECK(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    if (const char *e = std::getenv("NUFI_B200_PDL")) h->pdl = std::atoi(e) != 0;
    CREATE_CHECK(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    const size_t hist_bytes = (c.Nt + 1) * h->level_stride * sizeof(double);
    CREATE_CHECK(cudaMalloc(&h->d_hist, hist_bytes));
    CREATE_CHECK(cudaMemsetAsync(h->d_hist, 0, hist_bytes, h->stream));
    if ((dim == 1 && order == 4) || h->xpp) {
        CREATE_CHECK(cudaMalloc(&h->d_raw, (c.Nt + 1) * h->raw_stride * sizeof(double)));
        CREATE_CHECK(cudaMemsetAsync(h->d_raw, 0, (c.Nt + 1) * h->raw_stride * sizeof(double), h->stream));
    }
    CREATE_CHECK(cudaMalloc(&h->d_rho_partial, h->n_nodes * sizeof(double)));
    CREATE_CHECK(cudaMalloc(&h->d_rho_full, h->n_nodes * sizeof(double)));
    CREATE_CHECK(cudaMemsetAsync(h->d_rho_partial, 0, h->n_nodes * sizeof(double), h->stream));
    CREATE_CHECK(cudaMemsetAsync(h->d_rho_full, 0, h->n_nodes * sizeof(double), h->stream));
    CREATE_CHECK(cudaMalloc(&h->d_metrics, 4 * sizeof(double)));
    CREATE_CHECK(cudaMemsetAsync(h->d_metrics, 0, 4 * sizeof(double), h->stream));
    CREATE_CHECK(cudaMalloc(&h->d_mpartials, 4 * sizeof(double) * h->sm_count));
    CREATE_CHECK(cudaMalloc(&h->d_energy, (c.Nt + 1) * sizeof(double)));
    CREATE_CHECK(cudaMemsetAsync(h->d_energy, 0, (c.Nt + 1) * sizeof(double), h->stream));
    CREATE_CHECK(cudaMalloc(&h->d_stage, h->stride_t * sizeof(double)));
    CREATE_CHECK(cudaMalloc(&h->d_done, sizeof(unsigned int)));
    CREATE_CHECK(cudaMemsetAsync(h->d_done, 0, sizeof(unsigned int), h->stream));
    CREATE_CHECK(cudaMalloc(&h->d_ll_status, sizeof(int)));
    CREATE_CHECK(cudaMemsetAsync(h->d_ll_status, 0, sizeof(int), h->stream));
    h->h_pinned_cap = h->stride_t + h->n_nodes + 2; // step_host returns a level, rho and the energy with one copy-back
    if (h->h_pinned_cap < c.Nt + 1) h->h_pinned_cap = c.Nt + 1;
    CREATE_CHECK(cudaMallocHost(&h->h_pinned, h->h_pinned_cap * sizeof(double)));
    CREATE_CHECK(cudaMallocHost(&h->h_up, h->h_pinned_cap * sizeof(double)));
    CREATE_CHECK(cudaEventCreateWithFlags(&h->ev_up, cudaEventDisableTiming));
#undef CREATE_CHECK
    int rc = tail_init(h);
    if (rc != NUFI_B200_OK) {
        std::string m = h->err;
        free_all(h);
        return fail(nullptr, rc, m);
    }
    cudaStreamSynchronize(h->stream);
    *out = reinterpret_cast<nufi_b200_handle *>(h);
    return NUFI_B200_OK;
}

inline Handle *H(nufi_b200_handle *h) { return reinterpret_cast<Handle *>(h); }
inline const Handle *H(const nufi_b200_handle *h) { return reinterpret_cast<const Handle *>(h); }

// host -> device through the upload staging buffer: the caller's buffer is consumed on return, the stream is not synchronised
int stage_upload(Handle *h, double *d_dst, const double *host_src, size_t count)
{
    if (h->ev_up_pending) {
        NUFI_CUDA_CHECK(h, cudaEventSynchronize(h->ev_up)); // the previous copy out of h_up (normally long finished)
        h->ev_up_pending = false;
    }
    std::memcpy(h->h_up, host_src, sizeof(double) * count);
    NUFI_CUDA_CHECK(h, cudaMemcpyAsync(d_dst, h->h_up, sizeof(double) * count, cudaMemcpyHostToDevice, h->stream));
    NUFI_CUDA_CHECK(h, cudaEventRecord(h->ev_up, h->stream));
    h->ev_up_pending = true;
    return NUFI_B200_OK;
}

int check_levels(Handle *h, size_t first_needed_exclusive_end, const char *what)
{
    for (size_t m = 0; m < first_needed_exclusive_end; ++m)
        if (!h->level_valid[m])
            return fail(h, NUFI_B200_ERR_RANGE, std::string(what) + ": history level " + std::to_string(m) +
                                                    " was never uploaded or computed");
    return NUFI_B200_OK;
}

} // namespace
} // namespace nufi_b200

using namespace nufi_b200;

#define ENTER(h)                                                         \
    Handle *hh = H(h);                                                   \
    if (!hh) return fail(nullptr, NUFI_B200_ERR_ARG, "handle is NULL");  \
    NUFI_CUDA_CHECK(hh, cudaSetDevice(hh->device))

extern "C" {

int nufi_b200_create_1d(const nufi_b200_config1d *conf, int order, const nufi_b200_f0 *f0, int device, nufi_b200_handle **out)
{
    if (!conf) return fail(nullptr, NUFI_B200_ERR_ARG, "conf is NULL");
    nufi_b200_config3d c{};
    c.Nx = conf->Nx; c.Ny = 1; c.Nz = 1; c.Nu = conf->Nu; c.Nv = 1; c.Nw = 1; c.Nt = conf->Nt; c.dt = conf->dt;
    c.x_min = conf->x_min; c.x_max = conf->x_max; c.u_min = conf->u_min; c.u_max = conf->u_max;
    c.dx = conf->dx; c.dx_inv = conf->dx_inv; c.Lx = conf->Lx; c.Lx_inv = conf->Lx_inv; c.du = conf->du;
    c.dy = c.dz = c.dy_inv = c.dz_inv = c.Ly = c.Lz = c.Ly_inv = c.Lz_inv = 1; c.dv = c.dw = 1;
    c.v_max = c.w_max = 1;
    return create_common(c, 1, order, f0, device, out);
}

int nufi_b200_create_2d(const nufi_b200_config2d *conf, int order, const nufi_b200_f0 *f0, int device, nufi_b200_handle **out)
{
    if (!conf) return fail(nullptr, NUFI_B200_ERR_ARG, "conf is NULL");
    nufi_b200_config3d c{};
    c.Nx = conf->Nx; c.Ny = conf->Ny; c.Nz = 1; c.Nu = conf->Nu; c.Nv = conf->Nv; c.Nw = 1; c.Nt = conf->Nt; c.dt = conf->dt;
    c.x_min = conf->x_min; c.x_max = conf->x_max; c.y_min = conf->y_min; c.y_max = conf->y_max;
    c.u_min = conf->u_min; c.u_max = conf->u_max; c.v_min = conf->v_min; c.v_max = conf->v_max;
    c.dx = conf->dx; c.dx_inv = conf->dx_inv; c.Lx = conf->Lx; c.Lx_inv = conf->Lx_inv;
    c.dy = conf->dy; c.dy_inv = conf->dy_inv; c.Ly = conf->Ly; c.Ly_inv = conf->Ly_inv;
    c.du = conf->du; c.dv = conf->dv;
    c.dz = c.dz_inv = c.Lz = c.Lz_inv = 1; c.dw = 1; c.w_max = 1;
    return create_common(c, 2, order, f0, device, out);
}

int nufi_b200_create_3d(const nufi_b200_config3d *conf, int order, const nufi_b200_f0 *f0, int device, nufi_b200_handle **out)
{
    if (!conf) return fail(nullptr, NUFI_B200_ERR_ARG, "conf is NULL");
    return create_common(*conf, 3, order, f0, device, out);
}

void nufi_b200_destroy(nufi_b200_handle *h) { free_all(H(h)); }

const char *nufi_b200_last_error(const nufi_b200_handle *h) { return h ? H(h)->err.c_str() : g_create_error.c_str(); }

int nufi_b200_compute_rho(nufi_b200_handle *h, size_t n, size_t q_begin, size_t q_end)
{
    ENTER(h);
    if (n > hh->Nt) return fail(hh, NUFI_B200_ERR_RANGE, "Time-step out of range."); // cuda_kernel.cu:115-116
    const size_t nq = hh->n_nodes * hh->n_vel;
    if (q_begin > q_end || q_end > nq) return fail(hh, NUFI_B200_ERR_RANGE, "quadrature range out of bounds");
    if (q_begin == q_end) { // reference: no-op after the scheduler's early return; the partial is defined as zero
        NUFI_CUDA_CHECK(hh, cudaMemsetAsync(hh->d_rho_partial, 0, sizeof(double) * hh->n_nodes, hh->stream));
        return NUFI_B200_OK;
    }
    int rc = check_levels(hh, n, "compute_rho");
    if (rc) return rc;
    return launch_backtrace(hh, n, q_begin, q_end, false);
}

int nufi_b200_download_rho(nufi_b200_handle *h, double *rho_host)
{
    ENTER(h);
    if (!rho_host) return fail(hh, NUFI_B200_ERR_ARG, "rho_host is NULL");
    NUFI_CUDA_CHECK(hh, cudaMemcpyAsync(hh->h_pinned, hh->d_rho_partial, sizeof(double) * hh->n_nodes, cudaMemcpyDeviceToHost, hh->stream));
    NUFI_CUDA_CHECK(hh, cudaStreamSynchronize(hh->stream));
    for (size_t i = 0; i < hh->n_nodes; ++i) rho_host[i] += hh->h_pinned[i]; // cuda_kernel.cu:143-144
    return NUFI_B200_OK;
}

int nufi_b200_upload_phi(nufi_b200_handle *h, size_t n, const double *coeffs_base)
{
    ENTER(h);
    if (n > hh->Nt) return fail(hh, NUFI_B200_ERR_RANGE, "Time-step out of range.");
    if (!coeffs_base) return fail(hh, NUFI_B200_ERR_ARG, "coeffs is NULL");
    // slice n (cuda_kernel.cu:154-155).  The reference's cudaMemcpy blocks; here the slice is staged in pinned memory, so the
    // caller may reuse its array on return while copy and layout conversion run asynchronously on the stream.
    int rc = stage_upload(hh, hh->d_stage, coeffs_base + n * hh->stride_t, hh->stride_t);
    if (rc) return rc;
    rc = convert_level_to_device(hh, n, hh->d_stage);
    if (rc) return rc;
    hh->level_valid[n] = 1;
    return NUFI_B200_OK;
}

int nufi_b200_compute_metrics(nufi_b200_handle *h, size_t n, size_t q_begin, size_t q_end)
{
    ENTER(h);
    if (n > hh->Nt) return fail(hh, NUFI_B200_ERR_RANGE, "Time-step out of range.");
    const size_t nq = hh->mgrid_set ? hh->mconf.Nx * hh->mconf.Nu : hh->n_nodes * hh->n_vel;
    if (q_begin > q_end || q_end > nq) return fail(hh, NUFI_B200_ERR_RANGE, "quadrature range out of bounds");
    if (q_begin == q_end) {
        NUFI_CUDA_CHECK(hh, cudaMemsetAsync(hh->d_metrics, 0, sizeof(double) * 4, hh->stream));
        return NUFI_B200_OK;
    }
    int rc = check_levels(hh, n == 0 ? 0 : n + 1, "compute_metrics");
    if (rc) return rc;
    return launch_backtrace(hh, n, q_begin, q_end, true);
}

int nufi_b200_set_metrics_grid_1d(nufi_b200_handle *h, const nufi_b200_config1d *conf_metrics)
{
    ENTER(h);
    if (hh->dim != 1) return fail(hh, NUFI_B200_ERR_ARG, "a separate metrics grid exists for dim 1 only (nufi/cuda_scheduler.hpp:65-85)");
    if (!conf_metrics) { // back to the grid of the handle's own configuration
        hh->mgrid_set = false;
        return NUFI_B200_OK;
    }
    if (conf_metrics->Nx == 0 || conf_metrics->Nu == 0 || conf_metrics->Nx >= (1ull << 31) || conf_metrics->Nu >= (1ull << 31))
        return fail(hh, NUFI_B200_ERR_ARG, "metrics grid: Nx and Nu must be positive (and below 2^31)");
    hh->mconf = *conf_metrics;
    hh->mgrid_set = true;
    return NUFI_B200_OK;
}

int nufi_b200_download_metrics(nufi_b200_handle *h, double *m)
{
    ENTER(h);
    if (!m) return fail(hh, NUFI_B200_ERR_ARG, "metrics is NULL");
    NUFI_CUDA_CHECK(hh, cudaMemcpyAsync(hh->h_pinned, hh->d_metrics, sizeof(double) * 4, cudaMemcpyDeviceToHost, hh->stream));
    NUFI_CUDA_CHECK(hh, cudaStreamSynchronize(hh->stream));
    for (int i = 0; i < 4; ++i) m[i] += hh->h_pinned[i]; // accumulate (cuda_kernel.cu:186-188)
    return NUFI_B200_OK;
}

int nufi_b200_eval_rho_all(nufi_b200_handle *h, size_t n, double *rho_host)
{
    ENTER(h);
    if (n > hh->Nt) return fail(hh, NUFI_B200_ERR_RANGE, "Time-step out of range.");
    int rc = check_levels(hh, n, "eval_rho_all");
    if (rc) return rc;
    rc = launch_backtrace(hh, n, 0, hh->n_nodes * hh->n_vel, false);
    if (rc) return rc;
    if (rho_host) {
        NUFI_CUDA_CHECK(hh, cudaMemcpyAsync(hh->h_pinned, hh->d_rho_full, sizeof(double) * hh->n_nodes, cudaMemcpyDeviceToHost, hh->stream));
        NUFI_CUDA_CHECK(hh, cudaStreamSynchronize(hh->stream));
        std::memcpy(rho_host, hh->h_pinned, sizeof(double) * hh->n_nodes);
    }
    return NUFI_B200_OK;
}

int nufi_b200_download_rho_full(nufi_b200_handle *h, double *rho_host)
{
    ENTER(h);
    if (!rho_host) return fail(hh, NUFI_B200_ERR_ARG, "rho_host is NULL");
    NUFI_CUDA_CHECK(hh, cudaMemcpyAsync(hh->h_pinned, hh->d_rho_full, sizeof(double) * hh->n_nodes, cudaMemcpyDeviceToHost, hh->stream));
    NUFI_CUDA_CHECK(hh, cudaStreamSynchronize(hh->stream));
    std::memcpy(rho_host, hh->h_pinned, sizeof(double) * hh->n_nodes);
    return NUFI_B200_OK;
}

int nufi_b200_solve_interpolate(nufi_b200_handle *h, size_t n, double *energy)
{
    ENTER(h);
    if (n > hh->Nt) return fail(hh, NUFI_B200_ERR_RANGE, "Time-step out of range.");
    int rc = tail_run(hh, n, hh->d_rho_full);
    if (rc) return rc;
    if (energy) {
        NUFI_CUDA_CHECK(hh, cudaMemcpyAsync(hh->h_pinned, hh->d_energy + n, sizeof(double), cudaMemcpyDeviceToHost, hh->stream));
        NUFI_CUDA_CHECK(hh, cudaStreamSynchronize(hh->stream));
        *energy = hh->h_pinned[0];
    }
    return NUFI_B200_OK;
}

int nufi_b200_solve_interpolate_host(nufi_b200_handle *h, size_t n, const double *rho_host, double *energy)
{
    ENTER(h);
    if (!rho_host) return fail(hh, NUFI_B200_ERR_ARG, "rho_host is NULL");
    int rc = stage_upload(hh, hh->d_rho_full, rho_host, hh->n_nodes);
    if (rc) return rc;
    double e = 0;
    rc = nufi_b200_solve_interpolate(h, n, energy ? &e : nullptr);
    if (rc) return rc;
    if (energy) *energy = e;
    return NUFI_B200_OK;
}

int nufi_b200_poisson_solve(nufi_b200_handle *h, double *data_host, double *energy)
{
    ENTER(h);
    if (!data_host) return fail(hh, NUFI_B200_ERR_ARG, "data is NULL");
    std::memcpy(hh->h_pinned, data_host, sizeof(double) * hh->n_nodes);
    NUFI_CUDA_CHECK(hh, cudaMemcpyAsync(hh->d_rho_full, hh->h_pinned, sizeof(double) * hh->n_nodes, cudaMemcpyHostToDevice, hh->stream));
    int rc = tail_filter(hh, hh->d_rho_full, 1);
    if (rc) return rc;
    NUFI_CUDA_CHECK(hh, cudaStreamSynchronize(hh->stream)); // h_pinned is reused for the way back
    NUFI_CUDA_CHECK(hh, cudaMemcpyAsync(hh->h_pinned, hh->d_field, sizeof(double) * hh->n_nodes, cudaMemcpyDeviceToHost, hh->stream));
    NUFI_CUDA_CHECK(hh, cudaStreamSynchronize(hh->stream));
    std::memcpy(data_host, hh->h_pinned, sizeof(double) * hh->n_nodes);
    if (energy) {
        NUFI_CUDA_CHECK(hh, cudaMemcpyAsync(hh->h_pinned, tail_energy_scratch(hh), sizeof(double), cudaMemcpyDeviceToHost, hh->stream));
        NUFI_CUDA_CHECK(hh, cudaStreamSynchronize(hh->stream));
        *energy = hh->h_pinned[0];
    }
    return NUFI_B200_OK;
}

int nufi_b200_interpolate(nufi_b200_handle *h, const double *values_host, double *coeffs_level_host)
{
    ENTER(h);
    if (!values_host || !coeffs_level_host) return fail(hh, NUFI_B200_ERR_ARG, "values / coeffs is NULL");
    std::memcpy(hh->h_pinned, values_host, sizeof(double) * hh->n_nodes);
    NUFI_CUDA_CHECK(hh, cudaMemcpyAsync(hh->d_rho_full, hh->h_pinned, sizeof(double) * hh->n_nodes, cudaMemcpyHostToDevice, hh->stream));
    int rc = tail_filter(hh, hh->d_rho_full, 2);
    if (rc) return rc;
    rc = expand_field_to_stage(hh);
    if (rc) return rc;
    NUFI_CUDA_CHECK(hh, cudaStreamSynchronize(hh->stream));
    NUFI_CUDA_CHECK(hh, cudaMemcpyAsync(hh->h_pinned, hh->d_stage, sizeof(double) * hh->stride_t, cudaMemcpyDeviceToHost, hh->stream));
    NUFI_CUDA_CHECK(hh, cudaStreamSynchronize(hh->stream));
    std::memcpy(coeffs_level_host, hh->h_pinned, sizeof(double) * hh->stride_t);
    return NUFI_B200_OK;
}

int nufi_b200_step(nufi_b200_handle *h, size_t n)
{
    ENTER(h);
    if (n > hh->Nt) return fail(hh, NUFI_B200_ERR_RANGE, "Time-step out of range.");
    int rc = check_levels(hh, n, "step");
    if (rc) return rc;
    rc = launch_backtrace(hh, n, 0, hh->n_nodes * hh->n_vel, false, /*defer_finish=*/true);
    if (rc) return rc;
    return tail_run(hh, n, nullptr);
}

// One time step for a caller that owns the history on the HOST (the reference drivers' loop body, bin/test_nufi_gpu_3d.cpp:
// 154-162, in one call): level n-1 host -> device (the only level the device has not seen: levels below were pushed by
// earlier calls), fused step on the device, level n + rho + energy device -> host, ONE stream synchronisation.
int nufi_b200_step_host(nufi_b200_handle *h, size_t n, double *coeffs_base, double *rho_host, double *energy, int peer)
{
    ENTER(h);
    if (n > hh->Nt) return fail(hh, NUFI_B200_ERR_RANGE, "Time-step out of range.");
    if (!coeffs_base) return fail(hh, NUFI_B200_ERR_ARG, "coeffs is NULL");
    int rc;
    if (n > 0) {
        rc = nufi_b200_upload_phi(h, n - 1, coeffs_base);
        if (rc) return rc;
    }
    rc = peer ? nufi_b200_peer_step(h, n) : nufi_b200_step(h, n);
    if (rc) return rc;
    rc = convert_level_from_device(hh, n, hh->d_stage);
    if (rc) return rc;
    double *lvl = hh->h_pinned, *rho = lvl + hh->stride_t, *en = rho + hh->n_nodes;
    NUFI_CUDA_CHECK(hh, cudaMemcpyAsync(lvl, hh->d_stage, sizeof(double) * hh->stride_t, cudaMemcpyDeviceToHost, hh->stream));
    if (rho_host) NUFI_CUDA_CHECK(hh, cudaMemcpyAsync(rho, hh->d_rho_full, sizeof(double) * hh->n_nodes, cudaMemcpyDeviceToHost, hh->stream));
    NUFI_CUDA_CHECK(hh, cudaMemcpyAsync(en, hh->d_energy + n, sizeof(double), cudaMemcpyDeviceToHost, hh->stream));
    NUFI_CUDA_CHECK(hh, cudaStreamSynchronize(hh->stream));
    std::memcpy(coeffs_base + n * hh->stride_t, lvl, sizeof(double) * hh->stride_t);
    if (rho_host) std::memcpy(rho_host, rho, sizeof(double) * hh->n_nodes);
    if (energy) *energy = *en;
    return NUFI_B200_OK;
}

int nufi_b200_download_energy(nufi_b200_handle *h, size_t n_begin, size_t n_end, double *energies)
{
    ENTER(h);
    if (n_begin > n_end || n_end > hh->Nt + 1) return fail(hh, NUFI_B200_ERR_RANGE, "energy range out of bounds");
    if (n_begin == n_end) return NUFI_B200_OK;
    if (!energies) return fail(hh, NUFI_B200_ERR_ARG, "energies is NULL");
    NUFI_CUDA_CHECK(hh, cudaMemcpyAsync(hh->h_pinned, hh->d_energy + n_begin, sizeof(double) * (n_end - n_begin), cudaMemcpyDeviceToHost, hh->stream));
    NUFI_CUDA_CHECK(hh, cudaStreamSynchronize(hh->stream));
    std::memcpy(energies, hh->h_pinned, sizeof(double) * (n_end - n_begin));
    return NUFI_B200_OK;
}

int nufi_b200_download_phi(nufi_b200_handle *h, size_t n, double *coeffs_level)
{
    ENTER(h);
    if (n > hh->Nt) return fail(hh, NUFI_B200_ERR_RANGE, "Time-step out of range.");
    if (!coeffs_level) return fail(hh, NUFI_B200_ERR_ARG, "coeffs_level is NULL");
    int rc = convert_level_from_device(hh, n, hh->d_stage);
    if (rc) return rc;
    NUFI_CUDA_CHECK(hh, cudaMemcpyAsync(hh->h_pinned, hh->d_stage, sizeof(double) * hh->stride_t, cudaMemcpyDeviceToHost, hh->stream));
    NUFI_CUDA_CHECK(hh, cudaStreamSynchronize(hh->stream));
    std::memcpy(coeffs_level, hh->h_pinned, sizeof(double) * hh->stride_t);
    return NUFI_B200_OK;
}

// ---- bulk history transfer (checkpoint / restart): the coefficient history is the complete simulation state
int nufi_b200_download_history(nufi_b200_handle *h, size_t n_levels, double *coeffs_host)
{
    ENTER(h);
    if (n_levels > hh->Nt + 1) return fail(hh, NUFI_B200_ERR_RANGE, "more levels requested than the history holds");
    if (n_levels && !coeffs_host) return fail(hh, NUFI_B200_ERR_ARG, "coeffs is NULL");
    int rc = check_levels(hh, n_levels, "download_history");
    if (rc) return rc;
    for (size_t m = 0; m < n_levels; ++m) {
        rc = nufi_b200_download_phi(h, m, coeffs_host + m * hh->stride_t);
        if (rc) return rc;
    }
    return NUFI_B200_OK;
}

int nufi_b200_upload_history(nufi_b200_handle *h, size_t n_levels, const double *coeffs_host)
{
    ENTER(h);
    if (n_levels > hh->Nt + 1) return fail(hh, NUFI_B200_ERR_RANGE, "more levels than the history can hold");
    if (n_levels && !coeffs_host) return fail(hh, NUFI_B200_ERR_ARG, "coeffs is NULL");
    for (size_t m = 0; m < n_levels; ++m) {
        int rc = nufi_b200_upload_phi(h, m, coeffs_host);
        if (rc) return rc;
    }
    return NUFI_B200_OK;
}

// ---- sampling for plots / diagnostics (host buffers, blocking)
namespace
{
struct DeviceScratch
{
    double *p = nullptr;
    ~DeviceScratch() { cudaFree(p); }
};
} // namespace

int nufi_b200_eval_f(nufi_b200_handle *h, size_t n, size_t npts, const double *points_host, double *f_host, int with_first_half_kick)
{
    ENTER(h);
    if (n > hh->Nt) return fail(hh, NUFI_B200_ERR_RANGE, "Time-step out of range.");
    if (npts == 0) return NUFI_B200_OK;
    if (!points_host || !f_host) return fail(hh, NUFI_B200_ERR_ARG, "points / f is NULL");
    int rc = check_levels(hh, with_first_half_kick ? (n == 0 ? 0 : n + 1) : n, "eval_f");
    if (rc) return rc;
    const size_t w = 2 * static_cast<size_t>(hh->dim);
    DeviceScratch d;
    if (cudaMalloc(&d.p, sizeof(double) * npts * (w + 1)) != cudaSuccess) return fail(hh, NUFI_B200_ERR_ALLOC, "cudaMalloc of the sample points failed");
    NUFI_CUDA_CHECK(hh, cudaMemcpyAsync(d.p, points_host, sizeof(double) * npts * w, cudaMemcpyHostToDevice, hh->stream));
    rc = launch_sample_f(hh, n, npts, d.p, d.p + npts * w, with_first_half_kick != 0);
    if (rc) return rc;
    NUFI_CUDA_CHECK(hh, cudaMemcpyAsync(f_host, d.p + npts * w, sizeof(double) * npts, cudaMemcpyDeviceToHost, hh->stream));
    NUFI_CUDA_CHECK(hh, cudaStreamSynchronize(hh->stream));
    return NUFI_B200_OK;
}

int nufi_b200_eval_phase_flow(nufi_b200_handle *h, size_t n, size_t npts, const double *points_host, double *feet_host)
{
    ENTER(h);
    if (n > hh->Nt) return fail(hh, NUFI_B200_ERR_RANGE, "Time-step out of range.");
    if (npts == 0) return NUFI_B200_OK;
    if (!points_host || !feet_host) return fail(hh, NUFI_B200_ERR_ARG, "points / feet is NULL");
    const bool traced = n > 1; // the reference traces nothing for n <= 1 (nufi/rho.hpp:102) and only reduces x into the box
    int rc = check_levels(hh, traced ? n + 1 : 0, "eval_phase_flow");
    if (rc) return rc;
    const size_t w = 2 * static_cast<size_t>(hh->dim);
    DeviceScratch d;
    if (cudaMalloc(&d.p, sizeof(double) * npts * 2 * w) != cudaSuccess) return fail(hh, NUFI_B200_ERR_ALLOC, "cudaMalloc of the sample points failed");
    NUFI_CUDA_CHECK(hh, cudaMemcpyAsync(d.p, points_host, sizeof(double) * npts * w, cudaMemcpyHostToDevice, hh->stream));
    // n <= 1: the kernel runs with no level to read (n = 0, eval_ftilda form) and just locates / re-assembles the point
    rc = launch_sample_f(hh, traced ? n : 0, npts, d.p, d.p + npts * w, /*full=*/traced, /*feet=*/true);
    if (rc) return rc;
    NUFI_CUDA_CHECK(hh, cudaMemcpyAsync(feet_host, d.p + npts * w, sizeof(double) * npts * w, cudaMemcpyDeviceToHost, hh->stream));
    NUFI_CUDA_CHECK(hh, cudaStreamSynchronize(hh->stream));
    return NUFI_B200_OK;
}

int nufi_b200_eval_field(nufi_b200_handle *h, size_t n, int derivative_axis, size_t npts, const double *points_host, double *values_host)
{
    ENTER(h);
    if (n > hh->Nt) return fail(hh, NUFI_B200_ERR_RANGE, "Time-step out of range.");
    if (derivative_axis < -1 || derivative_axis >= hh->dim) return fail(hh, NUFI_B200_ERR_ARG, "derivative axis must be -1 (value) or < dim");
    if (npts == 0) return NUFI_B200_OK;
    if (!points_host || !values_host) return fail(hh, NUFI_B200_ERR_ARG, "points / values is NULL");
    if (!hh->level_valid[n]) return fail(hh, NUFI_B200_ERR_RANGE, "eval_field: level was never uploaded or computed");
    const size_t w = static_cast<size_t>(hh->dim);
    DeviceScratch d;
    if (cudaMalloc(&d.p, sizeof(double) * npts * (w + 1)) != cudaSuccess) return fail(hh, NUFI_B200_ERR_ALLOC, "cudaMalloc of the sample points failed");
    NUFI_CUDA_CHECK(hh, cudaMemcpyAsync(d.p, points_host, sizeof(double) * npts * w, cudaMemcpyHostToDevice, hh->stream));
    const double *level = hh->d_raw ? hh->d_raw + n * hh->raw_stride : hh->d_hist + n * hh->level_stride; // reference layout
    int rc = launch_sample_field(hh, level, derivative_axis, npts, d.p, d.p + npts * w);
    if (rc) return rc;
    NUFI_CUDA_CHECK(hh, cudaMemcpyAsync(values_host, d.p + npts * w, sizeof(double) * npts, cudaMemcpyDeviceToHost, hh->stream));
    NUFI_CUDA_CHECK(hh, cudaStreamSynchronize(hh->stream));
    return NUFI_B200_OK;
}

int nufi_b200_sync(nufi_b200_handle *h)
{
    ENTER(h);
    NUFI_CUDA_CHECK(hh, cudaStreamSynchronize(hh->stream));
    return NUFI_B200_OK;
}

int nufi_b200_set_stream(nufi_b200_handle *h, void *cuda_stream)
{
    ENTER(h);
    NUFI_CUDA_CHECK(hh, cudaStreamSynchronize(hh->stream));
    {
        int rc = ev_drain(hh);
        if (rc) return rc;
    }
    hh->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : hh->own_stream;
    return NUFI_B200_OK;
}

int nufi_b200_rho_device(nufi_b200_handle *h, double **d_rho)
{
    ENTER(h);
    if (!d_rho) return fail(hh, NUFI_B200_ERR_ARG, "d_rho is NULL");
    *d_rho = hh->d_rho_partial;
    return NUFI_B200_OK;
}

int nufi_b200_field_tail_device(nufi_b200_handle *h, size_t n, const double *d_rho_partial_sum)
{
    ENTER(h);
    if (n > hh->Nt) return fail(hh, NUFI_B200_ERR_RANGE, "Time-step out of range.");
    if (!d_rho_partial_sum) return fail(hh, NUFI_B200_ERR_ARG, "d_rho_partial_sum is NULL");
    int rc = make_full_rho(hh, d_rho_partial_sum, hh->d_rho_full);
    if (rc) return rc;
    return tail_run(hh, n, hh->d_rho_full);
}

uint64_t nufi_b200_launch_count(const nufi_b200_handle *h) { return h ? H(h)->launches : 0; }

int nufi_b200_set_kernel_timing(nufi_b200_handle *h, int on)
{
    ENTER(h);
    NUFI_CUDA_CHECK(hh, cudaStreamSynchronize(hh->stream));
    int rc = ev_drain(hh);
    if (rc) return rc;
    hh->kernel_timing = on != 0;
    return NUFI_B200_OK;
}

int nufi_b200_last_backtrace_ms(nufi_b200_handle *h, float *ms)
{
    ENTER(h);
    if (!ms) return fail(hh, NUFI_B200_ERR_ARG, "ms is NULL");
    int rc = ev_drain(hh);
    if (rc) return rc;
    if (hh->bt_count == 0) return fail(hh, NUFI_B200_ERR_RANGE, "no timed backtrace launch yet (enable nufi_b200_set_kernel_timing first)");
    *ms = static_cast<float>(hh->bt_ms_last);
    return NUFI_B200_OK;
}

int nufi_b200_backtrace_time(nufi_b200_handle *h, double *total_ms, uint64_t *count, int reset)
{
    ENTER(h);
    int rc = ev_drain(hh);
    if (rc) return rc;
    if (total_ms) *total_ms = hh->bt_ms_total;
    if (count) *count = hh->bt_count;
    if (reset) {
        hh->bt_ms_total = 0;
        hh->bt_count = 0;
    }
    return NUFI_B200_OK;
}

const char *nufi_b200_last_variant(const nufi_b200_handle *h) { return h ? H(h)->last_variant : "none"; }

int nufi_b200_set_variant(nufi_b200_handle *h, int variant)
{
    Handle *hh = H(h);
    if (!hh) return fail(nullptr, NUFI_B200_ERR_ARG, "handle is NULL");
    if (variant < 0 || variant > 2) return fail(hh, NUFI_B200_ERR_ARG, "variant must be 0 (auto), 1 (global) or 2 (staged)");
    hh->variant_force = variant;
    return NUFI_B200_OK;
}

int nufi_b200_set_tile_nodes(nufi_b200_handle *h, int nodes_per_tile)
{
    Handle *hh = H(h);
    if (!hh) return fail(nullptr, NUFI_B200_ERR_ARG, "handle is NULL");
    const int t = nodes_per_tile;
    if (!(t == 0 || t == 1 || t == 2 || t == 4 || t == 8 || t == 16 || t == 32))
        return fail(hh, NUFI_B200_ERR_ARG, "nodes per tile must be 0 (automatic), 1, 2, 4, 8, 16 or 32");
    hh->tn_force = t;
    return NUFI_B200_OK;
}

int nufi_b200_set_tail_variant(nufi_b200_handle *h, int variant)
{
    Handle *hh = H(h);
    if (!hh) return fail(nullptr, NUFI_B200_ERR_ARG, "handle is NULL");
    if (variant < 0 || variant > 2) return fail(hh, NUFI_B200_ERR_ARG, "tail variant must be 0 (auto), 1 (cuFFT) or 2 (fused single CTA)");
    hh->tail_force = variant;
    return NUFI_B200_OK;
}

const char *nufi_b200_last_tail_variant(const nufi_b200_handle *h) { return h ? H(h)->last_tail : "none"; }

int nufi_b200_measure_fp64_peak(int device, double *tflops) { return measure_fp64_peak(device, tflops); }

int nufi_b200_device_count(int *count)
{
    if (!count) return fail(nullptr, NUFI_B200_ERR_ARG, "count is NULL");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(nullptr, NUFI_B200_ERR_CUDA, std::string(cudaGetErrorName(e)) + ": " + cudaGetErrorString(e));
    }
    return NUFI_B200_OK;
}

int nufi_b200_device_of(const nufi_b200_handle *h) { return h ? H(h)->device : -1; }

const char *nufi_b200_version(void) { return "nufi_b200 0.2 (sm_100a)"; }

} // extern "C"
