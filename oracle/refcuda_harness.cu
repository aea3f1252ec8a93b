// refcuda_harness.cu -- C interface around the REAL reference CUDA path (nufi/cuda_kernel.cu, compiled unmodified and in place
// from /root/reference by oracle/Makefile; nothing of it is copied into this repository).  Test/bench infrastructure only: it is
// the informational "reference's existing CUDA path" column of the north star -- the same isolated step, CUDA-event timed on the
// same GPU -- and a second, GPU-side parity witness.  Never linked into libnufi_b200.
//
// f0 is the one COMMITTED in the reference's nufi/config.hpp (1d: two-stream alpha=0.01 k=0.5 -- exactly workload C2; 2d: Landau
// amplitude 0.5; 3d: bump-on-tail); it is evaluated once per point per step, so the kernel time does not depend on the choice.
#include <cstddef>
#include <cstdio>
#include <exception>
#include <memory>

#include <nufi/config.hpp>
#include <nufi/cuda_kernel.hpp>

namespace
{

template <typename Kernel, typename Conf> struct Box
{
    Conf conf;
    std::unique_ptr<Kernel> k;
};

// reps timed calls of compute_rho (memset + launch on the default stream, as the reference does), then one download
template <typename B> int timed_rho(B *b, size_t n, size_t q0, size_t q1, double *rho_accum, int reps, float *ms_per_call)
{
    try {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        b->k->compute_rho(n, q0, q1); // warm-up
        cudaDeviceSynchronize();
        cudaEventRecord(e0, 0);
        for (int r = 0; r < reps; ++r) b->k->compute_rho(n, q0, q1);
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms_per_call) *ms_per_call = ms / (reps > 0 ? reps : 1);
        if (rho_accum) b->k->download_rho(rho_accum);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        return cudaGetLastError() == cudaSuccess ? 0 : 2;
    } catch (const std::exception &ex) {
        std::fprintf(stderr, "refcuda: %s\n", ex.what());
        return 1;
    }
}

template <typename B> int upload(B *b, size_t n_levels, const double *coeffs)
{
    try {
        for (size_t n = 0; n < n_levels; ++n) b->k->upload_phi(n, coeffs);
        return 0;
    } catch (const std::exception &ex) {
        std::fprintf(stderr, "refcuda: %s\n", ex.what());
        return 1;
    }
}

} // namespace

#define REFCUDA_DIM(D)                                                                                                     \
    using Box##D = Box<nufi::dim##D::cuda_kernel<double, 4>, nufi::dim##D::config_t<double>>;                              \
    extern "C" void *refcuda_create_##D##d(const void *conf, int dev)                                                       \
    {                                                                                                                      \
        try {                                                                                                              \
            auto *b = new Box##D;                                                                                          \
            b->conf = *static_cast<const nufi::dim##D::config_t<double> *>(conf);                                          \
            b->k.reset(new nufi::dim##D::cuda_kernel<double, 4>(b->conf, dev));                                            \
            return b;                                                                                                      \
        } catch (const std::exception &ex) {                                                                               \
            std::fprintf(stderr, "refcuda: %s\n", ex.what());                                                              \
            return nullptr;                                                                                                \
        }                                                                                                                  \
    }                                                                                                                      \
    extern "C" void refcuda_destroy_##D##d(void *b) { delete static_cast<Box##D *>(b); }                                    \
    extern "C" int refcuda_upload_##D##d(void *b, size_t n_levels, const double *coeffs)                                    \
    {                                                                                                                      \
        return upload(static_cast<Box##D *>(b), n_levels, coeffs);                                                         \
    }                                                                                                                      \
    extern "C" int refcuda_rho_##D##d(void *b, size_t n, size_t q0, size_t q1, double *rho_accum, int reps, float *ms)      \
    {                                                                                                                      \
        return timed_rho(static_cast<Box##D *>(b), n, q0, q1, rho_accum, reps, ms);                                        \
    }

REFCUDA_DIM(1)
REFCUDA_DIM(2)
REFCUDA_DIM(3)
