/*
 * ref_harness.cpp -- compiles the REAL reference headers (read in place from /root/reference, never
 * copied) into oracle/_ref/libnufi_ref*.so so the C restatement in nufi_oracle.c can be pinned against
 * them and so bench.py --impl reference can time the reference's own eval_rho.  TEST INFRASTRUCTURE ONLY.
 *
 * What is the reference's own code here: nufi/config.hpp, splines.hpp, fields.hpp (eval + interpolate
 * via lsmr.hpp/.tpp), rho.hpp (eval_ftilda, eval_f, eval_rho).
 * What is NOT (dependencies absent from /root/reference and from this image):
 *   - OpenBLAS: the four BLAS-1 calls LSMR makes (nufi/blas.hpp:34-51) are plain loops below;
 *   - FFTW3: nufi/poisson.cpp cannot be compiled; ref_run_* uses oracle/nufi_oracle.c's DHT restatement
 *     for the solve stage (so the solve is a "port", everything else the reference).
 *
 * Two builds (oracle/Makefile):
 *   libnufi_ref_asis.so   f0 exactly as committed in nufi/config.hpp (two-stream / alpha=0.5 Landau /
 *                         bump-on-tail); ref_set_f0 is a no-op that reports failure.
 *   libnufi_ref.so        -DREF_F0_SELECTABLE: config_t<double>::f0 is explicitly specialised to a
 *                         run-time selectable expression (same formulas as oracle/nufi_oracle.c) so the
 *                         weak-Landau configurations of BASELINE.json can run through the reference code.
 */
#include <cmath>
#include <cstddef>
#include <cstring>
#include <memory>
#include <iostream>

#include <nufi/config.hpp>

#include "nufi_oracle.h"

#ifdef REF_F0_SELECTABLE
static orc_f0 g_f0[4] = {{0, {0, 0, 0, 0}}, {1, {0.01, 0.5, 0, 0}}, {0, {0.5, 0.5, 0, 0}}, {2, {0.03, 0.3, 0, 0}}};
namespace nufi
{
namespace dim1 { template <> double config_t<double>::f0(double x, double u) noexcept { return orc_f0_1d(&g_f0[1], x, u); } }
namespace dim2 { template <> double config_t<double>::f0(double x, double y, double u, double v) noexcept { return orc_f0_2d(&g_f0[2], x, y, u, v); } }
namespace dim3 { template <> double config_t<double>::f0(double x, double y, double z, double u, double v, double w) noexcept { return orc_f0_3d(&g_f0[3], x, y, z, u, v, w); } }
}
#endif

#include <nufi/fields.hpp>
#include <nufi/rho.hpp>

namespace nufi
{
namespace blas
{
double dot(const size_t n, const double *x, size_t incx, const double *y, size_t incy)
{
    double s = 0;
    for (size_t i = 0; i < n; ++i) s += x[i * incx] * y[i * incy];
    return s;
}
void axpy(size_t n, double alpha, const double *x, size_t incx, double *y, size_t incy)
{
    for (size_t i = 0; i < n; ++i) y[i * incy] += alpha * x[i * incx];
}
void scal(size_t n, double alpha, double *x, size_t incx)
{
    for (size_t i = 0; i < n; ++i) x[i * incx] *= alpha;
}
void copy(size_t n, const double *x, size_t incx, double *y, size_t incy)
{
    for (size_t i = 0; i < n; ++i) y[i * incy] = x[i * incx];
}
}
}

namespace
{
template <typename C, typename O> C to_ref(const O *o)
{
    static_assert(sizeof(C) == sizeof(O), "config mirror must match config_t<double> layout");
    C c;
    std::memcpy(static_cast<void *>(&c), o, sizeof(C));
    return c;
}
using c1 = nufi::dim1::config_t<double>;
using c2 = nufi::dim2::config_t<double>;
using c3 = nufi::dim3::config_t<double>;
}

extern "C" {

int ref_set_f0(int dim, const orc_f0 *f)
{
#ifdef REF_F0_SELECTABLE
    if (dim < 1 || dim > 3) return 1;
    g_f0[dim] = *f;
    return 0;
#else
    (void)dim; (void)f;
    return 1;
#endif
}

int ref_f0_selectable(void)
{
#ifdef REF_F0_SELECTABLE
    return 1;
#else
    return 0;
#endif
}

/* default-constructed reference configs (nufi/config.hpp:55-70,117-138,196-219) */
void ref_default_conf1d(orc_conf1d *o) { c1 c; std::memcpy(o, &c, sizeof(c)); }
void ref_default_conf2d(orc_conf2d *o) { c2 c; std::memcpy(o, &c, sizeof(c)); }
void ref_default_conf3d(orc_conf3d *o) { c3 c; std::memcpy(o, &c, sizeof(c)); }

void ref_basis4(int der, double x, double *out)
{
    if (der == 0) nufi::splines1d::N<double, 4, 0>(x, out);
    else nufi::splines1d::N<double, 4, 1>(x, out);
}

double ref_f0_1d(double x, double u) { return c1::f0(x, u); }
double ref_f0_2d(double x, double y, double u, double v) { return c2::f0(x, y, u, v); }
double ref_f0_3d(double x, double y, double z, double u, double v, double w) { return c3::f0(x, y, z, u, v, w); }

double ref_field_1d(int dx, double x, const double *level, const orc_conf1d *cf)
{
    c1 c = to_ref<c1>(cf);
    return dx ? nufi::dim1::eval<double, 4, 1>(x, level, c) : nufi::dim1::eval<double, 4, 0>(x, level, c);
}
double ref_field_2d(int dx, int dy, double x, double y, const double *level, const orc_conf2d *cf)
{
    c2 c = to_ref<c2>(cf);
    if (dx) return nufi::dim2::eval<double, 4, 1, 0>(x, y, level, c);
    if (dy) return nufi::dim2::eval<double, 4, 0, 1>(x, y, level, c);
    return nufi::dim2::eval<double, 4, 0, 0>(x, y, level, c);
}
double ref_field_3d(int dx, int dy, int dz, double x, double y, double z, const double *level, const orc_conf3d *cf)
{
    c3 c = to_ref<c3>(cf);
    if (dx) return nufi::dim3::eval<double, 4, 1, 0, 0>(x, y, z, level, c);
    if (dy) return nufi::dim3::eval<double, 4, 0, 1, 0>(x, y, z, level, c);
    if (dz) return nufi::dim3::eval<double, 4, 0, 0, 1>(x, y, z, level, c);
    return nufi::dim3::eval<double, 4, 0, 0, 0>(x, y, z, level, c);
}

double ref_ftilda_1d(size_t n, double x, double u, const double *coeffs, const orc_conf1d *cf)
{
    c1 c = to_ref<c1>(cf);
    return nufi::dim1::eval_ftilda<double, 4>(n, x, u, coeffs, c);
}
double ref_ftilda_2d(size_t n, double x, double y, double u, double v, const double *coeffs, const orc_conf2d *cf)
{
    c2 c = to_ref<c2>(cf);
    return nufi::dim2::eval_ftilda<double, 4>(n, x, y, u, v, coeffs, c);
}
double ref_ftilda_3d(size_t n, double x, double y, double z, double u, double v, double w, const double *coeffs,
                     const orc_conf3d *cf)
{
    c3 c = to_ref<c3>(cf);
    return nufi::dim3::eval_ftilda<double, 4>(n, x, y, z, u, v, w, coeffs, c);
}
double ref_f_1d(size_t n, double x, double u, const double *coeffs, const orc_conf1d *cf)
{
    c1 c = to_ref<c1>(cf);
    return nufi::dim1::eval_f<double, 4>(n, x, u, coeffs, c);
}
void ref_phase_flow_1d(size_t n, double *x, double *u, const double *coeffs, const orc_conf1d *cf)
{
    c1 c = to_ref<c1>(cf);
    nufi::dim1::eval_phase_flow<double, 4>(n, *x, *u, coeffs, c);
}
double ref_f_2d(size_t n, double x, double y, double u, double v, const double *coeffs, const orc_conf2d *cf)
{
    c2 c = to_ref<c2>(cf);
    return nufi::dim2::eval_f<double, 4>(n, x, y, u, v, coeffs, c);
}
double ref_f_3d(size_t n, double x, double y, double z, double u, double v, double w, const double *coeffs,
                const orc_conf3d *cf)
{
    c3 c = to_ref<c3>(cf);
    return nufi::dim3::eval_f<double, 4>(n, x, y, z, u, v, w, coeffs, c);
}

/* The reference drivers' OpenMP sweep (bin/test_nufi_cpu_{1,2,3}d.cpp) over nodes [l_begin,l_end). */
void ref_rho_sweep_1d(size_t n, const double *coeffs, const orc_conf1d *cf, size_t l_begin, size_t l_end, double *rho)
{
    c1 c = to_ref<c1>(cf);
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t l = l_begin; l < l_end; ++l) rho[l] = nufi::dim1::eval_rho<double, 4>(n, l, coeffs, c);
}
void ref_rho_sweep_2d(size_t n, const double *coeffs, const orc_conf2d *cf, size_t l_begin, size_t l_end, double *rho)
{
    c2 c = to_ref<c2>(cf);
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t l = l_begin; l < l_end; ++l) rho[l] = nufi::dim2::eval_rho<double, 4>(n, l, coeffs, c);
}
void ref_rho_sweep_3d(size_t n, const double *coeffs, const orc_conf3d *cf, size_t l_begin, size_t l_end, double *rho)
{
    c3 c = to_ref<c3>(cf);
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t l = l_begin; l < l_end; ++l) rho[l] = nufi::dim3::eval_rho<double, 4>(n, l, coeffs, c);
}

/* The reference's interpolate (LSMR) writing one level with halo. */
void ref_interpolate_1d(double *level, const double *values, const orc_conf1d *cf)
{
    c1 c = to_ref<c1>(cf);
    nufi::dim1::interpolate<double, 4>(level, values, c);
}
void ref_interpolate_2d(double *level, const double *values, const orc_conf2d *cf)
{
    c2 c = to_ref<c2>(cf);
    nufi::dim2::interpolate<double, 4>(level, values, c);
}
void ref_interpolate_3d(double *level, const double *values, const orc_conf3d *cf)
{
    c3 c = to_ref<c3>(cf);
    nufi::dim3::interpolate<double, 4>(level, values, c);
}

/* ---- the same reference templates at the other spline orders they are written for (nufi/splines.hpp:39-110 is generic; the
 *      reference's CUDA side instantiates 3..8, nufi/cuda_kernel.cu:191-203): pins the order-generic paths of the oracle. */
#define REF_ORDER_SWITCH(order, CALL) \
    switch (order) {                  \
    case 3: { constexpr size_t O = 3; CALL; } break; \
    case 4: { constexpr size_t O = 4; CALL; } break; \
    case 5: { constexpr size_t O = 5; CALL; } break; \
    case 6: { constexpr size_t O = 6; CALL; } break; \
    case 7: { constexpr size_t O = 7; CALL; } break; \
    case 8: { constexpr size_t O = 8; CALL; } break; \
    default: return 1;                \
    }

int ref_basis_order(int order, int der, double x, double *out)
{
    REF_ORDER_SWITCH(order, if (der == 0) (nufi::splines1d::N<double, O, 0>(x, out)); else if (der == 1) (nufi::splines1d::N<double, O, 1>(x, out)); else (nufi::splines1d::N<double, O, 2>(x, out)))
    return 0;
}
int ref_rho_sweep_order_1d(int order, size_t n, const double *coeffs, const orc_conf1d *cf, size_t l_begin, size_t l_end, double *rho)
{
    c1 c = to_ref<c1>(cf);
    REF_ORDER_SWITCH(order, _Pragma("omp parallel for schedule(dynamic, 1)") for (size_t l = l_begin; l < l_end; ++l) rho[l] = (nufi::dim1::eval_rho<double, O>(n, l, coeffs, c)))
    return 0;
}
int ref_rho_sweep_order_2d(int order, size_t n, const double *coeffs, const orc_conf2d *cf, size_t l_begin, size_t l_end, double *rho)
{
    c2 c = to_ref<c2>(cf);
    REF_ORDER_SWITCH(order, _Pragma("omp parallel for schedule(dynamic, 1)") for (size_t l = l_begin; l < l_end; ++l) rho[l] = (nufi::dim2::eval_rho<double, O>(n, l, coeffs, c)))
    return 0;
}
int ref_rho_sweep_order_3d(int order, size_t n, const double *coeffs, const orc_conf3d *cf, size_t l_begin, size_t l_end, double *rho)
{
    c3 c = to_ref<c3>(cf);
    REF_ORDER_SWITCH(order, _Pragma("omp parallel for schedule(dynamic, 1)") for (size_t l = l_begin; l < l_end; ++l) rho[l] = (nufi::dim3::eval_rho<double, O>(n, l, coeffs, c)))
    return 0;
}
int ref_interpolate_order_1d(int order, double *level, const double *values, const orc_conf1d *cf)
{
    c1 c = to_ref<c1>(cf);
    REF_ORDER_SWITCH(order, (nufi::dim1::interpolate<double, O>(level, values, c)))
    return 0;
}
int ref_interpolate_order_2d(int order, double *level, const double *values, const orc_conf2d *cf)
{
    c2 c = to_ref<c2>(cf);
    REF_ORDER_SWITCH(order, (nufi::dim2::interpolate<double, O>(level, values, c)))
    return 0;
}
int ref_interpolate_order_3d(int order, double *level, const double *values, const orc_conf3d *cf)
{
    c3 c = to_ref<c3>(cf);
    REF_ORDER_SWITCH(order, (nufi::dim3::interpolate<double, O>(level, values, c)))
    return 0;
}

/* The CPU drivers' time loop with the solve stage ported (FFTW absent): reference eval_rho ->
 * orc_poisson_* -> reference interpolate. */
void ref_run_1d(const orc_conf1d *cf, size_t n_begin, size_t n_end, double *coeffs, double *energy, double *rho_out)
{
    c1 c = to_ref<c1>(cf);
    const size_t stride_t = c.Nx + 3, N = c.Nx;
    std::unique_ptr<double[]> rho{new double[N]};
    for (size_t n = n_begin; n < n_end; ++n) {
        ref_rho_sweep_1d(n, coeffs, cf, 0, N, rho.get());
        if (rho_out && n + 1 == n_end) std::memcpy(rho_out, rho.get(), sizeof(double) * N);
        double e = orc_poisson_1d(cf, rho.get());
        if (energy) energy[n] = e;
        nufi::dim1::interpolate<double, 4>(coeffs + n * stride_t, rho.get(), c);
    }
}
void ref_run_2d(const orc_conf2d *cf, size_t n_begin, size_t n_end, double *coeffs, double *energy, double *rho_out)
{
    c2 c = to_ref<c2>(cf);
    const size_t stride_t = (c.Nx + 3) * (c.Ny + 3), N = c.Nx * c.Ny;
    std::unique_ptr<double[]> rho{new double[N]};
    for (size_t n = n_begin; n < n_end; ++n) {
        ref_rho_sweep_2d(n, coeffs, cf, 0, N, rho.get());
        if (rho_out && n + 1 == n_end) std::memcpy(rho_out, rho.get(), sizeof(double) * N);
        double e = orc_poisson_2d(cf, rho.get());
        if (energy) energy[n] = e;
        nufi::dim2::interpolate<double, 4>(coeffs + n * stride_t, rho.get(), c);
    }
}
void ref_run_3d(const orc_conf3d *cf, size_t n_begin, size_t n_end, double *coeffs, double *energy, double *rho_out)
{
    c3 c = to_ref<c3>(cf);
    const size_t stride_t = (c.Nx + 3) * (c.Ny + 3) * (c.Nz + 3), N = c.Nx * c.Ny * c.Nz;
    std::unique_ptr<double[]> rho{new double[N]};
    for (size_t n = n_begin; n < n_end; ++n) {
        ref_rho_sweep_3d(n, coeffs, cf, 0, N, rho.get());
        if (rho_out && n + 1 == n_end) std::memcpy(rho_out, rho.get(), sizeof(double) * N);
        double e = orc_poisson_3d(cf, rho.get());
        if (energy) energy[n] = e;
        nufi::dim3::interpolate<double, 4>(coeffs + n * stride_t, rho.get(), c);
    }
}

} /* extern "C" */
