// nufi/fields.hpp -- interpolate and eval with the reference's signatures (nufi/fields.hpp:36-61, 63-142, 149-184,
// 186-300, 308-350, 352-490).
//
// interpolate<real,order>(coeffs_level, values, conf): nodal values -> one level of periodic B-spline coefficients with
// the (order-1) halo.  The reference solves the collocation system with LSMR on the host; here it is an exact circulant
// solve on the device (cuFFT), agreeing with LSMR to ~1e-13.
// eval<real,order,dx[,dy,dz]>(x..., coeffs_level, conf): host-side evaluation of the spline or one first derivative at a
// point -- the drivers use it for plots and diagnostics only (bin/test_nufi_cpu_1d.cpp:82-119); cubic, closed-form basis.
#ifndef NUFI_B200_NUFI_FIELDS_HPP
#define NUFI_B200_NUFI_FIELDS_HPP

#include <cmath>

#include "device_context.hpp"

namespace nufi
{

namespace detail
{

// cubic B-spline basis values (der = 0) or first derivatives (der = 1) at reference coordinate t in [0,1)
inline void basis4(double t, int der, double *N)
{
    const double s = 1 - t;
    if (der == 0) {
        N[0] = s * s * s / 6; N[3] = t * t * t / 6;
        N[1] = (3 * t * t * t - 6 * t * t + 4) / 6; N[2] = (3 * s * s * s - 6 * s * s + 4) / 6;
    } else {
        N[0] = -s * s / 2; N[3] = t * t / 2;
        N[1] = (3 * t * t - 4 * t) / 2; N[2] = -(3 * s * s - 4 * s) / 2;
    }
}
// periodic wrap + cell + reference coordinate, as nufi/fields.hpp:315-331
inline void locate(double x, double x_min, double L, double L_inv, double dx_inv, size_t N, size_t &k, double &t)
{
    x -= x_min;
    x -= L * std::floor(x * L_inv);
    const double kf = std::floor(x * dx_inv);
    k = static_cast<size_t>(kf);
    t = x * dx_inv - kf;
    if (k >= N) { k = 0; } // x rounded up to exactly L
}

} // namespace detail

namespace dim1
{
template <typename real, size_t order> void interpolate(real *coeffs, const real *values, const config_t<real> &conf)
{
    static_assert(std::is_same<real, double>::value && order == 4, "libnufi_b200: FP64, cubic");
    auto &k = detail::context<config_t<real>, order>(conf).kernel();
    cuda::check(nufi_b200_interpolate(k.handle(), values, coeffs), nufi_b200_last_error(k.handle()));
}
template <typename real, size_t order, size_t dx = 0> real eval(real x, const real *coeffs, const config_t<real> &conf)
{
    static_assert(order == 4 && dx <= 1, "host eval: cubic, value or first derivative");
    size_t k; double t, N[4];
    detail::locate(x, conf.x_min, conf.Lx, conf.Lx_inv, conf.dx_inv, conf.Nx, k, t);
    detail::basis4(t, dx, N);
    double r = 0;
    for (int a = 0; a < 4; ++a) r += coeffs[k + a] * N[a];
    return dx ? r * conf.dx_inv : r;
}
} // namespace dim1

namespace dim2
{
template <typename real, size_t order> void interpolate(real *coeffs, const real *values, const config_t<real> &conf)
{
    static_assert(std::is_same<real, double>::value && order == 4, "libnufi_b200: FP64, cubic");
    auto &k = detail::context<config_t<real>, order>(conf).kernel();
    cuda::check(nufi_b200_interpolate(k.handle(), values, coeffs), nufi_b200_last_error(k.handle()));
}
template <typename real, size_t order, size_t dx = 0, size_t dy = 0> real eval(real x, real y, const real *coeffs, const config_t<real> &conf)
{
    static_assert(order == 4 && dx + dy <= 1, "host eval: cubic, value or one first derivative");
    size_t kx, ky; double tx, ty, Nx[4], Ny[4];
    detail::locate(x, conf.x_min, conf.Lx, conf.Lx_inv, conf.dx_inv, conf.Nx, kx, tx);
    detail::locate(y, conf.y_min, conf.Ly, conf.Ly_inv, conf.dy_inv, conf.Ny, ky, ty);
    detail::basis4(tx, dx, Nx);
    detail::basis4(ty, dy, Ny);
    const size_t sy = conf.Nx + order - 1;
    double r = 0;
    for (int b = 0; b < 4; ++b) {
        double row = 0;
        for (int a = 0; a < 4; ++a) row += coeffs[(ky + b) * sy + kx + a] * Nx[a];
        r += row * Ny[b];
    }
    return r * (dx ? conf.dx_inv : 1) * (dy ? conf.dy_inv : 1);
}
} // namespace dim2

namespace dim3
{
template <typename real, size_t order> void interpolate(real *coeffs, const real *values, const config_t<real> &conf)
{
    static_assert(std::is_same<real, double>::value && order == 4, "libnufi_b200: FP64, cubic");
    auto &k = detail::context<config_t<real>, order>(conf).kernel();
    cuda::check(nufi_b200_interpolate(k.handle(), values, coeffs), nufi_b200_last_error(k.handle()));
}
template <typename real, size_t order, size_t dx = 0, size_t dy = 0, size_t dz = 0>
real eval(real x, real y, real z, const real *coeffs, const config_t<real> &conf)
{
    static_assert(order == 4 && dx + dy + dz <= 1, "host eval: cubic, value or one first derivative");
    size_t kx, ky, kz; double tx, ty, tz, Nx[4], Ny[4], Nz[4];
    detail::locate(x, conf.x_min, conf.Lx, conf.Lx_inv, conf.dx_inv, conf.Nx, kx, tx);
    detail::locate(y, conf.y_min, conf.Ly, conf.Ly_inv, conf.dy_inv, conf.Ny, ky, ty);
    detail::locate(z, conf.z_min, conf.Lz, conf.Lz_inv, conf.dz_inv, conf.Nz, kz, tz);
    detail::basis4(tx, dx, Nx);
    detail::basis4(ty, dy, Ny);
    detail::basis4(tz, dz, Nz);
    const size_t sy = conf.Nx + order - 1, sz = sy * (conf.Ny + order - 1);
    double r = 0;
    for (int c = 0; c < 4; ++c) {
        double plane = 0;
        for (int b = 0; b < 4; ++b) {
            double row = 0;
            for (int a = 0; a < 4; ++a) row += coeffs[(kz + c) * sz + (ky + b) * sy + kx + a] * Nx[a];
            plane += row * Ny[b];
        }
        r += plane * Nz[c];
    }
    return r * (dx ? conf.dx_inv : 1) * (dy ? conf.dy_inv : 1) * (dz ? conf.dz_inv : 1);
}
} // namespace dim3

} // namespace nufi

#endif
