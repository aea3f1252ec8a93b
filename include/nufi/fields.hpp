// nufi/fields.hpp -- interpolate and eval with the reference's signatures (nufi/fields.hpp:36-61, 63-142, 149-184,
// 186-300, 308-350, 352-490).
//
// interpolate<real,order>(coeffs_level, values, conf): nodal values -> one level of periodic B-spline coefficients with
// the (order-1) halo.  The reference solves the collocation system with LSMR on the host; here it is an exact circulant
// solve on the device (cuFFT), agreeing with LSMR to ~1e-13.
// eval<real,order,dx[,dy,dz]>(x..., coeffs_level, conf): host-side evaluation of the spline or any of its derivatives at a
// point -- the drivers use it for plots and diagnostics only (bin/test_nufi_cpu_1d.cpp:82-119, and
// bin/test_nufi_gpu_1d.cpp:301 with dx = 2 for rho = -phi'').  Any order 1..8, any derivative order per dimension.
#ifndef NUFI_B200_NUFI_FIELDS_HPP
#define NUFI_B200_NUFI_FIELDS_HPP

#include <cmath>

#include "device_context.hpp"

namespace nufi
{

namespace detail
{

// Values (der = 0) or der-th derivatives of the `order` B-splines that overlap a cell, at the reference coordinate
// t in [0,1) of the cell (what nufi/splines.hpp:39-79 returns).  Uniform knots: with B_p[i] = M_p(t + p - i), M_p the
// cardinal B-spline of degree p,  B_p[i] = ((t+p-i) B_{p-1}[i-1] + (1+i-t) B_{p-1}[i]) / p,  and the der-th derivative of
// degree p is the der-fold backward difference of degree p - der.
template <size_t order> inline void bspline_basis(double t, size_t der, double *N)
{
    static_assert(order >= 1 && order <= 8, "spline order 1..8");
    const int K = static_cast<int>(order);
    if (der >= order) {
        for (int i = 0; i < K; ++i) N[i] = 0;
        return;
    }
    double b[order + 1] = {};
    b[0] = 1;
    const int deg = K - 1 - static_cast<int>(der); // degree reached by the recurrence
    for (int p = 1; p <= deg; ++p) {
        b[p] = t * b[p - 1] / p;
        for (int i = p - 1; i >= 1; --i) b[i] = ((t + (p - i)) * b[i - 1] + ((1 + i) - t) * b[i]) / p;
        b[0] = (1 - t) * b[0] / p;
    }
    for (int len = deg + 1; len < K; ++len) { // one difference per derivative: `len` values -> len + 1
        b[len] = b[len - 1];
        for (int i = len - 1; i >= 1; --i) b[i] = b[i - 1] - b[i];
        b[0] = -b[0];
    }
    for (int i = 0; i < K; ++i) N[i] = b[i];
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

inline double ipow(double b, size_t e)
{
    double r = 1;
    for (size_t i = 0; i < e; ++i) r *= b;
    return r;
}

} // namespace detail

namespace dim1
{
template <typename real, size_t order> void interpolate(real *coeffs, const real *values, const config_t<real> &conf)
{
    static_assert(std::is_same<real, double>::value, "libnufi_b200 computes in FP64");
    detail::context<config_t<real>, order>(conf)->interpolate(values, coeffs);
}
template <typename real, size_t order, size_t dx = 0> real eval(real x, const real *coeffs, const config_t<real> &conf)
{
    size_t k; double t, N[order];
    detail::locate(x, conf.x_min, conf.Lx, conf.Lx_inv, conf.dx_inv, conf.Nx, k, t);
    detail::bspline_basis<order>(t, dx, N);
    double r = 0;
    for (size_t a = 0; a < order; ++a) r += coeffs[k + a] * N[a];
    return r * detail::ipow(conf.dx_inv, dx);
}
} // namespace dim1

namespace dim2
{
template <typename real, size_t order> void interpolate(real *coeffs, const real *values, const config_t<real> &conf)
{
    static_assert(std::is_same<real, double>::value, "libnufi_b200 computes in FP64");
    detail::context<config_t<real>, order>(conf)->interpolate(values, coeffs);
}
template <typename real, size_t order, size_t dx = 0, size_t dy = 0> real eval(real x, real y, const real *coeffs, const config_t<real> &conf)
{
    size_t kx, ky; double tx, ty, Nx[order], Ny[order];
    detail::locate(x, conf.x_min, conf.Lx, conf.Lx_inv, conf.dx_inv, conf.Nx, kx, tx);
    detail::locate(y, conf.y_min, conf.Ly, conf.Ly_inv, conf.dy_inv, conf.Ny, ky, ty);
    detail::bspline_basis<order>(tx, dx, Nx);
    detail::bspline_basis<order>(ty, dy, Ny);
    const size_t sy = conf.Nx + order - 1;
    double r = 0;
    for (size_t b = 0; b < order; ++b) {
        double row = 0;
        for (size_t a = 0; a < order; ++a) row += coeffs[(ky + b) * sy + kx + a] * Nx[a];
        r += row * Ny[b];
    }
    return r * detail::ipow(conf.dx_inv, dx) * detail::ipow(conf.dy_inv, dy);
}
} // namespace dim2

namespace dim3
{
template <typename real, size_t order> void interpolate(real *coeffs, const real *values, const config_t<real> &conf)
{
    static_assert(std::is_same<real, double>::value, "libnufi_b200 computes in FP64");
    detail::context<config_t<real>, order>(conf)->interpolate(values, coeffs);
}
template <typename real, size_t order, size_t dx = 0, size_t dy = 0, size_t dz = 0>
real eval(real x, real y, real z, const real *coeffs, const config_t<real> &conf)
{
    size_t kx, ky, kz; double tx, ty, tz, Nx[order], Ny[order], Nz[order];
    detail::locate(x, conf.x_min, conf.Lx, conf.Lx_inv, conf.dx_inv, conf.Nx, kx, tx);
    detail::locate(y, conf.y_min, conf.Ly, conf.Ly_inv, conf.dy_inv, conf.Ny, ky, ty);
    detail::locate(z, conf.z_min, conf.Lz, conf.Lz_inv, conf.dz_inv, conf.Nz, kz, tz);
    detail::bspline_basis<order>(tx, dx, Nx);
    detail::bspline_basis<order>(ty, dy, Ny);
    detail::bspline_basis<order>(tz, dz, Nz);
    const size_t sy = conf.Nx + order - 1, sz = sy * (conf.Ny + order - 1);
    double r = 0;
    for (size_t c = 0; c < order; ++c) {
        double plane = 0;
        for (size_t b = 0; b < order; ++b) {
            double row = 0;
            for (size_t a = 0; a < order; ++a) row += coeffs[(kz + c) * sz + (ky + b) * sy + kx + a] * Nx[a];
            plane += row * Ny[b];
        }
        r += plane * Nz[c];
    }
    return r * detail::ipow(conf.dx_inv, dx) * detail::ipow(conf.dy_inv, dy) * detail::ipow(conf.dz_inv, dz);
}
} // namespace dim3

} // namespace nufi

#endif
