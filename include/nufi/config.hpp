// nufi/config.hpp -- nufi::dim{1,2,3}::config_t<real> for the B200 library.
//
// Same public, mutable fields in the same order, same defaults and derived quantities as the reference's
// nufi/config.hpp:33-70, 91-138, 167-219, so config_t<double> is layout-identical to nufi_b200_config{1,2,3}d and a
// driver written against the reference compiles unchanged.  One deliberate difference: the reference picks the initial
// condition by editing the body of the static f0 and recompiling; here f0 is selected at run time through the static
// member `f0_sel` (kind + parameters, every expression of the reference's config.hpp is available, the default is
// the line the reference has active).  `f0` itself stays a static member with the reference's signature.
#ifndef NUFI_B200_NUFI_CONFIG_HPP
#define NUFI_B200_NUFI_CONFIG_HPP

#include <cmath>
#include <cstddef>

#include "../nufi_b200.h"

namespace nufi
{

namespace dim1
{

template <typename real> struct config_t
{
    size_t Nx, Nu, Nt;
    real dt;
    real x_min, x_max;
    real u_min, u_max;
    real dx, dx_inv, Lx, Lx_inv;
    real du;

    config_t() noexcept
    {
        Nx = 256; Nu = 512;
        u_min = -10; u_max = 10;
        x_min = 0; x_max = 4 * M_PI;
        dt = 1. / 16.; Nt = 100 / dt;
        derive();
    }
    // recompute the derived fields after editing the primary ones (the reference's tests do this by hand)
    void derive() noexcept
    {
        Lx = x_max - x_min; Lx_inv = 1 / Lx;
        dx = Lx / Nx; dx_inv = 1 / dx;
        du = (u_max - u_min) / Nu;
    }

    // kind 0: Landau (config.hpp:83), 1: two-stream (:82, active in the reference); p = {alpha, k}
    static inline nufi_b200_f0 f0_sel{1, {0.01, 0.5, 0, 0}};
    static real f0(real x, real u) noexcept
    {
        const real alpha = f0_sel.p[0], k = f0_sel.p[1];
        real r = 0.39894228040143267793994 * (1. + alpha * std::cos(k * x)) * std::exp(-u * u / 2.);
        return f0_sel.kind == 1 ? r * u * u : r;
    }
};

} // namespace dim1

namespace dim2
{

template <typename real> struct config_t
{
    size_t Nx, Ny, Nu, Nv, Nt;
    real dt;
    real x_min, x_max, y_min, y_max;
    real u_min, u_max, v_min, v_max;
    real dx, dx_inv, Lx, Lx_inv;
    real dy, dy_inv, Ly, Ly_inv;
    real du, dv;

    config_t() noexcept
    {
        Nx = Ny = 32; Nu = Nv = 128;
        u_min = v_min = -6; u_max = v_max = 6;
        x_min = y_min = 0; x_max = y_max = 4 * M_PI;
        dt = 1. / 16.; Nt = 50 / dt;
        derive();
    }
    void derive() noexcept
    {
        Lx = x_max - x_min; Lx_inv = 1 / Lx;
        Ly = y_max - y_min; Ly_inv = 1 / Ly;
        dx = Lx / Nx; dx_inv = 1 / dx;
        dy = Ly / Ny; dy_inv = 1 / dy;
        du = (u_max - u_min) / Nu;
        dv = (v_max - v_min) / Nv;
    }

    // kind 0: Landau (config.hpp:148-149, active: alpha = 0.5, k = 0.5), 1: two-stream (:151-158); p = {alpha, k, v0}
    static inline nufi_b200_f0 f0_sel{0, {0.5, 0.5, 2.4, 0}};
    static real f0(real x, real y, real u, real v) noexcept
    {
        const real alpha = f0_sel.p[0], k = f0_sel.p[1];
        const real pert = 1.0 + alpha * (std::cos(k * x) + std::cos(k * y));
        if (f0_sel.kind == 1) {
            const real v0 = f0_sel.p[2], c = 1.0 / (8.0 * M_PI);
            return c * pert * (std::exp(-0.5 * (v - v0) * (v - v0)) + std::exp(-0.5 * (v + v0) * (v + v0))) *
                   (std::exp(-0.5 * (u - v0) * (u - v0)) + std::exp(-0.5 * (u + v0) * (u + v0)));
        }
        return 1.0 / (2.0 * M_PI) * std::exp(-0.5 * (u * u + v * v)) * pert;
    }
};

} // namespace dim2

namespace dim3
{

template <typename real> struct config_t
{
    size_t Nx, Ny, Nz, Nu, Nv, Nw, Nt;
    real dt;
    real x_min, x_max, y_min, y_max, z_min, z_max;
    real u_min, u_max, v_min, v_max, w_min, w_max;
    real dx, dx_inv, Lx, Lx_inv;
    real dy, dy_inv, Ly, Ly_inv;
    real dz, dz_inv, Lz, Lz_inv;
    real du, dv, dw;

    config_t() noexcept
    {
        Nx = Ny = Nz = 8; Nu = Nv = Nw = 8;
        u_min = v_min = w_min = -9; u_max = v_max = w_max = 0; // as committed in the reference (config.hpp:201-203)
        x_min = y_min = z_min = 0; x_max = y_max = z_max = 20 * M_PI / 3.0;
        dt = 1. / 10.; Nt = 5 / dt;
        derive();
    }
    void derive() noexcept
    {
        Lx = x_max - x_min; Lx_inv = 1 / Lx;
        Ly = y_max - y_min; Ly_inv = 1 / Ly;
        Lz = z_max - z_min; Lz_inv = 1 / Lz;
        dx = Lx / Nx; dx_inv = 1 / dx;
        dy = Ly / Ny; dy_inv = 1 / dy;
        dz = Lz / Nz; dz_inv = 1 / dz;
        du = (u_max - u_min) / Nu;
        dv = (v_max - v_min) / Nv;
        dw = (w_max - w_min) / Nw;
    }

    // kind 0: Landau (config.hpp:233-234), 1: two-stream (:237-242), 2: bump-on-tail (:244-246, active); p = {alpha, k, v0}
    static inline nufi_b200_f0 f0_sel{2, {0.03, 0.3, 2.4, 0}};
    static real f0(real x, real y, real z, real u, real v, real w) noexcept
    {
        const real alpha = f0_sel.p[0], k = f0_sel.p[1];
        const real pert = 1 + alpha * (std::cos(k * x) + std::cos(k * y) + std::cos(k * z));
        if (f0_sel.kind == 1) {
            const real c = 0.03174681796712048489288165246732, v0 = f0_sel.p[2];
            return c * (std::exp(-(v - v0) * (v - v0) / 2.0) + std::exp(-(v + v0) * (v + v0) / 2.0)) * std::exp(-(u * u + w * w) / 2) * pert;
        }
        const real c = 0.06349363593424096978576330493464;
        if (f0_sel.kind == 2)
            return c * (0.9 * std::exp(-0.5 * u * u) + 0.2 * std::exp(-2 * (u - 4.5) * (u - 4.5))) * std::exp(-0.5 * (v * v + w * w)) * pert;
        return c * pert * std::exp(-(u * u + v * v + w * w) / 2);
    }
};

} // namespace dim3

} // namespace nufi

#endif
