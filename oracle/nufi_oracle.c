/*
 * nufi_oracle.c -- CPU restatement of the NuFI hot path.  TEST INFRASTRUCTURE ONLY
 * (see nufi_oracle.h for the rules and the list of deviations).
 *
 * Every function cites the reference lines it follows; the floating-point expression
 * order of the reference is kept (canonical build: gcc -O2 -ffp-contract=off, i.e. what the
 * reference's autotools build produces on x86-64: no FMA contraction).
 */
#include "nufi_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_ORDER 8

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* OpenMP thread count of the sweeps below: torchrun exports OMP_NUM_THREADS=1 to every rank, so the CPU-baseline legs of
 * bench.py set the count explicitly (to the host's core count) instead of inheriting it. */
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static double inv_factorial(int n) /* splines.hpp:31-37 */
{
    double f = 1.0;
    for (int i = 2; i <= n; ++i) f = (double)i * f; /* real(n)*faculty(n-1), innermost first */
    return 1.0 / f;
}

/* The reference's faculty() recursion multiplies n*(n-1)*...; for n <= 7 every product is an
 * exact integer in double, so evaluation order is irrelevant. */

/* splines.hpp:39-79  B-spline basis values (or der-th derivatives) on the reference cell. */
void orc_bspline_basis(int order, int der, double x, double *out)
{
    const int n = order, d = der;
    if (d >= n)
        for (int i = 0; i < n; ++i) out[i] = 0.0;
    if (n == 1) { out[0] = 1.0; return; }

    double v[ORC_MAX_ORDER];
    v[n - 1] = 1.0;
    for (int k = 1; k < n - d; ++k) {
        v[n - k - 1] = (1 - x) * v[n - k];
        for (int i = 1 - k; i < 0; ++i)
            v[n - 1 + i] = (x - i) * v[n - 1 + i] + (k + 1 + i - x) * v[n + i];
        v[n - 1] *= x;
    }
    for (int j = d; j-- > 0;) {
        v[j] = -v[j + 1];
        for (int i = j + 1; i < n - 1; ++i) v[i] = v[i] - v[i + 1];
    }
    const double factor = inv_factorial(n - d - 1);
    for (int i = 0; i < n; ++i) out[i] = v[i] * factor;
}

/* splines.hpp:81-110  de Boor evaluation of sum_i c_i N_i^(der)(x). */
double orc_deboor(int order, int der, double x, const double *coef, size_t stride)
{
    const size_t n = (size_t)order, d = (size_t)der;
    if (d >= n) return 0.0;
    if (n == 1) return coef[0];

    double c[ORC_MAX_ORDER];
    for (size_t j = 0; j < n; ++j) c[j] = coef[stride * j];
    for (size_t j = 1; j <= d; ++j)
        for (size_t i = n; i-- > j;) c[i] = c[i] - c[i - 1];
    for (size_t j = 1; j < n - d; ++j)
        for (size_t i = n - d; i-- > j;)
            c[d + i] = (x + n - d - 1 - i) * c[d + i] + (i - j + 1 - x) * c[d + i - 1];
    return inv_factorial((int)(n - d - 1)) * c[n - 1];
}

/* splines.hpp:117-137 */
static double spline2d(int order, int dx, int dy, double x, double y, const double *coef, size_t stride_y)
{
    if (dx >= order || dy >= order) return 0.0;
    if (order == 1) return coef[0];
    double c[ORC_MAX_ORDER] = {0}, Nx[ORC_MAX_ORDER];
    orc_bspline_basis(order, dx, x, Nx);
    for (int j = 0; j < order; ++j)
        for (int i = 0; i < order; ++i) c[j] += coef[(size_t)j * stride_y + (size_t)i] * Nx[i];
    return orc_deboor(order, dy, y, c, 1);
}

/* splines.hpp:144-172 */
static double spline3d(int order, int dx, int dy, int dz, double x, double y, double z,
                       const double *coef, size_t stride_z, size_t stride_y)
{
    if (dx >= order || dy >= order || dz >= order) return 0.0;
    if (order == 1) return coef[0];
    double czy[ORC_MAX_ORDER * ORC_MAX_ORDER] = {0}, cz[ORC_MAX_ORDER] = {0}, N[ORC_MAX_ORDER];
    orc_bspline_basis(order, dx, x, N);
    for (int k = 0; k < order; ++k)
        for (int j = 0; j < order; ++j)
            for (int i = 0; i < order; ++i)
                czy[k * order + j] += coef[(size_t)k * stride_z + (size_t)j * stride_y + (size_t)i] * N[i];
    orc_bspline_basis(order, dy, y, N);
    for (int k = 0; k < order; ++k)
        for (int j = 0; j < order; ++j) cz[k] += czy[k * order + j] * N[j];
    return orc_deboor(order, dz, z, cz, 1);
}

/* fields.hpp:36-61 */
double orc_field_1d(int order, int dx, double x, const double *level, const orc_conf1d *cf)
{
    x -= cf->x_min;
    x = x - cf->Lx * floor(x * cf->Lx_inv);
    double x_knot = floor(x * cf->dx_inv);
    size_t ii = (size_t)x_knot;
    x = x * cf->dx_inv - x_knot;
    double factor = 1;
    for (int i = 0; i < dx; ++i) factor *= cf->dx_inv;
    return factor * orc_deboor(order, dx, x, level + ii, 1);
}

/* fields.hpp:149-184 */
double orc_field_2d(int order, int dx, int dy, double x, double y, const double *level, const orc_conf2d *cf)
{
    x -= cf->x_min;
    y -= cf->y_min;
    x = x - cf->Lx * floor(x * cf->Lx_inv);
    y = y - cf->Ly * floor(y * cf->Ly_inv);
    double x_knot = floor(x * cf->dx_inv);
    double y_knot = floor(y * cf->dy_inv);
    size_t ii = (size_t)x_knot, jj = (size_t)y_knot;
    x = x * cf->dx_inv - x_knot;
    y = y * cf->dy_inv - y_knot;
    const size_t stride_y = cf->Nx + (size_t)order - 1;
    double factor = 1;
    for (int i = 0; i < dx; ++i) factor *= cf->dx_inv;
    for (int j = 0; j < dy; ++j) factor *= cf->dy_inv;
    return factor * spline2d(order, dx, dy, x, y, level + jj * stride_y + ii, stride_y);
}

/* fields.hpp:308-350 */
double orc_field_3d(int order, int dx, int dy, int dz, double x, double y, double z,
                    const double *level, const orc_conf3d *cf)
{
    x -= cf->x_min;
    y -= cf->y_min;
    z -= cf->z_min;
    x = x - cf->Lx * floor(x * cf->Lx_inv);
    y = y - cf->Ly * floor(y * cf->Ly_inv);
    z = z - cf->Lz * floor(z * cf->Lz_inv);
    double x_knot = floor(x * cf->dx_inv);
    double y_knot = floor(y * cf->dy_inv);
    double z_knot = floor(z * cf->dz_inv);
    size_t ii = (size_t)x_knot, jj = (size_t)y_knot, kk = (size_t)z_knot;
    x = x * cf->dx_inv - x_knot;
    y = y * cf->dy_inv - y_knot;
    z = z * cf->dz_inv - z_knot;
    const size_t stride_y = cf->Nx + (size_t)order - 1;
    const size_t stride_z = (cf->Ny + (size_t)order - 1) * stride_y;
    double factor = 1;
    for (int i = 0; i < dx; ++i) factor *= cf->dx_inv;
    for (int j = 0; j < dy; ++j) factor *= cf->dy_inv;
    for (int k = 0; k < dz; ++k) factor *= cf->dz_inv;
    return factor * spline3d(order, dx, dy, dz, x, y, z, level + kk * stride_z + jj * stride_y + ii,
                             stride_z, stride_y);
}

/* ------------------------------------------------------------------ f0 (config.hpp) */

double orc_f0_1d(const orc_f0 *f, double x, double u)
{
    const double alpha = f->p[0], k = f->p[1];
    if (f->kind == 1) /* config.hpp:82 */
        return 0.39894228040143267793994 * (1. + alpha * cos(k * x)) * exp(-u * u / 2.) * u * u;
    return 0.39894228040143267793994 * (1. + alpha * cos(k * x)) * exp(-u * u / 2); /* :83 */
}

double orc_f0_2d(const orc_f0 *f, double x, double y, double u, double v)
{
    const double alpha = f->p[0], k = f->p[1];
    if (f->kind == 1) { /* config.hpp:151-158 */
        const double v0 = f->p[2];
        const double c = 1.0 / (8.0 * M_PI);
        double pertube = 1.0 + alpha * (cos(k * x) + cos(k * y));
        double feq = (exp(-0.5 * (v - v0) * (v - v0)) + exp(-0.5 * (v + v0) * (v + v0))) *
                     (exp(-0.5 * (u - v0) * (u - v0)) + exp(-0.5 * (u + v0) * (u + v0)));
        return c * pertube * feq;
    }
    /* config.hpp:148-149 */
    return 1.0 / (2.0 * M_PI) * exp(-0.5 * (u * u + v * v)) * (1 + alpha * (cos(k * x) + cos(k * y)));
}

double orc_f0_3d(const orc_f0 *f, double x, double y, double z, double u, double v, double w)
{
    const double alpha = f->p[0], k = f->p[1];
    if (f->kind == 1) { /* config.hpp:237-242 */
        const double c = 0.03174681796712048489288165246732;
        const double v0 = f->p[2];
        return c * ((exp(-(v - v0) * (v - v0) / 2.0) + exp(-(v + v0) * (v + v0) / 2.0))) *
               exp(-(u * u + w * w) / 2) * (1 + alpha * (cos(k * x) + cos(k * y) + cos(k * z)));
    }
    if (f->kind == 2) { /* config.hpp:244-246 */
        const double c = 0.06349363593424096978576330493464;
        return c * (0.9 * exp(-0.5 * u * u) + 0.2 * exp(-2 * (u - 4.5) * (u - 4.5))) *
               exp(-0.5 * (v * v + w * w)) * (1 + alpha * (cos(k * x) + cos(k * y) + cos(k * z)));
    }
    /* config.hpp:233-234 */
    const double c = 0.06349363593424096978576330493464;
    return c * (1. + alpha * cos(k * x) + alpha * cos(k * y) + alpha * cos(k * z)) *
           exp(-(u * u + v * v + w * w) / 2);
}

/* ------------------------------------------------------------------ rho.hpp backtrace */

static size_t stride1(int order, const orc_conf1d *cf) { return cf->Nx + (size_t)order - 1; }
static size_t stride2(int order, const orc_conf2d *cf)
{
    return (cf->Nx + (size_t)order - 1) * (cf->Ny + (size_t)order - 1);
}
static size_t stride3(int order, const orc_conf3d *cf)
{
    return (cf->Nx + (size_t)order - 1) * (cf->Ny + (size_t)order - 1) * (cf->Nz + (size_t)order - 1);
}

/* rho.hpp:31-61 */
double orc_ftilda_1d(int order, size_t n, double x, double u, const double *coeffs, const orc_conf1d *cf, const orc_f0 *f)
{
    if (n == 0) return orc_f0_1d(f, x, u);
    const size_t stride_t = stride1(order, cf);
    double Ex;
    while (--n) {
        x = x - cf->dt * u;
        Ex = -orc_field_1d(order, 1, x, coeffs + n * stride_t, cf);
        u = u + cf->dt * Ex;
    }
    x -= cf->dt * u;
    Ex = -orc_field_1d(order, 1, x, coeffs, cf);
    u += 0.5 * cf->dt * Ex;
    return orc_f0_1d(f, x, u);
}

/* rho.hpp:63-96 */
double orc_f_1d(int order, size_t n, double x, double u, const double *coeffs, const orc_conf1d *cf, const orc_f0 *f)
{
    if (n == 0) return orc_f0_1d(f, x, u);
    const size_t stride_t = stride1(order, cf);
    double Ex = -orc_field_1d(order, 1, x, coeffs + n * stride_t, cf);
    u += 0.5 * cf->dt * Ex;
    while (--n) {
        x -= cf->dt * u;
        Ex = -orc_field_1d(order, 1, x, coeffs + n * stride_t, cf);
        u += cf->dt * Ex;
    }
    x -= cf->dt * u;
    Ex = -orc_field_1d(order, 1, x, coeffs, cf);
    u += 0.5 * cf->dt * Ex;
    return orc_f0_1d(f, x, u);
}

/* rho.hpp:98-131: the flow map itself (dim1 only in the reference): foot (x, u) of the characteristic through (x, u) at t_n.
 * Reference quirks kept: nothing is traced for n <= 1; the final x is wrapped with Lx*floor(x*Lx_inv) WITHOUT subtracting x_min. */
void orc_phase_flow_1d(int order, size_t n, double *px, double *pu, const double *coeffs, const orc_conf1d *cf)
{
    double x = *px, u = *pu;
    if (n > 1) {
        const size_t stride_t = stride1(order, cf);
        double Ex = -orc_field_1d(order, 1, x, coeffs + n * stride_t, cf);
        u += 0.5 * cf->dt * Ex;
        while (--n) {
            x -= cf->dt * u;
            Ex = -orc_field_1d(order, 1, x, coeffs + n * stride_t, cf);
            u += cf->dt * Ex;
        }
        x -= cf->dt * u;
        Ex = -orc_field_1d(order, 1, x, coeffs, cf);
        u += 0.5 * cf->dt * Ex;
    }
    x = x - cf->Lx * floor(x * cf->Lx_inv);
    *px = x;
    *pu = u;
}

/* rho.hpp:191-232 */
double orc_ftilda_2d(int order, size_t n, double x, double y, double u, double v, const double *coeffs,
                     const orc_conf2d *cf, const orc_f0 *f)
{
    if (n == 0) return orc_f0_2d(f, x, y, u, v);
    const size_t stride_t = stride2(order, cf);
    double Ex, Ey;
    const double *c;
    while (--n) {
        x -= cf->dt * u;
        y -= cf->dt * v;
        c = coeffs + n * stride_t;
        Ex = -orc_field_2d(order, 1, 0, x, y, c, cf);
        Ey = -orc_field_2d(order, 0, 1, x, y, c, cf);
        u += cf->dt * Ex;
        v += cf->dt * Ey;
    }
    x -= cf->dt * u;
    y -= cf->dt * v;
    Ex = -orc_field_2d(order, 1, 0, x, y, coeffs, cf);
    Ey = -orc_field_2d(order, 0, 1, x, y, coeffs, cf);
    u += 0.5 * cf->dt * Ex;
    v += 0.5 * cf->dt * Ey;
    return orc_f0_2d(f, x, y, u, v);
}

/* rho.hpp:234-281 */
double orc_f_2d(int order, size_t n, double x, double y, double u, double v, const double *coeffs,
                const orc_conf2d *cf, const orc_f0 *f)
{
    if (n == 0) return orc_f0_2d(f, x, y, u, v);
    const size_t stride_t = stride2(order, cf);
    const double *c = coeffs + n * stride_t;
    double Ex = -orc_field_2d(order, 1, 0, x, y, c, cf);
    double Ey = -orc_field_2d(order, 0, 1, x, y, c, cf);
    u += 0.5 * cf->dt * Ex;
    v += 0.5 * cf->dt * Ey;
    while (--n) {
        x -= cf->dt * u;
        y -= cf->dt * v;
        c = coeffs + n * stride_t;
        Ex = -orc_field_2d(order, 1, 0, x, y, c, cf);
        Ey = -orc_field_2d(order, 0, 1, x, y, c, cf);
        u += cf->dt * Ex;
        v += cf->dt * Ey;
    }
    x -= cf->dt * u;
    y -= cf->dt * v;
    Ex = -orc_field_2d(order, 1, 0, x, y, coeffs, cf);
    Ey = -orc_field_2d(order, 0, 1, x, y, coeffs, cf);
    u += 0.5 * cf->dt * Ex;
    v += 0.5 * cf->dt * Ey;
    return orc_f0_2d(f, x, y, u, v);
}

/* rho.hpp:318-367 */
double orc_ftilda_3d(int order, size_t n, double x, double y, double z, double u, double v, double w,
                     const double *coeffs, const orc_conf3d *cf, const orc_f0 *f)
{
    if (n == 0) return orc_f0_3d(f, x, y, z, u, v, w);
    const size_t stride_t = stride3(order, cf);
    double Ex, Ey, Ez;
    const double *c;
    while (--n) {
        x -= cf->dt * u;
        y -= cf->dt * v;
        z -= cf->dt * w;
        c = coeffs + n * stride_t;
        Ex = -orc_field_3d(order, 1, 0, 0, x, y, z, c, cf);
        Ey = -orc_field_3d(order, 0, 1, 0, x, y, z, c, cf);
        Ez = -orc_field_3d(order, 0, 0, 1, x, y, z, c, cf);
        u += cf->dt * Ex;
        v += cf->dt * Ey;
        w += cf->dt * Ez;
    }
    x -= cf->dt * u;
    y -= cf->dt * v;
    z -= cf->dt * w;
    Ex = -orc_field_3d(order, 1, 0, 0, x, y, z, coeffs, cf);
    Ey = -orc_field_3d(order, 0, 1, 0, x, y, z, coeffs, cf);
    Ez = -orc_field_3d(order, 0, 0, 1, x, y, z, coeffs, cf);
    u += 0.5 * cf->dt * Ex;
    v += 0.5 * cf->dt * Ey;
    w += 0.5 * cf->dt * Ez;
    return orc_f0_3d(f, x, y, z, u, v, w);
}

/* rho.hpp:369-426 */
double orc_f_3d(int order, size_t n, double x, double y, double z, double u, double v, double w,
                const double *coeffs, const orc_conf3d *cf, const orc_f0 *f)
{
    if (n == 0) return orc_f0_3d(f, x, y, z, u, v, w);
    const size_t stride_t = stride3(order, cf);
    const double *c = coeffs + n * stride_t;
    double Ex = -orc_field_3d(order, 1, 0, 0, x, y, z, c, cf);
    double Ey = -orc_field_3d(order, 0, 1, 0, x, y, z, c, cf);
    double Ez = -orc_field_3d(order, 0, 0, 1, x, y, z, c, cf);
    u += 0.5 * cf->dt * Ex;
    v += 0.5 * cf->dt * Ey;
    w += 0.5 * cf->dt * Ez;
    while (--n) {
        x -= cf->dt * u;
        y -= cf->dt * v;
        z -= cf->dt * w;
        c = coeffs + n * stride_t;
        Ex = -orc_field_3d(order, 1, 0, 0, x, y, z, c, cf);
        Ey = -orc_field_3d(order, 0, 1, 0, x, y, z, c, cf);
        Ez = -orc_field_3d(order, 0, 0, 1, x, y, z, c, cf);
        u += cf->dt * Ex;
        v += cf->dt * Ey;
        w += cf->dt * Ez;
    }
    x -= cf->dt * u;
    y -= cf->dt * v;
    z -= cf->dt * w;
    Ex = -orc_field_3d(order, 1, 0, 0, x, y, z, coeffs, cf);
    Ey = -orc_field_3d(order, 0, 1, 0, x, y, z, coeffs, cf);
    Ez = -orc_field_3d(order, 0, 0, 1, x, y, z, coeffs, cf);
    u += 0.5 * cf->dt * Ex;
    v += 0.5 * cf->dt * Ey;
    w += 0.5 * cf->dt * Ez;
    return orc_f0_3d(f, x, y, z, u, v, w);
}

/* ------------------------------------------------------------------ rho.hpp quadrature */

/* rho.hpp:133-146 */
double orc_rho_1d(int order, size_t n, size_t i, const double *coeffs, const orc_conf1d *cf, const orc_f0 *f)
{
    const double x = cf->x_min + i * cf->dx;
    const double du = (cf->u_max - cf->u_min) / cf->Nu;
    const double u_min = cf->u_min + 0.5 * du;
    double rho = 0;
    for (size_t ii = 0; ii < cf->Nu; ++ii) rho += orc_ftilda_1d(order, n, x, u_min + ii * du, coeffs, cf, f);
    rho = 1 - du * rho;
    return rho;
}

/* rho.hpp:283-310 */
double orc_rho_2d(int order, size_t n, size_t l, const double *coeffs, const orc_conf2d *cf, const orc_f0 *f)
{
    const size_t i = l % cf->Nx, j = l / cf->Nx;
    const double x = cf->x_min + i * cf->dx;
    const double y = cf->y_min + j * cf->dy;
    const double du = (cf->u_max - cf->u_min) / cf->Nu;
    const double dv = (cf->v_max - cf->v_min) / cf->Nv;
    const double u_min = cf->u_min + 0.5 * du;
    const double v_min = cf->v_min + 0.5 * dv;
    double rho = 0;
    for (size_t jj = 0; jj < cf->Nv; ++jj)
        for (size_t ii = 0; ii < cf->Nu; ++ii) {
            double u = u_min + ii * du;
            double v = v_min + jj * dv;
            rho += orc_ftilda_2d(order, n, x, y, u, v, coeffs, cf, f);
        }
    rho = 1 - du * dv * rho;
    return rho;
}

/* rho.hpp:428-462 */
double orc_rho_3d(int order, size_t n, size_t l, const double *coeffs, const orc_conf3d *cf, const orc_f0 *f)
{
    const size_t k = l / (cf->Nx * cf->Ny);
    const size_t tmp = l % (cf->Nx * cf->Ny);
    const size_t j = tmp / cf->Nx, i = tmp % cf->Nx;
    const double x = cf->x_min + i * cf->dx;
    const double y = cf->y_min + j * cf->dy;
    const double z = cf->z_min + k * cf->dz;
    const double du = (cf->u_max - cf->u_min) / cf->Nu;
    const double dv = (cf->v_max - cf->v_min) / cf->Nv;
    const double dw = (cf->w_max - cf->w_min) / cf->Nw;
    const double u_min = cf->u_min + 0.5 * du;
    const double v_min = cf->v_min + 0.5 * dv;
    const double w_min = cf->w_min + 0.5 * dw;
    double rho = 0;
    for (size_t kk = 0; kk < cf->Nw; ++kk)
        for (size_t jj = 0; jj < cf->Nv; ++jj)
            for (size_t ii = 0; ii < cf->Nu; ++ii) {
                double u = u_min + ii * du;
                double v = v_min + jj * dv;
                double w = w_min + kk * dw;
                rho += orc_ftilda_3d(order, n, x, y, z, u, v, w, coeffs, cf, f);
            }
    rho = 1 - du * dv * dw * rho;
    return rho;
}

/* bin/test_nufi_cpu_1d.cpp:65-70 */
void orc_rho_sweep_1d(int order, size_t n, const double *coeffs, const orc_conf1d *cf, const orc_f0 *f,
                      size_t l_begin, size_t l_end, double *rho)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t l = l_begin; l < l_end; ++l) rho[l] = orc_rho_1d(order, n, l, coeffs, cf, f);
}
/* bin/test_nufi_cpu_2d.cpp:68-72 */
void orc_rho_sweep_2d(int order, size_t n, const double *coeffs, const orc_conf2d *cf, const orc_f0 *f,
                      size_t l_begin, size_t l_end, double *rho)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t l = l_begin; l < l_end; ++l) rho[l] = orc_rho_2d(order, n, l, coeffs, cf, f);
}
/* bin/test_nufi_cpu_3d.cpp:68-72 */
void orc_rho_sweep_3d(int order, size_t n, const double *coeffs, const orc_conf3d *cf, const orc_f0 *f,
                      size_t l_begin, size_t l_end, double *rho)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t l = l_begin; l < l_end; ++l) rho[l] = orc_rho_3d(order, n, l, coeffs, cf, f);
}

/* ---- extended-precision sums: the same eval_ftilda values, added up in x87 long double (64-bit mantissa) instead of the
 *      reference's running double (rho.hpp:299-306).  NOT the reference's arithmetic: a yardstick that tells whose rounding a
 *      difference between two FP64 implementations is -- the reference's sequential sum of Nu*Nv*Nw terms carries up to
 *      ~sqrt(Nvel) eps of relative error in dV*sum f, i.e. that divided by the perturbation amplitude in rho. ---- */
static double rho_ext_node(int dim, int order, size_t n, size_t l, const double *coeffs, const void *cfv, const orc_f0 *f)
{
    long double sum = 0.0L;
    if (dim == 1) {
        const orc_conf1d *cf = (const orc_conf1d *)cfv;
        const double x = cf->x_min + l * cf->dx;
        const double du = (cf->u_max - cf->u_min) / cf->Nu, u_min = cf->u_min + 0.5 * du;
        for (size_t ii = 0; ii < cf->Nu; ++ii) sum += (long double)orc_ftilda_1d(order, n, x, u_min + ii * du, coeffs, cf, f);
        return (double)(1.0L - (long double)du * sum);
    }
    if (dim == 2) {
        const orc_conf2d *cf = (const orc_conf2d *)cfv;
        const size_t i = l % cf->Nx, j = l / cf->Nx;
        const double x = cf->x_min + i * cf->dx, y = cf->y_min + j * cf->dy;
        const double du = (cf->u_max - cf->u_min) / cf->Nu, dv = (cf->v_max - cf->v_min) / cf->Nv;
        const double u_min = cf->u_min + 0.5 * du, v_min = cf->v_min + 0.5 * dv;
        for (size_t jj = 0; jj < cf->Nv; ++jj)
            for (size_t ii = 0; ii < cf->Nu; ++ii)
                sum += (long double)orc_ftilda_2d(order, n, x, y, u_min + ii * du, v_min + jj * dv, coeffs, cf, f);
        return (double)(1.0L - (long double)(du * dv) * sum);
    }
    const orc_conf3d *cf = (const orc_conf3d *)cfv;
    const size_t k = l / (cf->Nx * cf->Ny), tmp = l % (cf->Nx * cf->Ny), j = tmp / cf->Nx, i = tmp % cf->Nx;
    const double x = cf->x_min + i * cf->dx, y = cf->y_min + j * cf->dy, z = cf->z_min + k * cf->dz;
    const double du = (cf->u_max - cf->u_min) / cf->Nu, dv = (cf->v_max - cf->v_min) / cf->Nv, dw = (cf->w_max - cf->w_min) / cf->Nw;
    const double u_min = cf->u_min + 0.5 * du, v_min = cf->v_min + 0.5 * dv, w_min = cf->w_min + 0.5 * dw;
    for (size_t kk = 0; kk < cf->Nw; ++kk)
        for (size_t jj = 0; jj < cf->Nv; ++jj)
            for (size_t ii = 0; ii < cf->Nu; ++ii)
                sum += (long double)orc_ftilda_3d(order, n, x, y, z, u_min + ii * du, v_min + jj * dv, w_min + kk * dw, coeffs, cf, f);
    return (double)(1.0L - (long double)(du * dv * dw) * sum);
}

/* dim = 1, 2, 3; cf points at the matching orc_conf{1,2,3}d */
void orc_rho_sweep_extended(int dim, int order, size_t n, const double *coeffs, const void *cf, const orc_f0 *f, size_t l_begin,
                            size_t l_end, double *rho)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t l = l_begin; l < l_end; ++l) rho[l] = rho_ext_node(dim, order, n, l, coeffs, cf, f);
}

/* ---- flat-q partial sums, the accumulate convention of the reference GPU path
 *      (cuda_kernel.cu:31-51, 210-237, 393-426; weights :46, :233, :422).  Sequential in q. ---- */

void orc_rho_partial_1d(int order, size_t n, const double *coeffs, const orc_conf1d *cf, const orc_f0 *f,
                        size_t q_begin, size_t q_end, double *rho)
{
    const double du = (cf->u_max - cf->u_min) / cf->Nu;
    const double u_min = cf->u_min + 0.5 * du;
    for (size_t q = q_begin; q < q_end; ++q) {
        const size_t ix = q / cf->Nu, iu = q % cf->Nu;
        const double x = cf->x_min + ix * cf->dx;
        const double v = orc_ftilda_1d(order, n, x, u_min + iu * du, coeffs, cf, f);
        rho[ix] += -cf->du * v;
    }
}

void orc_rho_partial_2d(int order, size_t n, const double *coeffs, const orc_conf2d *cf, const orc_f0 *f,
                        size_t q_begin, size_t q_end, double *rho)
{
    const double du = (cf->u_max - cf->u_min) / cf->Nu;
    const double dv = (cf->v_max - cf->v_min) / cf->Nv;
    const double u_min = cf->u_min + 0.5 * du, v_min = cf->v_min + 0.5 * dv;
    const double weight = cf->du * cf->dv;
    for (size_t q = q_begin; q < q_end; ++q) {
        size_t tmp = q;
        const size_t iy = tmp / (cf->Nx * cf->Nv * cf->Nu);
        tmp = tmp % (cf->Nx * cf->Nv * cf->Nu);
        const size_t ix = tmp / (cf->Nv * cf->Nu);
        tmp = tmp % (cf->Nv * cf->Nu);
        const size_t iv = tmp / cf->Nu, iu = tmp % cf->Nu;
        const double x = cf->x_min + ix * cf->dx, y = cf->y_min + iy * cf->dy;
        const double val = orc_ftilda_2d(order, n, x, y, u_min + iu * du, v_min + iv * dv, coeffs, cf, f);
        rho[iy * cf->Nx + ix] += -weight * val;
    }
}

void orc_rho_partial_3d(int order, size_t n, const double *coeffs, const orc_conf3d *cf, const orc_f0 *f,
                        size_t q_begin, size_t q_end, double *rho)
{
    const double du = (cf->u_max - cf->u_min) / cf->Nu;
    const double dv = (cf->v_max - cf->v_min) / cf->Nv;
    const double dw = (cf->w_max - cf->w_min) / cf->Nw;
    const double u_min = cf->u_min + 0.5 * du, v_min = cf->v_min + 0.5 * dv, w_min = cf->w_min + 0.5 * dw;
    const double weight = cf->du * cf->dv * cf->dw;
    const size_t Nvel = cf->Nu * cf->Nv * cf->Nw;
    for (size_t q = q_begin; q < q_end; ++q) {
        const size_t l = q / Nvel;
        size_t tmp = q % Nvel;
        const size_t iz = l / (cf->Nx * cf->Ny), iy = (l % (cf->Nx * cf->Ny)) / cf->Nx, ix = l % cf->Nx;
        const size_t iw = tmp / (cf->Nv * cf->Nu);
        tmp = tmp % (cf->Nv * cf->Nu);
        const size_t iv = tmp / cf->Nu, iu = tmp % cf->Nu;
        const double x = cf->x_min + ix * cf->dx, y = cf->y_min + iy * cf->dy, z = cf->z_min + iz * cf->dz;
        const double val = orc_ftilda_3d(order, n, x, y, z, u_min + iu * du, v_min + iv * dv, w_min + iw * dw,
                                         coeffs, cf, f);
        rho[l] += -weight * val;
    }
}

/* ---- metrics (cuda_kernel.cu:53-79, 239-271, 428-466), sequential in q; velocity nodes in the GPU
 *      rounding order (u_min + iu*du + du/2), weights exactly as the reference writes them
 *      (the 3d weight omits dx*dy*dz, cuda_kernel.cu:457). ---- */

static void metrics_add(double *m, double weight, double f, double vsq)
{
    m[0] += weight * f;
    m[1] += weight * f * f;
    m[2] += weight * vsq * f / 2;
    m[3] += (f > 0) ? -weight * f * log(f) : 0;
}

void orc_metrics_1d(int order, size_t n, const double *coeffs, const orc_conf1d *cf, const orc_f0 *f,
                    size_t q_begin, size_t q_end, double *m)
{
    const double weight = cf->du * cf->dx;
    for (size_t q = q_begin; q < q_end; ++q) {
        const size_t ix = q / cf->Nu, iu = q % cf->Nu;
        const double x = cf->x_min + ix * cf->dx;
        const double u = cf->u_min + iu * cf->du + cf->du / 2;
        const double val = orc_f_1d(order, n, x, u, coeffs, cf, f);
        m[0] += weight * val;
        m[1] += weight * val * val;
        m[2] += weight * (u * u * val / 2);
        m[3] += (val > 0) ? -weight * val * log(val) : 0;
    }
}

void orc_metrics_2d(int order, size_t n, const double *coeffs, const orc_conf2d *cf, const orc_f0 *f,
                    size_t q_begin, size_t q_end, double *m)
{
    const double weight = cf->dx * cf->dy * cf->du * cf->dv;
    for (size_t q = q_begin; q < q_end; ++q) {
        size_t tmp = q;
        const size_t iy = tmp / (cf->Nx * cf->Nv * cf->Nu);
        tmp = tmp % (cf->Nx * cf->Nv * cf->Nu);
        const size_t ix = tmp / (cf->Nv * cf->Nu);
        tmp = tmp % (cf->Nv * cf->Nu);
        const size_t iv = tmp / cf->Nu, iu = tmp % cf->Nu;
        const double x = cf->x_min + ix * cf->dx, y = cf->y_min + iy * cf->dy;
        const double u = cf->u_min + iu * cf->du + cf->du / 2;
        const double v = cf->v_min + iv * cf->dv + cf->dv / 2;
        const double val = orc_f_2d(order, n, x, y, u, v, coeffs, cf, f);
        metrics_add(m, weight, val, u * u + v * v);
    }
}

void orc_metrics_3d(int order, size_t n, const double *coeffs, const orc_conf3d *cf, const orc_f0 *f,
                    size_t q_begin, size_t q_end, double *m)
{
    const double weight = cf->du * cf->dv * cf->dw;
    const size_t Nvel = cf->Nu * cf->Nv * cf->Nw;
    for (size_t q = q_begin; q < q_end; ++q) {
        const size_t l = q / Nvel;
        size_t tmp = q % Nvel;
        const size_t iz = l / (cf->Nx * cf->Ny), iy = (l % (cf->Nx * cf->Ny)) / cf->Nx, ix = l % cf->Nx;
        const size_t iw = tmp / (cf->Nv * cf->Nu);
        tmp = tmp % (cf->Nv * cf->Nu);
        const size_t iv = tmp / cf->Nu, iu = tmp % cf->Nu;
        const double x = cf->x_min + ix * cf->dx, y = cf->y_min + iy * cf->dy, z = cf->z_min + iz * cf->dz;
        const double u = cf->u_min + iu * cf->du + cf->du / 2;
        const double v = cf->v_min + iv * cf->dv + cf->dv / 2;
        const double w = cf->w_min + iw * cf->dw + cf->dw / 2;
        const double val = orc_f_3d(order, n, x, y, z, u, v, w, coeffs, cf, f);
        metrics_add(m, weight, val, u * u + v * v + w * w);
    }
}

/* ------------------------------------------------------------------ poisson.cpp
 * FFTW_DHT restated: H[k] = sum_j x[j] (cos(2 pi j k/N) + sin(2 pi j k/N)), unnormalised; a
 * multi-dimensional r2r DHT plan is the separable product of 1-d DHTs along each axis. */

static void dht_axis(double *data, size_t n, size_t stride, size_t count_outer, size_t outer_stride,
                     size_t count_inner, size_t inner_stride)
{
    double *cas = (double *)malloc(sizeof(double) * n);
    double *tmp = (double *)malloc(sizeof(double) * n);
    for (size_t m = 0; m < n; ++m) {
        const double a = 2.0 * M_PI * (double)m / (double)n;
        cas[m] = cos(a) + sin(a);
    }
    for (size_t o = 0; o < count_outer; ++o)
        for (size_t in = 0; in < count_inner; ++in) {
            double *line = data + o * outer_stride + in * inner_stride;
            for (size_t k = 0; k < n; ++k) {
                double s = 0;
                for (size_t j = 0; j < n; ++j) s += line[j * stride] * cas[(j * k) % n];
                tmp[k] = s;
            }
            for (size_t k = 0; k < n; ++k) line[k * stride] = tmp[k];
        }
    free(cas);
    free(tmp);
}

/* poisson.cpp:66-89 */
double orc_poisson_1d(const orc_conf1d *cf, double *data)
{
    dht_axis(data, cf->Nx, 1, 1, 0, 1, 0);
    const double fac_N = 1.0 / (double)cf->Nx;
    const double fac_x = 2 * M_PI * cf->Lx_inv;
    double energy = 0;
    for (size_t i = 1; i < cf->Nx; i++) {
        double ii = (2 * i < cf->Nx) ? i : cf->Nx - i;
        double fac = fac_N / (ii * ii * fac_x * fac_x);
        data[i] *= fac;
        double Ex = ii * fac_x * data[i];
        energy += Ex * Ex;
    }
    data[0] = 0;
    dht_axis(data, cf->Nx, 1, 1, 0, 1, 0);
    energy *= cf->Lx / 2;
    return energy;
}

/* poisson.cpp:190-219 */
double orc_poisson_2d(const orc_conf2d *cf, double *data)
{
    const size_t Nx = cf->Nx, Ny = cf->Ny;
    dht_axis(data, Nx, 1, Ny, Nx, 1, 0);
    dht_axis(data, Ny, Nx, 1, 0, Nx, 1);
    const double fac_N = 1.0 / (double)(Nx * Ny);
    const double fac_x = 2 * M_PI * cf->Lx_inv, fac_y = 2 * M_PI * cf->Ly_inv;
    double energy = 0;
    for (size_t j = 0; j < Ny; j++)
        for (size_t i = 0; i < Nx; i++) {
            if (j == 0 && i == 0) continue;
            double ii = (2 * i < Nx) ? i : Nx - i;
            double jj = (2 * j < Ny) ? j : Ny - j;
            double fac = fac_N / (ii * ii * fac_x * fac_x + jj * jj * fac_y * fac_y);
            data[j * Nx + i] *= fac;
            double Ex = ii * fac_x * data[j * Nx + i];
            double Ey = jj * fac_y * data[j * Nx + i];
            energy += Ex * Ex + Ey * Ey;
        }
    data[0] = 0;
    dht_axis(data, Nx, 1, Ny, Nx, 1, 0);
    dht_axis(data, Ny, Nx, 1, 0, Nx, 1);
    energy *= cf->Lx * cf->Ly / 2;
    return energy;
}

static void dht3(double *data, size_t Nx, size_t Ny, size_t Nz)
{
    dht_axis(data, Nx, 1, Ny * Nz, Nx, 1, 0);
    dht_axis(data, Ny, Nx, Nz, Nx * Ny, Nx, 1);
    dht_axis(data, Nz, Nx * Ny, 1, 0, Nx * Ny, 1);
}

/* poisson.cpp:328-362 */
double orc_poisson_3d(const orc_conf3d *cf, double *data)
{
    const size_t Nx = cf->Nx, Ny = cf->Ny, Nz = cf->Nz;
    dht3(data, Nx, Ny, Nz);
    const double fac_N = 1.0 / (double)(Nx * Ny * Nz);
    const double fac_x = 2 * M_PI * cf->Lx_inv, fac_y = 2 * M_PI * cf->Ly_inv, fac_z = 2 * M_PI * cf->Lz_inv;
    double energy = 0;
    for (size_t k = 0; k < Nz; k++)
        for (size_t j = 0; j < Ny; j++)
            for (size_t i = 0; i < Nx; i++) {
                if (i == 0 && j == 0 && k == 0) continue;
                double ii = (2 * i < Nx) ? i : Nx - i;
                double jj = (2 * j < Ny) ? j : Ny - j;
                double kk = (2 * k < Nz) ? k : Nz - k;
                double fac = fac_N / (ii * ii * fac_x * fac_x + jj * jj * fac_y * fac_y + kk * kk * fac_z * fac_z);
                double *d = data + k * Nx * Ny + j * Nx + i;
                *d *= fac;
                double Ex = ii * fac_x * *d, Ey = jj * fac_y * *d, Ez = kk * fac_z * *d;
                energy += Ex * Ex + Ey * Ey + Ez * Ez;
            }
    data[0] = 0;
    dht3(data, Nx, Ny, Nz);
    energy *= cf->Lx * cf->Ly * cf->Lz / 2;
    return energy;
}

/* ------------------------------------------------------------------ fields.hpp interpolate
 * The collocation system (fields.hpp:76-95, 200-241, 370-421): sum_ii N_ii(0) c[(i+ii) mod N] = values[i]
 * per dimension (tensor product).  The reference solves it with LSMR to eps; here it is solved exactly,
 * dimension by dimension, with a dense LU of the N x N circulant matrix (deviation stated in the header).
 * Odd orders on an even grid: the stencil N_ii(0) is symmetric about a half-integer, so its symbol vanishes at the
 * Nyquist mode s_i = (-1)^i and the system is singular.  LSMR started from zero (lsmr.tpp) converges to the minimum-norm
 * least-squares solution A^+ v; A being normal with the single null vector s, A^+ v = (A + s s^T / n)^{-1} (v - s (s.v)/n),
 * which is what is factored / solved here in that case. */

typedef struct { size_t n; double *lu; size_t *piv; int singular; } circ_lu;

static circ_lu circ_factor(int order, size_t n)
{
    circ_lu F;
    F.n = n;
    F.lu = (double *)calloc(n * n, sizeof(double));
    F.piv = (size_t *)malloc(sizeof(size_t) * n);
    double N0[ORC_MAX_ORDER];
    orc_bspline_basis(order, 0, 0.0, N0);
    for (size_t i = 0; i < n; ++i)
        for (int ii = 0; ii < order; ++ii) F.lu[i * n + (i + (size_t)ii) % n] += N0[ii];
    F.singular = (order % 2 == 1) && (n % 2 == 0);
    if (F.singular)
        for (size_t i = 0; i < n; ++i)
            for (size_t j = 0; j < n; ++j) F.lu[i * n + j] += (((i + j) & 1) ? -1.0 : 1.0) / (double)n;
    for (size_t k = 0; k < n; ++k) {
        size_t p = k;
        for (size_t r = k + 1; r < n; ++r)
            if (fabs(F.lu[r * n + k]) > fabs(F.lu[p * n + k])) p = r;
        F.piv[k] = p;
        if (p != k)
            for (size_t c = 0; c < n; ++c) {
                double t = F.lu[k * n + c];
                F.lu[k * n + c] = F.lu[p * n + c];
                F.lu[p * n + c] = t;
            }
        for (size_t r = k + 1; r < n; ++r) {
            double m = F.lu[r * n + k] / F.lu[k * n + k];
            F.lu[r * n + k] = m;
            if (m != 0.0)
                for (size_t c = k + 1; c < n; ++c) F.lu[r * n + c] -= m * F.lu[k * n + c];
        }
    }
    return F;
}

static void circ_solve(const circ_lu *F, double *x, size_t stride, double *work)
{
    const size_t n = F->n;
    for (size_t i = 0; i < n; ++i) work[i] = x[i * stride];
    if (F->singular) { /* remove the Nyquist component of the right-hand side */
        double a = 0;
        for (size_t i = 0; i < n; ++i) a += (i & 1) ? -work[i] : work[i];
        a /= (double)n;
        for (size_t i = 0; i < n; ++i) work[i] -= (i & 1) ? -a : a;
    }
    for (size_t k = 0; k < n; ++k) { /* whole rows were swapped while factoring: permute first */
        size_t p = F->piv[k];
        if (p != k) { double t = work[k]; work[k] = work[p]; work[p] = t; }
    }
    for (size_t k = 0; k < n; ++k)
        for (size_t r = k + 1; r < n; ++r) work[r] -= F->lu[r * n + k] * work[k];
    for (size_t k = n; k-- > 0;) {
        double s = work[k];
        for (size_t c = k + 1; c < n; ++c) s -= F->lu[k * n + c] * work[c];
        work[k] = s / F->lu[k * n + k];
    }
    for (size_t i = 0; i < n; ++i) x[i * stride] = work[i];
}

static void circ_free(circ_lu *F) { free(F->lu); free(F->piv); }

/* fields.hpp:63-142 */
void orc_interpolate_1d(int order, double *level, const double *values, const orc_conf1d *cf)
{
    const size_t Nx = cf->Nx;
    double *tmp = (double *)malloc(sizeof(double) * Nx), *work = (double *)malloc(sizeof(double) * Nx);
    memcpy(tmp, values, sizeof(double) * Nx);
    circ_lu F = circ_factor(order, Nx);
    circ_solve(&F, tmp, 1, work);
    circ_free(&F);
    for (size_t i = 0; i < Nx + (size_t)order - 1; ++i) level[i] = tmp[i % Nx]; /* :140-141 */
    free(tmp);
    free(work);
}

/* fields.hpp:186-300 */
void orc_interpolate_2d(int order, double *level, const double *values, const orc_conf2d *cf)
{
    const size_t Nx = cf->Nx, Ny = cf->Ny, nmax = Nx > Ny ? Nx : Ny;
    double *tmp = (double *)malloc(sizeof(double) * Nx * Ny), *work = (double *)malloc(sizeof(double) * nmax);
    memcpy(tmp, values, sizeof(double) * Nx * Ny);
    circ_lu Fx = circ_factor(order, Nx), Fy = circ_factor(order, Ny);
    for (size_t j = 0; j < Ny; ++j) circ_solve(&Fx, tmp + j * Nx, 1, work);
    for (size_t i = 0; i < Nx; ++i) circ_solve(&Fy, tmp + i, Nx, work);
    circ_free(&Fx);
    circ_free(&Fy);
    const size_t stride_y = Nx + (size_t)order - 1;
    for (size_t j = 0; j < Ny + (size_t)order - 1; ++j) /* :294-299 */
        for (size_t i = 0; i < Nx + (size_t)order - 1; ++i)
            level[j * stride_y + i] = tmp[(j % Ny) * Nx + (i % Nx)];
    free(tmp);
    free(work);
}

/* fields.hpp:352-490 */
void orc_interpolate_3d(int order, double *level, const double *values, const orc_conf3d *cf)
{
    const size_t Nx = cf->Nx, Ny = cf->Ny, Nz = cf->Nz;
    size_t nmax = Nx > Ny ? Nx : Ny;
    if (Nz > nmax) nmax = Nz;
    double *tmp = (double *)malloc(sizeof(double) * Nx * Ny * Nz), *work = (double *)malloc(sizeof(double) * nmax);
    memcpy(tmp, values, sizeof(double) * Nx * Ny * Nz);
    circ_lu Fx = circ_factor(order, Nx), Fy = circ_factor(order, Ny), Fz = circ_factor(order, Nz);
    for (size_t r = 0; r < Ny * Nz; ++r) circ_solve(&Fx, tmp + r * Nx, 1, work);
    for (size_t k = 0; k < Nz; ++k)
        for (size_t i = 0; i < Nx; ++i) circ_solve(&Fy, tmp + k * Nx * Ny + i, Nx, work);
    for (size_t r = 0; r < Nx * Ny; ++r) circ_solve(&Fz, tmp + r, Nx * Ny, work);
    circ_free(&Fx);
    circ_free(&Fy);
    circ_free(&Fz);
    const size_t stride_y = Nx + (size_t)order - 1;
    const size_t stride_z = (Ny + (size_t)order - 1) * stride_y;
    for (size_t k = 0; k < Nz + (size_t)order - 1; ++k) /* :482-489 */
        for (size_t j = 0; j < Ny + (size_t)order - 1; ++j)
            for (size_t i = 0; i < Nx + (size_t)order - 1; ++i)
                level[k * stride_z + j * stride_y + i] = tmp[(k % Nz) * Nx * Ny + (j % Ny) * Nx + (i % Nx)];
    free(tmp);
    free(work);
}

/* ------------------------------------------------------------------ the drivers' time loop */

/* bin/test_nufi_cpu_1d.cpp:60-76 */
void orc_run_1d(int order, const orc_conf1d *cf, const orc_f0 *f, size_t n_begin, size_t n_end,
                double *coeffs, double *energy, double *rho_out)
{
    const size_t stride_t = stride1(order, cf), N = cf->Nx;
    double *rho = (double *)malloc(sizeof(double) * N);
    for (size_t n = n_begin; n < n_end; ++n) {
        orc_rho_sweep_1d(order, n, coeffs, cf, f, 0, N, rho);
        if (rho_out && n + 1 == n_end) memcpy(rho_out, rho, sizeof(double) * N);
        double e = orc_poisson_1d(cf, rho);
        if (energy) energy[n] = e;
        orc_interpolate_1d(order, coeffs + n * stride_t, rho, cf);
    }
    free(rho);
}

/* bin/test_nufi_cpu_2d.cpp:63-76 */
void orc_run_2d(int order, const orc_conf2d *cf, const orc_f0 *f, size_t n_begin, size_t n_end,
                double *coeffs, double *energy, double *rho_out)
{
    const size_t stride_t = stride2(order, cf), N = cf->Nx * cf->Ny;
    double *rho = (double *)malloc(sizeof(double) * N);
    for (size_t n = n_begin; n < n_end; ++n) {
        orc_rho_sweep_2d(order, n, coeffs, cf, f, 0, N, rho);
        if (rho_out && n + 1 == n_end) memcpy(rho_out, rho, sizeof(double) * N);
        double e = orc_poisson_2d(cf, rho);
        if (energy) energy[n] = e;
        orc_interpolate_2d(order, coeffs + n * stride_t, rho, cf);
    }
    free(rho);
}

/* bin/test_nufi_cpu_3d.cpp:63-76 */
void orc_run_3d(int order, const orc_conf3d *cf, const orc_f0 *f, size_t n_begin, size_t n_end,
                double *coeffs, double *energy, double *rho_out)
{
    const size_t stride_t = stride3(order, cf), N = cf->Nx * cf->Ny * cf->Nz;
    double *rho = (double *)malloc(sizeof(double) * N);
    for (size_t n = n_begin; n < n_end; ++n) {
        orc_rho_sweep_3d(order, n, coeffs, cf, f, 0, N, rho);
        if (rho_out && n + 1 == n_end) memcpy(rho_out, rho, sizeof(double) * N);
        double e = orc_poisson_3d(cf, rho);
        if (energy) energy[n] = e;
        orc_interpolate_3d(order, coeffs + n * stride_t, rho, cf);
    }
    free(rho);
}
