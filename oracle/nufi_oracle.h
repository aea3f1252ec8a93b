/*
 * nufi_oracle.h -- CPU restatement of the NuFI hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the parity oracle for numericalflowiteration_b200.  It restates, in plain C,
 * the arithmetic of the reference (paulwilhelmvlasov/NumericalFlowIteration) for
 *   splines.hpp  -> orc_bspline_basis / orc_deboor            (nufi/splines.hpp:39-110)
 *   fields.hpp   -> orc_field_{1,2,3}d                        (nufi/fields.hpp:36-61,149-184,308-350)
 *   rho.hpp      -> orc_ftilda_*, orc_f_*, orc_rho_*          (nufi/rho.hpp:31-96,133-146,191-310,318-462)
 *   poisson.cpp  -> orc_poisson_*                             (nufi/poisson.cpp:66-89,190-219,328-362)
 *   fields.hpp   -> orc_interpolate_*                         (nufi/fields.hpp:63-142,186-300,352-490)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (libnufi_b200.so) never links or calls it.
 *
 * Pinning: oracle/_ref (the real reference headers compiled in place from /root/reference)
 * is compared against this restatement by tests/test_oracle_vs_ref.py and was used to
 * generate tests/golden/*.npz (tests/golden/make_golden.py).
 *
 * Deviations from the reference, all stated here:
 *   - poisson: FFTW3 (un-vendored third-party dependency, absent) is replaced by a direct
 *     O(N^2) separable discrete Hartley transform, H[k] = sum_j x[j] cas(2 pi j k / N),
 *     which is FFTW's documented FFTW_DHT definition (r2r kind, unnormalised, separable in
 *     several dimensions).
 *   - interpolate: LSMR (iterative, tolerance eps) is replaced by the exact solve of the
 *     same periodic collocation system (cyclic Thomas / dense LU per dimension);
 *     agreement with the reference's LSMR is checked through oracle/_ref to ~1e-13.
 *   - f0 is selectable at run time (kind + parameters) instead of by editing config.hpp;
 *     every expression is one of the (active or commented) lines of nufi/config.hpp.
 */
#ifndef NUFI_ORACLE_H
#define NUFI_ORACLE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Field order mirrors nufi::dim{1,2,3}::config_t<double> (nufi/config.hpp:33-53,91-114,167-193). */
typedef struct {
    size_t Nx, Nu, Nt;
    double dt;
    double x_min, x_max;
    double u_min, u_max;
    double dx, dx_inv, Lx, Lx_inv;
    double du;
} orc_conf1d;

typedef struct {
    size_t Nx, Ny, Nu, Nv, Nt;
    double dt;
    double x_min, x_max, y_min, y_max;
    double u_min, u_max, v_min, v_max;
    double dx, dx_inv, Lx, Lx_inv;
    double dy, dy_inv, Ly, Ly_inv;
    double du, dv;
} orc_conf2d;

typedef struct {
    size_t Nx, Ny, Nz, Nu, Nv, Nw, Nt;
    double dt;
    double x_min, x_max, y_min, y_max, z_min, z_max;
    double u_min, u_max, v_min, v_max, w_min, w_max;
    double dx, dx_inv, Lx, Lx_inv;
    double dy, dy_inv, Ly, Ly_inv;
    double dz, dz_inv, Lz, Lz_inv;
    double du, dv, dw;
} orc_conf3d;

/* Initial condition selector.  kinds per dimension (expressions from nufi/config.hpp):
 *  1d: 0 = Landau      c(1+a cos kx) exp(-u^2/2)              (:83)   p = {alpha, k}
 *      1 = two-stream  c(1+a cos kx) exp(-u^2/2) u^2          (:82)   p = {alpha, k}
 *  2d: 0 = Landau      1/(2pi) exp(-(u^2+v^2)/2)(1+a(cos kx+cos ky))  (:148-149) p = {alpha, k}
 *      1 = two-stream  (:151-158)                                      p = {alpha, k, v0}
 *  3d: 0 = Landau      c(1+a cos kx+a cos ky+a cos kz)exp(-|v|^2/2)   (:233-234) p = {alpha, k}
 *      1 = two-stream  (:237-242)                                      p = {alpha, k, v0}
 *      2 = bump-on-tail (:244-246)                                     p = {0.03, 0.3} fixed shape
 */
typedef struct { int kind; double p[4]; } orc_f0;

/* ---- splines.hpp ---- */
void   orc_bspline_basis(int order, int der, double x, double *out);
double orc_deboor(int order, int der, double x, const double *c, size_t stride);

/* ---- fields.hpp eval: value or first derivatives of the spline potential ---- */
double orc_field_1d(int order, int dx, double x, const double *level, const orc_conf1d *cf);
double orc_field_2d(int order, int dx, int dy, double x, double y, const double *level, const orc_conf2d *cf);
double orc_field_3d(int order, int dx, int dy, int dz, double x, double y, double z,
                    const double *level, const orc_conf3d *cf);

/* ---- rho.hpp ---- */
double orc_f0_1d(const orc_f0 *f, double x, double u);
double orc_f0_2d(const orc_f0 *f, double x, double y, double u, double v);
double orc_f0_3d(const orc_f0 *f, double x, double y, double z, double u, double v, double w);

double orc_ftilda_1d(int order, size_t n, double x, double u, const double *coeffs, const orc_conf1d *cf, const orc_f0 *f);
void orc_phase_flow_1d(int order, size_t n, double *x, double *u, const double *coeffs, const orc_conf1d *cf); /* rho.hpp:98-131 */
double orc_ftilda_2d(int order, size_t n, double x, double y, double u, double v, const double *coeffs, const orc_conf2d *cf, const orc_f0 *f);
double orc_ftilda_3d(int order, size_t n, double x, double y, double z, double u, double v, double w,
                     const double *coeffs, const orc_conf3d *cf, const orc_f0 *f);
double orc_f_1d(int order, size_t n, double x, double u, const double *coeffs, const orc_conf1d *cf, const orc_f0 *f);
double orc_f_2d(int order, size_t n, double x, double y, double u, double v, const double *coeffs, const orc_conf2d *cf, const orc_f0 *f);
double orc_f_3d(int order, size_t n, double x, double y, double z, double u, double v, double w,
                const double *coeffs, const orc_conf3d *cf, const orc_f0 *f);

double orc_rho_1d(int order, size_t n, size_t i, const double *coeffs, const orc_conf1d *cf, const orc_f0 *f);
double orc_rho_2d(int order, size_t n, size_t l, const double *coeffs, const orc_conf2d *cf, const orc_f0 *f);
double orc_rho_3d(int order, size_t n, size_t l, const double *coeffs, const orc_conf3d *cf, const orc_f0 *f);

/* The drivers' "#pragma omp parallel for" sweep over spatial nodes l in [l_begin,l_end)
 * (bin/test_nufi_cpu_1d.cpp:65-70, _2d.cpp:68-72, _3d.cpp:68-72).  rho has Nx[*Ny[*Nz]] entries;
 * only [l_begin,l_end) is written. */
void orc_rho_sweep_1d(int order, size_t n, const double *coeffs, const orc_conf1d *cf, const orc_f0 *f, size_t l_begin, size_t l_end, double *rho);
void orc_rho_sweep_2d(int order, size_t n, const double *coeffs, const orc_conf2d *cf, const orc_f0 *f, size_t l_begin, size_t l_end, double *rho);
void orc_rho_sweep_3d(int order, size_t n, const double *coeffs, const orc_conf3d *cf, const orc_f0 *f, size_t l_begin, size_t l_end, double *rho);

/* GPU-flavoured flat-q partial sums (nufi/cuda_kernel.cu:31-51,210-237,393-426): rho[l] += -dV f for q in
 * [q_begin,q_end); velocity nodes use the CPU rounding order (u_min + 0.5 du + i du). */
void orc_rho_partial_1d(int order, size_t n, const double *coeffs, const orc_conf1d *cf, const orc_f0 *f, size_t q_begin, size_t q_end, double *rho);
void orc_rho_partial_2d(int order, size_t n, const double *coeffs, const orc_conf2d *cf, const orc_f0 *f, size_t q_begin, size_t q_end, double *rho);
void orc_rho_partial_3d(int order, size_t n, const double *coeffs, const orc_conf3d *cf, const orc_f0 *f, size_t q_begin, size_t q_end, double *rho);

/* Metrics (nufi/cuda_kernel.cu:53-79,239-271,428-466): m[0..3] += int f, int f^2, kinetic, entropy, summed
 * sequentially in flat-q order. */
void orc_metrics_1d(int order, size_t n, const double *coeffs, const orc_conf1d *cf, const orc_f0 *f, size_t q_begin, size_t q_end, double *m);
void orc_metrics_2d(int order, size_t n, const double *coeffs, const orc_conf2d *cf, const orc_f0 *f, size_t q_begin, size_t q_end, double *m);
void orc_metrics_3d(int order, size_t n, const double *coeffs, const orc_conf3d *cf, const orc_f0 *f, size_t q_begin, size_t q_end, double *m);

/* ---- poisson.cpp: in-place -Laplace(phi) = rho; returns electric energy ---- */
double orc_poisson_1d(const orc_conf1d *cf, double *data);
double orc_poisson_2d(const orc_conf2d *cf, double *data);
double orc_poisson_3d(const orc_conf3d *cf, double *data);

/* the same sweep with the velocity sum carried in long double: a precision yardstick, not the reference's arithmetic */
void orc_rho_sweep_extended(int dim, int order, size_t n, const double *coeffs, const void *cf, const orc_f0 *f, size_t l_begin,
                            size_t l_end, double *rho);

/* ---- fields.hpp interpolate: values at nodes -> level coefficients with (order-1) periodic halo ---- */
void orc_interpolate_1d(int order, double *level, const double *values, const orc_conf1d *cf);
void orc_interpolate_2d(int order, double *level, const double *values, const orc_conf2d *cf);
void orc_interpolate_3d(int order, double *level, const double *values, const orc_conf3d *cf);

/* ---- the CPU drivers' time loop: steps n = n_begin .. n_end-1, each rho sweep -> solve -> interpolate
 *      (bin/test_nufi_cpu_{1,2,3}d.cpp).  coeffs holds (>= n_end) levels; energy[n], if non-NULL,
 *      receives the return value of solve; rho_out, if non-NULL, receives rho of the LAST step
 *      before the solve. */
void orc_run_1d(int order, const orc_conf1d *cf, const orc_f0 *f, size_t n_begin, size_t n_end, double *coeffs, double *energy, double *rho_out);
void orc_run_2d(int order, const orc_conf2d *cf, const orc_f0 *f, size_t n_begin, size_t n_end, double *coeffs, double *energy, double *rho_out);
void orc_run_3d(int order, const orc_conf3d *cf, const orc_f0 *f, size_t n_begin, size_t n_end, double *coeffs, double *energy, double *rho_out);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
