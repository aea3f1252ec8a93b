/*
 * nufi_b200.h -- C ABI of libnufi_b200.so, the B200 (sm_100a) implementation of NuFI's hot path:
 * per time step n, trace every quadrature point (x,v) back through the stored history of potential
 * spline coefficients (levels n-1..0), evaluate f0 at the foot, reduce into rho; then the field tail
 * (periodic Poisson solve + spline interpolation) producing level n -- all on the device.
 *
 * Boundary being replaced (reference file:line, paulwilhelmvlasov/NumericalFlowIteration):
 *   nufi::dim{1,2,3}::cuda_scheduler<real,order>   nufi/cuda_scheduler.hpp:33-164, 173-279, 286-392
 *   nufi::dim{1,2,3}::cuda_kernel<real,order>      nufi/cuda_kernel.hpp:33-56, 77-98, 119-140;
 *                                                   nufi/cuda_kernel.cu:31-51, 112-189, 210-237, 289-371,
 *                                                   393-426, 468-573
 *   and, for the CPU-shaped callers, the loop body of bin/test_nufi_cpu_{1,2,3}d.cpp:
 *   dimN::eval_rho (nufi/rho.hpp:133-146, 283-310, 428-462), dimN::poisson<double>::solve
 *   (nufi/poisson.cpp:66-89, 190-219, 328-362), dimN::interpolate (nufi/fields.hpp:63-142, 186-300, 352-490).
 *
 * Conventions kept from the reference:
 *   - nufi_b200_config{1,2,3}d has the memory layout of nufi::dim{1,2,3}::config_t<double>
 *     (nufi/config.hpp:33-53, 91-114, 167-193); a C++ caller may pass &conf reinterpret_cast.
 *   - history level m lives at coeffs + m*stride_t, stride_t = prod_d (N_d + order - 1), x fastest
 *     (nufi/rho.hpp:326-329).  Step n reads levels n-1..1 with full kicks, level 0 with a half kick.
 *   - flat quadrature index q: 1d q = ix*Nu + iu; 2d q = ((iy*Nx+ix)*Nv+iv)*Nu+iu;
 *     3d q = ((((iz*Ny+iy)*Nx+ix)*Nw+iw)*Nv+iv)*Nu+iu (nufi/cuda_kernel.cu:40-41, 219-225, 402-412).
 *   - velocity nodes are midpoints, computed as (u_min + 0.5*du) + i*du with du=(u_max-u_min)/Nu (the CPU
 *     form, nufi/rho.hpp:137-142) -- the parity target is the reference's CPU eval_rho.
 *   - compute_rho/download_rho use the reference GPU path's convention: partial rho[l] = -dV * sum f (no
 *     leading 1), download ACCUMULATES into the caller's array (nufi/cuda_kernel.cu:135-145).
 *     eval_rho_all / the fused step use the CPU convention rho = 1 - dV * sum f (nufi/rho.hpp:145).
 *
 * Every function returns NUFI_B200_OK or an error code; nufi_b200_last_error() gives the message
 * (the reference throws: cuda::exception -> ERR_CUDA, std::range_error -> ERR_RANGE, std::bad_alloc ->
 * ERR_ALLOC; nufi/cuda_runtime.hpp:93-98, nufi/cuda_kernel.cu:115-116, 306-307).
 * Calls on one handle are not re-entrant.  One handle drives one GPU; work is sharded over GPUs by
 * giving each handle its own [q_begin,q_end) (the reference's partition, nufi/cuda_scheduler.hpp:88-111)
 * and summing the partial rho vectors (NCCL all-reduce in the host layer).
 * There is no CPU fallback: without a CUDA device every entry point fails with ERR_CUDA.
 */
#ifndef NUFI_B200_H
#define NUFI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NUFI_B200_OK 0
#define NUFI_B200_ERR_RANGE 1 /* time step / level / index out of range           (std::range_error) */
#define NUFI_B200_ERR_CUDA 2  /* CUDA / cuFFT runtime failure, or no device        (cuda::exception)  */
#define NUFI_B200_ERR_ALLOC 3 /* host or device allocation failed                  (std::bad_alloc)   */
#define NUFI_B200_ERR_ARG 4   /* invalid or unsupported argument                   (std::invalid_argument) */

typedef struct nufi_b200_handle nufi_b200_handle;

/* nufi::dim1::config_t<double>, nufi/config.hpp:33-53 */
typedef struct {
    size_t Nx, Nu, Nt;
    double dt;
    double x_min, x_max;
    double u_min, u_max;
    double dx, dx_inv, Lx, Lx_inv;
    double du;
} nufi_b200_config1d;

/* nufi::dim2::config_t<double>, nufi/config.hpp:91-114 */
typedef struct {
    size_t Nx, Ny, Nu, Nv, Nt;
    double dt;
    double x_min, x_max, y_min, y_max;
    double u_min, u_max, v_min, v_max;
    double dx, dx_inv, Lx, Lx_inv;
    double dy, dy_inv, Ly, Ly_inv;
    double du, dv;
} nufi_b200_config2d;

/* nufi::dim3::config_t<double>, nufi/config.hpp:167-193 */
typedef struct {
    size_t Nx, Ny, Nz, Nu, Nv, Nw, Nt;
    double dt;
    double x_min, x_max, y_min, y_max, z_min, z_max;
    double u_min, u_max, v_min, v_max, w_min, w_max;
    double dx, dx_inv, Lx, Lx_inv;
    double dy, dy_inv, Ly, Ly_inv;
    double dz, dz_inv, Lz, Lz_inv;
    double du, dv, dw;
} nufi_b200_config3d;

/* Initial condition f0 (static member of config_t in the reference, chosen by editing nufi/config.hpp).
 *  1d: kind 0 Landau (:83), 1 two-stream (:82);                           p = {alpha, k}
 *  2d: kind 0 Landau (:148-149), 1 two-stream (:151-158);                 p = {alpha, k, v0}
 *  3d: kind 0 Landau (:233-234), 1 two-stream (:237-242), 2 bump-on-tail (:244-246); p = {alpha, k, v0}
 * f0 must be periodic in x with the box (it is for every configuration of the reference). */
typedef struct {
    int kind;
    double p[4];
} nufi_b200_f0;

/* ---- lifetime (cuda_scheduler ctor, nufi/cuda_scheduler.hpp:43-63; cuda_kernel ctor nufi/cuda_kernel.cu:81-110,
 *      273-287, 468-483).  Allocates the (Nt+1)-level device history, rho, metrics on `device`
 *      (-1 = current device).  order = 3..8, the orders the reference instantiates (nufi/cuda_kernel.cu:191-203, 373-385,
 *      575-587).  order 4 (cubic; what every reference driver runs) takes the specialised kernels (per-cell polynomial level
 *      formats, shared-memory staging); the other orders run the generic Cox-de Boor kernel on the reference layout (global
 *      memory variant, one point per thread).  stride_t = prod_d (N_d + order - 1) everywhere below.
 *      f0 must have the period of the box: k*L_d = 2*pi*m for every dimension d, else ERR_ARG. ---- */
int nufi_b200_create_1d(const nufi_b200_config1d *conf, int order, const nufi_b200_f0 *f0, int device, nufi_b200_handle **out);
int nufi_b200_create_2d(const nufi_b200_config2d *conf, int order, const nufi_b200_f0 *f0, int device, nufi_b200_handle **out);
int nufi_b200_create_3d(const nufi_b200_config3d *conf, int order, const nufi_b200_f0 *f0, int device, nufi_b200_handle **out);
void nufi_b200_destroy(nufi_b200_handle *h);
/* message of the last failure on h (h == NULL: of the last failed create on this thread) */
const char *nufi_b200_last_error(const nufi_b200_handle *h);

/* ---- the reference scheduler's five methods ---- */
/* cuda_kernel::compute_rho (nufi/cuda_kernel.cu:112-132): asynchronous; partial rho over flat q in [q_begin,q_end)
 * at time step n from device levels 0..n-1.  n > Nt -> ERR_RANGE.  Empty range: rho partial = 0. */
int nufi_b200_compute_rho(nufi_b200_handle *h, size_t n, size_t q_begin, size_t q_end);
/* cuda_kernel::download_rho (:135-145): blocking; rho_host[l] += partial[l], l < Nx*Ny*Nz. */
int nufi_b200_download_rho(nufi_b200_handle *h, double *rho_host);
/* cuda_kernel::upload_phi (:147-156): blocking; copies level n from the BASE pointer of the host history,
 * i.e. coeffs_base[n*stride_t .. (n+1)*stride_t). */
int nufi_b200_upload_phi(nufi_b200_handle *h, size_t n, const double *coeffs_base);
/* cuda_kernel::compute_metrics / download_metrics (:158-189): eval_f backtrace (needs levels 0..n), then
 * metrics[0..3] += {int f, int f^2, kinetic energy, entropy}.  Weights as the reference writes them. */
int nufi_b200_compute_metrics(nufi_b200_handle *h, size_t n, size_t q_begin, size_t q_end);
int nufi_b200_download_metrics(nufi_b200_handle *h, double *metrics4_host);
/* dim 1 only: the reference's second constructor cuda_scheduler(conf, conf_metrics) / cuda_kernel(conf, conf_metrics, dev)
 * (nufi/cuda_scheduler.hpp:65-85, nufi/cuda_kernel.cu:97-110) integrates the metrics over the (x,u) grid of conf_metrics --
 * nodes x_min + ix*dx, u_min + iu*du + du/2, weight du*dx, flat index q = ix*Nu + iu, all taken from conf_metrics -- while
 * eval_f uses the field grid of conf (nufi/cuda_kernel.cu:55-70).  After this call compute_metrics does that;
 * conf_metrics == NULL restores the handle's own grid. */
int nufi_b200_set_metrics_grid_1d(nufi_b200_handle *h, const nufi_b200_config1d *conf_metrics);

/* ---- CPU-driver-shaped entry points (bin/test_nufi_cpu_{1,2,3}d.cpp loop body) ---- */
/* the whole "#pragma omp parallel for: rho[l] = eval_rho(n,l,coeffs,conf)" sweep in one call; blocking;
 * CPU convention (leading 1); rho_host may be NULL (result stays on the device for solve_interpolate). */
int nufi_b200_eval_rho_all(nufi_b200_handle *h, size_t n, double *rho_host);
/* poisson::solve + interpolate on the device rho left by eval_rho_all / step: writes device level n;
 * *energy (may be NULL) = return value of solve.  Blocking iff energy != NULL. */
int nufi_b200_solve_interpolate(nufi_b200_handle *h, size_t n, double *energy);
/* same, but from a host rho (CPU convention), e.g. after an MPI/host reduction; blocking */
int nufi_b200_solve_interpolate_host(nufi_b200_handle *h, size_t n, const double *rho_host, double *energy);

/* dimN::poisson<double>::solve alone (nufi/poisson.cpp:66-89, 190-219, 328-362): data (Nx*Ny*Nz nodal values of rho, host)
 * is overwritten with phi at the nodes; *energy (may be NULL) = the returned electric energy.  Blocking. */
int nufi_b200_poisson_solve(nufi_b200_handle *h, double *data_host, double *energy);
/* dimN::interpolate<double,4> alone (nufi/fields.hpp:63-142, 186-300, 352-490): nodal values (host) -> one level of
 * coefficients with the (order-1) periodic halo, reference layout, stride_t doubles (host).  Blocking; the device history
 * is not touched (push the level with nufi_b200_upload_phi as the reference loop does). */
int nufi_b200_interpolate(nufi_b200_handle *h, const double *values_host, double *coeffs_level_host);

/* ---- fused step: backtrace + reduce + Poisson + interpolate + store level n, no host round trip.
 *      Asynchronous; the electric energy of step n is kept on the device (nufi_b200_download_energy). ---- */
int nufi_b200_step(nufi_b200_handle *h, size_t n);
/* The same step for a caller that keeps the history on the HOST -- the loop body of the reference's GPU drivers in one call
 * (upload_phi(n-1) / compute_rho / download_rho / [MPI_Allreduce] / poisson.solve / interpolate, bin/test_nufi_gpu_3d.cpp:154-162):
 * level n-1 is copied from coeffs_base to the device (n > 0; lower levels were pushed by earlier calls or upload_phi), the fused
 * step runs, level n is written to coeffs_base + n*stride_t, rho (CPU convention, may be NULL) and the electric energy (may be
 * NULL) come back with it; one stream synchronisation, blocking.  peer != 0: the multi-GPU step (nufi_b200_peer_step). */
int nufi_b200_step_host(nufi_b200_handle *h, size_t n, double *coeffs_base, double *rho_host, double *energy, int peer);
/* blocking; energies[i] = electric energy of step n_begin+i, for steps run by step/solve_interpolate */
int nufi_b200_download_energy(nufi_b200_handle *h, size_t n_begin, size_t n_end, double *energies);
/* blocking; rho of the most recent step / peer_step / group_step / eval_rho_all as the field tail consumed it: CPU convention
 * (1 - dV * sum f, nufi/rho.hpp:145), Nx*Ny*Nz doubles; in a multi-GPU step the SUM over all ranks' shares (identical on every
 * GPU) -- what the reference holds in its host rho after download_rho + MPI_Allreduce (bin/test_nufi_gpu_3d.cpp:154-158) */
int nufi_b200_download_rho_full(nufi_b200_handle *h, double *rho_host);
/* blocking; level n (stride_t doubles, reference layout with halo) to the host */
int nufi_b200_download_phi(nufi_b200_handle *h, size_t n, double *coeffs_level);
int nufi_b200_sync(nufi_b200_handle *h);

/* ---- checkpoint / restart: the coefficient history IS the complete state.  Levels 0..n_levels-1 in the
 *      reference layout, n_levels*stride_t doubles -- what bin/test_nufi_gpu_1d.cpp:239,364-366 and
 *      bin/test_nufi_cpu_3d_isolated.cpp:64-73,160-163 write as text (readers/writers: include/nufi/history_io.hpp).
 *      Restart = upload_history + continue stepping at n_levels: bit-identical to the uninterrupted run.  Blocking. ---- */
int nufi_b200_download_history(nufi_b200_handle *h, size_t n_levels, double *coeffs_host);
int nufi_b200_upload_history(nufi_b200_handle *h, size_t n_levels, const double *coeffs_host);

/* ---- sampling for plots / diagnostics (bin/test_nufi_gpu_1d.cpp:307-349, bin/test_nufi_gpu_2d.cpp:174-200), host buffers,
 *      blocking.  eval_f: f(t_n, x, v) at npts phase-space points, points = [npts][2*dim] (x.., v..); with_first_half_kick
 *      1 = eval_f (needs levels 0..n; nufi/rho.hpp:63-96, 234-281, 369-426), 0 = eval_ftilda (levels 0..n-1).
 *      eval_field: phi_n (derivative_axis = -1) or d phi_n / d x_axis at npts positions, points = [npts][dim]
 *      (nufi/fields.hpp:36-61, 149-184, 308-350). ---- */
int nufi_b200_eval_f(nufi_b200_handle *h, size_t n, size_t npts, const double *points_host, double *f_host, int with_first_half_kick);
int nufi_b200_eval_field(nufi_b200_handle *h, size_t n, int derivative_axis, size_t npts, const double *points_host, double *values_host);
/* eval_phase_flow (nufi/rho.hpp:98-131, dim1 in the reference; bin/test_nufi_gpu_1d.cpp:155,190,339): the flow map itself --
 * feet[i] = (x.., v..) at t = 0 of the characteristic through points[i] = (x.., v..) at t_n (needs levels 0..n), both
 * [npts][2*dim].  As in the reference nothing is traced for n <= 1, and positions are reduced with L*floor(x/L).  The device
 * traces the periodic image inside [x_min, x_min + L), so the returned position is the reference's whenever x_min is a multiple
 * of L (every reference configuration has x_min = 0); otherwise it is that position modulo L, shifted into [0, L). */
int nufi_b200_eval_phase_flow(nufi_b200_handle *h, size_t n, size_t npts, const double *points_host, double *feet_host);

/* ---- device-pointer plumbing for one-process-per-GPU callers (torch.distributed / NCCL host layer) ---- */
/* run all work of h on this cudaStream_t (NULL = the handle's own stream) */
int nufi_b200_set_stream(nufi_b200_handle *h, void *cuda_stream);
/* device address of the partial rho (GPU convention, Nx*Ny*Nz doubles) written by compute_rho */
int nufi_b200_rho_device(nufi_b200_handle *h, double **d_rho);
/* field tail from a device vector holding the SUM over all shards of the partial rho (GPU convention):
 * rho = 1 + sum, solve, interpolate, store level n, record energy[n].  Asynchronous. */
int nufi_b200_field_tail_device(nufi_b200_handle *h, size_t n, const double *d_rho_partial_sum);

/* ---- multi-GPU step with the exchange of the partial rho FUSED into the path's kernels, over NVLink/NVSwitch peer memory
 *      (replaces cuda_kernel::download_rho + host add, nufi/cuda_kernel.cu:135-145, nufi/cuda_scheduler.hpp:113-118, and the
 *      MPI_Allreduce of bin/test_nufi_gpu_3d.cpp:158).  One process per GPU: every rank calls peer_export (allocates its exchange
 *      buffer, returns a 64-byte CUDA IPC handle), the host layer all-gathers the handles (rank order), every rank calls
 *      peer_attach.  peer_step(n) = backtrace of this rank's share of the quadrature points -- velocity nodes rank, rank+world,
 *      ... of every spatial node, so all GPUs trace statistically identical samples of phase space (the reference's contiguous
 *      flat-q split, nufi/cuda_scheduler.hpp:88-111, stays available through compute_rho(q_begin, q_end)); inside that kernel the
 *      last CTA to finish a tile of 32 spatial nodes adds the tile's partial sums in a fixed order and STORES them into every
 *      GPU's buffer, each 8-byte word carrying the step's epoch beside 32 bits of payload -- no fence to system scope, no flag, no
 *      serial epilogue on the sender -> field tail that adds the ranks' sums in rank order, polling each word until it carries
 *      this step's epoch, without waiting for its own GPU's backtrace grid to retire (bit-identical on all
 *      GPUs) -> level n.  Asynchronous, no collective call, no host synchronisation.  All ranks must call peer_step the same
 *      number of times; synchronise all ranks (host barrier) before peer_detach / destroy.  peer_status: blocking; *timed_out
 *      = 1 if a wait for a peer's flag ever gave up (a rank died or skipped a step; results are then invalid). ---- */
#define NUFI_B200_PEER_HANDLE_BYTES 64
int nufi_b200_peer_export(nufi_b200_handle *h, int world, void *ipc_handle);
int nufi_b200_peer_attach(nufi_b200_handle *h, int rank, int world, const void *ipc_handles /* world x 64 bytes, rank order */);
int nufi_b200_peer_step(nufi_b200_handle *h, size_t n);
int nufi_b200_peer_status(nufi_b200_handle *h, int *timed_out);
int nufi_b200_peer_detach(nufi_b200_handle *h);

/* ---- several GPUs driven by ONE process (the reference's cuda_scheduler shape: one host thread, all visible devices,
 *      nufi/cuda_scheduler.hpp:43-63).  A group ties one handle per device together with NCCL communicators
 *      (ncclCommInitAll; libnccl.so.2 is loaded on first use) or with direct peer access.  group_step = per device: backtrace of its contiguous
 *      share of the flat q range (the reference's split, cuda_scheduler.hpp:88-111) -> ncclAllReduce(sum) of the
 *      partial rho in place on the devices (replaces download_rho + host add + MPI_Allreduce,
 *      bin/test_nufi_gpu_3d.cpp:154-158) -> the replicated field tail on every device.  Asynchronous. ---- */
typedef struct nufi_b200_group nufi_b200_group;
int nufi_b200_group_create(nufi_b200_handle *const *handles, int n_handles, nufi_b200_group **out);
void nufi_b200_group_destroy(nufi_b200_group *g);
int nufi_b200_group_step(nufi_b200_group *g, size_t n);
int nufi_b200_group_sync(nufi_b200_group *g);
/* how group_step exchanges the partial rho: 0 = peer memory, fused into the kernels as in peer_step (default whenever every
 * device of the group can map every other one), 1 = NCCL all-reduce between backtrace and tail.  group_exchange names it. */
int nufi_b200_group_set_exchange(nufi_b200_group *g, int mode);
const char *nufi_b200_group_exchange(const nufi_b200_group *g);
const char *nufi_b200_group_last_error(const nufi_b200_group *g);

/* ---- introspection used by bench.py / tests ---- */
/* number of visible CUDA devices (cuda::device_count, nufi/cuda_runtime.hpp:196-203) */
int nufi_b200_device_count(int *count);
/* device index a handle lives on (-1 for NULL) */
int nufi_b200_device_of(const nufi_b200_handle *h);
/* number of kernels this library launched on h since creation */
uint64_t nufi_b200_launch_count(const nufi_b200_handle *h);
/* on != 0: bracket every backtrace launch with a CUDA event pair on its stream (read back lazily by the two calls below).
 * Off by default: an event between the backtrace kernel and the field tail would keep the tail from being launched
 * programmatically behind it (griddepcontrol), which hides its launch latency. */
int nufi_b200_set_kernel_timing(nufi_b200_handle *h, int on);
/* GPU time in ms of the most recent timed backtrace kernel; blocking */
int nufi_b200_last_backtrace_ms(nufi_b200_handle *h, float *ms);
/* accumulated GPU time and number of backtrace kernel launches since the last reset (every launch is bracketed by a
 * CUDA event pair on its stream; the pairs are read back lazily, so this call blocks, the launches do not) */
int nufi_b200_backtrace_time(nufi_b200_handle *h, double *total_ms, uint64_t *count, int reset);
/* which kernel variant the last compute_rho used: writes a short static string ("smem-tma", "global") */
const char *nufi_b200_last_variant(const nufi_b200_handle *h);
/* force a variant for A/B tests: 0 auto, 1 global-memory path, 2 shared-memory staged path */
int nufi_b200_set_variant(nufi_b200_handle *h, int variant);
/* lane layout of the backtrace kernel: nodes per tile = 32 (every lane of a warp its own spatial node, x-neighbours, one
 * velocity: rigid drift, conflict-free at any history depth) or 16/8/4/2/1 (32/TN lanes per node with neighbouring velocities:
 * their loads coincide early in the history and are served as broadcasts -- the lane layout of the reference's kernel,
 * nufi/cuda_kernel.cu:40-41, is TN = 1).  0 = automatic. */
int nufi_b200_set_tile_nodes(nufi_b200_handle *h, int nodes_per_tile);
/* field-tail implementation: 0 auto (fused single-CTA kernel for grids up to 4096 nodes, cuFFT otherwise),
 * 1 cuFFT D2Z + symbol + Z2D + expand, 2 fused single-CTA kernel */
int nufi_b200_set_tail_variant(nufi_b200_handle *h, int variant);
const char *nufi_b200_last_tail_variant(const nufi_b200_handle *h);
/* register-only DFMA loop on `device`: measured FP64 peak in TFLOP/s (FMA = 2 flop) */
int nufi_b200_measure_fp64_peak(int device, double *tflops);
const char *nufi_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NUFI_B200_H */
