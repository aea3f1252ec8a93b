// internal.cuh -- shared declarations of libnufi_b200 (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>
#include <cufft.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/nufi_b200.h"

namespace nufi_b200
{

// ----------------------------------------------------------------------------------------------
// Backtrace kernel parameters.  Positions are carried per dimension as (cell k, centred offset
// tau in [-1/2,1/2]) with xi = (x - x_min)/dx = k + 1/2 + tau; velocities stay physical.
// ----------------------------------------------------------------------------------------------
struct BtParams
{
    int dim;
    int Nx, Ny, Nz, Nu, Nv, Nw;
    int sx, sxy;                     // row / plane stride of a 2d/3d device level (doubles)
    unsigned int level_bytes;        // bytes of one device level (multiple of 16)
    const double *hist;              // device history, device level format, level m at m*level_bytes
    int first_level;                 // newest level read: n-1 (rho) or n (metrics); -1: none
    int metrics;                     // 0: eval_ftilda -> rho slots; 1: eval_f -> metric partials
    double ncx, ncy, ncz;            // -dt*dx_inv: cells per step per unit velocity, negated (drift is backwards)
    double gx, gy, gz;               // kick factor -dt*dx_inv / (12 | 72) (2d | 3d basis scaling folded in); xpp: gx / 3
    double x_min, y_min, z_min, dx, dy, dz;
    double u0, v0, w0, du, dv, dw;   // first midpoint velocity node and spacing, computed as rho.hpp does
    double ug0, vg0, wg0, dug, dvg, dwg; // metrics: GPU-form nodes u_min + i*du + du/2 (conf.du)
    int f0_kind;
    double f0p[4];
    // 1d metrics on a grid of their own (nufi/cuda_kernel.cu:55-70): node l sits at mx_min + l*mdx and is located in the field grid
    int mgrid;
    double mx_min, mdx, Lx, Lx_inv, dx_inv;
    // work decomposition (see backtrace.cu)
    unsigned long long q_begin, q_end, Nvel;
    unsigned long long Nvel_loc, vstride, voff; // this launch traces velocity nodes voff, voff+vstride, ... (Nvel_loc of them)
    unsigned long long l_first;      // first spatial node touched by [q_begin,q_end)
    unsigned long long l_last;       // last spatial node touched
    unsigned int n_tiles;            // tiles of TN consecutive nodes
    unsigned int TN, TNlog2;         // nodes per tile (32, 8, 4, 2 or 1): lane -> node lane % TN, velocity sub-index lane / TN
    unsigned int upt;                // warp-units per tile = ceil(ceil(Nvel_loc / (32/TN)) / ILP)
    unsigned int W;                  // consumer warps per CTA
    unsigned int rpt;                // CTA-rounds per tile = ceil(upt / W)
    unsigned int R;                  // CTA-rounds in total = rpt * n_tiles
    unsigned int rpc;                // CTA-rounds per CTA = ceil(R / grid)
    unsigned int Tmax;               // slots per CTA (tiles one CTA can touch)
    int interleave;                  // 1: warp-unit jc = jr + warp*rpt, point i of a thread at velocity jc + i*upt (balanced mix)
    double *slots;                   // [grid][Tmax][32]                (rho)
    uint4 *slots_ll;                 // non-null: the slots go here instead, each as a self-validating word pair carrying slot_flag
    unsigned int slot_flag;          //           (the fused tail polls them instead of waiting for this grid to retire)
    double *mpartials;               // [grid][4]                       (metrics)
    double mweight;
    // staged variant: shared-memory ring of `stages` stages, each a chunk of Lc consecutive levels
    int Lc, stages;
    unsigned int stage_bytes;
    int cluster;                     // 2: CTAs 2k, 2k+1 form a thread-block cluster and share every chunk (multicast halves); else 1
};

// sampling kernels (sample_f_kernel): f / ftilda / the flow map at arbitrary phase-space points
struct SampleParams
{
    BtParams P;
    double Lx, Ly, Lz, Lx_inv, Ly_inv, Lz_inv, dx_inv, dy_inv, dz_inv;
    const double *pts; // [npts][2*dim]: x.., v..
    double *out;
    size_t npts;
    int with_first_half_kick; // 1: eval_f (nufi/rho.hpp:63-96, 234-281, 369-426), 0: eval_ftilda
    int feet;                 // 1: write the foot (x.., v..) of the characteristic instead of f0 there (eval_phase_flow, rho.hpp:98-131)
};

struct FinishParams
{
    const double *slots;
    const uint4 *slots_ll; // non-null: self-validating slots (see BtParams::slots_ll); polled until they carry slot_flag
    unsigned int slot_flag;
    int *status;           // set to 1 if a polled word never arrived (bounded spin)
    double *rho_partial; // GPU convention: -dV * sum
    double *rho_full;    // CPU convention: 1 - dV * sum  (may be nullptr)
    double dV;
    unsigned long long l_first, l_last;
    unsigned int rpt, rpc, Tmax, n_tiles;
    unsigned int TN;     // nodes per tile: node of (tile, lane) = l_first + tile*TN + lane, lanes >= TN carry zeros
};

// ----------------------------------------------------------------------------------------------
// Self-validating words ("low latency" encoding) and the peer exchange of the partial rho over NVLink/NVSwitch peer memory.
//
// A double travels as one 16-byte vector store {lo, epoch, hi, epoch}: every 8-byte half carries 32 bits of payload and the
// 32-bit epoch of the step (an aligned 8-byte store is delivered whole), so a reader needs neither a flag nor a fence on the
// writer's side, and need not wait for the writing grid to retire: it polls the words it is about to add until they carry this
// step's epoch.  Two uses:
//  (1) the backtrace kernel's per-(CTA, tile) slots in a fused step (BtParams::slots_ll): the one-CTA field tail, resident
//      beside the backtrace grid since its start (programmatic dependent launch), adds them in the fixed order of
//      finish_rho_kernel as they land -- no kernel-boundary wait on the step's critical path;
//  (2) the exchange between GPUs: every GPU owns a buffer [2 parities][world][n_nodes] x 16 B mapped into all peers (one process:
//      cudaDeviceEnablePeerAccess; one process per GPU: CUDA IPC).  Per step with epoch e (parity e&1) the tail (large grids:
//      finish_rho_kernel) of rank r STORES its per-node sums straight into region [e&1][r] of every other GPU as soon as it has
//      them, then adds all ranks' sums in rank order (bit-identical on every GPU), polling the peers' words.  No collective call,
//      no system-scope fence (3.8 us measured), no flag, no extra launch.  Two parities make the buffers safe to reuse: a rank
//      writes epoch e+2 only after its tail of e+1 has read every peer's words of e+1, which those peers pushed after their
//      tails of e had finished reading epoch e.
// ----------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 16;

struct PeerPush // sender side (tail_small_kernel / finish_rho_kernel)
{
    uint4 *rho[kMaxPeers];     // region [e&1][my rank] inside every GPU's exchange buffer: one uint4 {lo, epoch, hi, epoch} per node
    int world, rank;           // world 0: not a peer step
    unsigned int flag;         // this step's epoch (never 0: the buffers start zeroed)
};

struct PeerRecv // receiver side (tail_small_kernel / peer_gather_kernel)
{
    int world;                 // 0: not a peer step
    unsigned int flag;         // this step's epoch
    const uint4 *rho;          // local region [e&1][0]; rank r at + r * n_nodes
    size_t n_nodes;
    double dV;                 // rho = 1 - dV * (sum over ranks, in rank order)
    double *rho_full;          // CPU-convention rho for later downloads
    int *status;               // set to 1 if a word never arrived (bounded spin)
};

#ifdef __CUDACC__
// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may start
// while its predecessor still runs; it must execute pdl_wait() before touching anything the predecessor writes (the wait
// returns once the predecessor grid has completed and its memory is visible).  pdl_trigger() in the predecessor lets the
// dependent grid be scheduled as soon as SM resources allow.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// s += x with the rounding error of the addition accumulated in lo (Knuth's branch-free two-sum; no fast-math: nvcc keeps the
// order of floating-point additions)
__device__ __forceinline__ void two_sum(double &s, double &lo, double x)
{
    const double t = s + x;
    const double xv = t - s;
    lo += (s - (t - xv)) + (x - xv);
    s = t;
}

// ---- the wire encoding
__device__ __forceinline__ void peer_store_double(uint4 *dst, double v, unsigned flag)
{
    const unsigned lo = static_cast<unsigned>(__double2loint(v)), hi = static_cast<unsigned>(__double2hiint(v));
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(lo), "r"(flag), "r"(hi), "r"(flag) : "memory");
}
// four slot loads in flight at once (null pointer: 0.0): v[u] = *src[u], polled like peer_rank_sum's
__device__ __forceinline__ void ll_load4(const uint4 *const (&src)[4], unsigned flag, int *status, double (&v)[4])
{
    unsigned a[4], fa[4], b[4], fb[4];
    long long t0 = 0;
    for (;;) {
        bool ok = true;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (src[u]) asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a[u]), "=r"(fa[u]), "=r"(b[u]), "=r"(fb[u]) : "l"(src[u]) : "memory");
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (src[u]) ok = ok && fa[u] == flag && fb[u] == flag;
        if (ok) break;
        if (t0 == 0) {
            if (*reinterpret_cast<volatile int *>(status) != 0) break;
            t0 = clock64();
        } else if (clock64() - t0 > (1ll << 35)) {
            *status = 1;
            break;
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = src[u] ? __hiloint2double(static_cast<int>(b[u]), static_cast<int>(a[u])) : 0.0;
}
// sum over the ranks, in rank order, of the word at `src + r * stride`: the loads poll (bounded, ~17 s; they give up at once after
// an earlier give-up) until both halves of every value carry this step's epoch.  Volatile loads: served by L2, where the peers'
// stores land.  Up to four ranks' loads are in flight at once.
// `self` >= 0: that rank's value is `own` (no load).
__device__ __forceinline__ double peer_rank_sum(const uint4 *src, size_t stride, int world, unsigned flag, int *status, int self, double own)
{
    double sum = 0, lo = 0; // compensated: the result is the ranks' sum rounded once
    for (int r0 = 0; r0 < world; r0 += 4) {
        unsigned a[4], fa[4], b[4], fb[4];
        long long t0 = 0;
        for (;;) {
            bool ok = true;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (r0 + u < world && r0 + u != self)
                    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a[u]), "=r"(fa[u]), "=r"(b[u]), "=r"(fb[u]) : "l"(src + (r0 + u) * stride) : "memory");
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (r0 + u < world && r0 + u != self) ok = ok && fa[u] == flag && fb[u] == flag;
            if (ok) break;
            if (t0 == 0) {
                if (*reinterpret_cast<volatile int *>(status) != 0) break;
                t0 = clock64();
            } else if (clock64() - t0 > (1ll << 35)) {
                *status = 1;
                break;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (r0 + u < world) two_sum(sum, lo, r0 + u == self ? own : __hiloint2double(static_cast<int>(b[u]), static_cast<int>(a[u])));
    }
    return sum + lo;
}
#endif

struct PeerState
{
    int world = 0, rank = -1;
    bool ipc = false;                   // peer buffers opened through CUDA IPC (must be closed), else direct peer access
    unsigned char *xb = nullptr;        // local exchange buffer
    size_t xb_bytes = 0;
    unsigned char *peer_xb[kMaxPeers] = {};
    unsigned long long epoch = 0;
    int *d_status = nullptr;
    PeerPush push{};                    // of the step being launched
    PeerRecv recv{};
};

// In-kernel epilogue of the backtrace kernel (rho mode).  mode 1: the last CTA to finish adds the slots of every tile in a fixed
// order (what finish_rho_kernel does as a separate launch).  mode 0: none -- the fused tail (or finish_rho_kernel) adds the slots.
struct EpilogueParams
{
    int mode;                 // 0 none, 1 finish (last-CTA slot reduction)
    unsigned int n_active;    // CTAs that take part (blockIdx.x < n_active)
    unsigned int *done;       // device counter, zero between launches
    FinishParams F;
};

struct Handle
{
    int dim = 0, order = 4, device = 0;
    nufi_b200_config3d c{}; // superset; unused dimensions have N = 1
    nufi_b200_f0 f0{};
    size_t Nt = 0;
    size_t n_nodes = 0, n_vel = 0, stride_t = 0; // reference-format level size
    // device level format
    int sx = 0, sxy = 0;
    bool xpp = false;        // 2d/3d: levels stored as per-(row, cell) cubics in the x offset (see tail.cu: row_poly)
    size_t level_stride = 0; // doubles
    size_t raw_stride = 0;   // 1d only: raw spline level kept beside the pp-form (doubles)
    double *d_hist = nullptr, *d_raw = nullptr;
    std::vector<unsigned char> level_valid;
    // rho / reduction
    double *d_rho_partial = nullptr, *d_rho_full = nullptr, *d_partials = nullptr;
    size_t partials_cap = 0; // doubles
    uint4 *d_partials_ll = nullptr; // the same slots as self-validating words (fused steps on grids the one-CTA tail handles)
    size_t partials_ll_cap = 0;
    unsigned int slot_epoch = 0;    // epoch of the last launch that wrote d_partials_ll (never 0: the buffer starts zeroed)
    int *d_ll_status = nullptr;     // set to 1 by a polling load that gave up
    double *d_metrics = nullptr, *d_mpartials = nullptr;
    double *d_energy = nullptr; // [Nt+1]
    double *d_stage = nullptr;  // device staging for one reference-format level
    double *h_pinned = nullptr; // pinned host staging (max(stride_t, n_nodes) doubles)
    size_t h_pinned_cap = 0;
    double *h_up = nullptr;     // second pinned staging buffer, host -> device only: uploads return without a stream sync
    cudaEvent_t ev_up = nullptr; // recorded behind the last copy out of h_up; waited for before h_up is overwritten
    bool ev_up_pending = false;
    // tail
    cufftHandle plan_fwd = 0, plan_inv = 0;
    bool plans = false;
    cufftDoubleComplex *d_spec = nullptr;
    double *d_symbol = nullptr; // real part tables, see tail.cu
    double *d_twiddle = nullptr; // per-dimension (cos, sin)(2 pi m / N_d) for the fused small-grid tail
    int tail_force = 0;          // 0 auto, 1 cuFFT path, 2 fused single-CTA path
    const char *last_tail = "none";
    FinishParams fin{};          // slot reduction of the last backtrace launch
    bool fin_pending = false;    // ... not yet run (the fused tail kernel does it itself)
    double *d_field = nullptr;
    double *d_epart = nullptr;
    size_t n_spec = 0;
    // streams / events
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // backtrace kernel timing: ring of CUDA event pairs on the launching stream, drained lazily
    std::vector<cudaEvent_t> ev_ring; // 2 per launch: [start, stop]
    size_t ev_head = 0, ev_pending = 0; // next slot (in pairs), pairs recorded but not yet read
    double bt_ms_total = 0, bt_ms_last = 0;
    uint64_t bt_count = 0;
    bool kernel_timing = false;  // bracket every backtrace launch with a CUDA event pair (nufi_b200_set_kernel_timing)
    bool pdl = true;             // programmatic dependent launch of finish / tail behind the backtrace kernel (NUFI_B200_PDL=0: off)
    int sm_count = 148;
    int pair_ctas = 0;           // CTAs of the persistent grid that can be resident as 2-CTA clusters (0: clusters unavailable)
    size_t smem_optin = 0;
    unsigned int *d_done = nullptr; // arrival counter of the backtrace kernel's last-CTA epilogue
    unsigned long long vstride = 1, voff = 0; // velocity share of the next backtrace launch (multi-GPU step), else 1, 0
    PeerState px;
    bool peer_push = false;      // the next backtrace launch is this GPU's share of a multi-GPU step: its sums go to every GPU (peer_step)
    int tn_force = 0;            // nodes per tile forced by nufi_b200_set_tile_nodes (0: automatic)
    bool mgrid_set = false;      // 1d: compute_metrics integrates over `mconf`'s (x,u) grid instead of the field grid's
    nufi_b200_config1d mconf{};
    int variant_force = 0;
    const char *last_variant = "none";
    char variant_buf[64] = {0};
    uint64_t launches = 0;
    std::string err;
};

// Launch `kern` on h->stream as a programmatic dependent of whatever precedes it there (safe after any predecessor: the
// kernel itself calls pdl_wait() before it reads or writes anything a predecessor touches).
template <typename... KArgs, typename... Args>
cudaError_t launch_chained(Handle *h, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args &&...args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = h->pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// error helpers (api.cu)
int fail(Handle *h, int code, const std::string &msg);
#define NUFI_CUDA_CHECK(h, expr)                                                                      \
    do {                                                                                              \
        cudaError_t e__ = (expr);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return ::nufi_b200::fail((h), NUFI_B200_ERR_CUDA,                                         \
                                     std::string(cudaGetErrorName(e__)) + ": " + cudaGetErrorString(e__) + \
                                         " [" #expr "]");                                             \
    } while (0)

// api.cu: event ring
constexpr size_t kEvRingPairs = 256;
int ev_acquire(Handle *h, cudaEvent_t *start, cudaEvent_t *stop); // next pair (drains the oldest when the ring is full)
int ev_drain(Handle *h);                                          // blocking: fold all pending pairs into the totals
// backtrace_generic.cu: the backtrace / sampling kernels for spline orders 3, 5..8 (global-memory variant, one point per thread)
cudaError_t launch_backtrace_generic(int order, int dim, const BtParams &P, const EpilogueParams &E, unsigned grid, unsigned threads,
                                     size_t smem_bytes, cudaStream_t st);
int generic_max_threads(int dim);
cudaError_t launch_sample_f_generic(int order, int dim, const SampleParams &S, unsigned blocks, cudaStream_t st);
// backtrace.cu
// defer_finish: leave the slot reduction to the field tail (fused step); otherwise finish_rho_kernel is launched
int launch_backtrace(Handle *h, size_t n, size_t q_begin, size_t q_end, bool metrics, bool defer_finish = false);
int launch_finish(Handle *h);
// sampling at arbitrary points (device pointers): f / ftilda at phase-space points, phi or a first derivative at positions
// feet: write the foot (x.., v..) of every characteristic ([npts][2 dim]) instead of f0 at the foot
int launch_sample_f(Handle *h, size_t n, size_t npts, const double *d_pts, double *d_out, bool full, bool feet = false);
int launch_sample_field(Handle *h, const double *d_ref_level, int der, size_t npts, const double *d_pts, double *d_out); // runs the pending slot reduction into d_rho_partial / d_rho_full
// tail.cu
int tail_init(Handle *h);
bool tail_is_small(const Handle *h); // the fused one-CTA tail will run (grid small enough, or forced)
void tail_destroy(Handle *h);
// d_rho_full == nullptr: take rho from the pending slot reduction of the last backtrace launch
// from_peer: rho = 1 + sum over ranks of the exchange buffer (h->px.recv), waited for inside the tail
int tail_run(Handle *h, size_t n, const double *d_rho_full, bool from_peer = false);
// peer.cu
void peer_free(Handle *h);
int peer_alloc(Handle *h, int world);
int peer_prepare_step(Handle *h);                 // next epoch: fills h->px.push / h->px.recv
int launch_peer_gather(Handle *h);                // exchange buffer -> d_rho_full (large grids, cuFFT tail)
int launch_peer_noop(Handle *h);                  // a rank without work still owes every GPU its (zero) sums of this epoch
int tail_filter(Handle *h, const double *d_values, int mode); // 1: poisson solve, 2: interpolate; result in d_field
int expand_field_to_stage(Handle *h);
double *tail_energy_scratch(Handle *h);
int convert_level_to_device(Handle *h, size_t n, const double *d_ref_level);  // reference format -> device format
int convert_level_from_device(Handle *h, size_t n, double *d_ref_level);      // device format -> reference format
int make_full_rho(Handle *h, const double *d_partial_sum, double *d_full);    // full = 1 + partial
// peak.cu
int measure_fp64_peak(int device, double *tflops);

} // namespace nufi_b200
