// backtrace.cu -- the hot path: backward Stoermer-Verlet through the coefficient history + f0 + rho reduction.
//
// Replaces nufi/cuda_kernel.cu:31-51, 210-237, 393-426 (cuda_eval_rho) and :53-79, 239-271, 428-466
// (cuda_eval_metrics), i.e. the device flavour of nufi/rho.hpp eval_ftilda / eval_f / eval_rho with
// nufi/fields.hpp eval and nufi/splines.hpp inlined.  Written from scratch for sm_100a:
//
//  * Work layout.  A warp-unit is 32 consecutive spatial nodes (x fastest, a "tile") sharing ILP consecutive
//    velocity nodes, so the lanes of a warp drift rigidly and read neighbouring coefficients (conflict-free)
//    for the whole history -- instead of the reference's "velocity fastest" layout whose lanes fan out and all
//    hit one rho address with atomics.  A CTA-round is W warp-units of the SAME tile; persistent CTAs (one per
//    SM) take contiguous runs of CTA-rounds.  Each thread sums f of its node over its velocities in registers;
//    when the CTA's tile changes the consumer warps combine their sums through shared memory in a fixed order
//    and write ONE slot per (CTA, tile); the slots of a tile are added in a fixed order by the fused field tail (which
//    polls them as self-validating words while this grid still runs), by this kernel's last CTA, or by
//    finish_rho_kernel.  The per-thread and per-CTA sums are compensated (two-sum).  No atomics on data: results
//    are run-to-run deterministic.
//  * History access.  Staged variant: a producer warp streams the history newest -> oldest from the HBM/L2-
//    resident ring into a shared-memory ring of stages with cp.async.bulk (TMA bulk copy, SASS UBLKCP) +
//    mbarrier full/empty pairs; a stage holds a CHUNK of several consecutive levels so barrier traffic and loop
//    bookkeeping are amortised over the chunk.  Global variant (levels too big for shared memory, large 3d):
//    read-only loads served by L1/L2.
//  * Arithmetic.  Position per dimension = (cell k, centred offset tau in [-1/2,1/2]); floor() and the
//    float->int conversion (quarter-rate pipes) are replaced by the 1.5*2^52 rounding trick on the FP64 pipe.
//    The regular full-kick step is the only thing in the inner loop: eval_f's initial half kick and the final
//    half kick on level 0 are peeled.  1d levels are stored as per-cell quadratics of dt*E ([Nx x (p1,p2)] [Nx x p0],
//    see tail.cu), so a 1d point-step is one 128-bit and one 64-bit load and 7 FP64 instructions; 2d/3d use the
//    cubic B-spline window (16/64 doubles) with value and derivative bases computed once per dimension and
//    shared by the field components, or the xpp format (per-row cubics in the x offset: Horner instead of basis).
//  * Launch.  Programmatic dependent launch in both directions: the field tail becomes resident while this grid
//    runs; the next step's grid becomes resident while that tail runs and waits (griddepcontrol.wait) before its
//    first global read.
#include "backtrace_kernel.cuh"

namespace nufi_b200
{

namespace
{

// Adds the per-(CTA, tile) slots of each tile in a fixed order.  One block (8 warps) per tile of 32 nodes.  Multi-GPU step on a
// grid too large for the one-CTA tail (X.world > 0): the sums also go into every GPU's exchange buffer (NVLink stores of
// self-validating words, internal.cuh), where peer_gather_kernel adds them in rank order.
__global__ void __launch_bounds__(256) finish_rho_kernel(const __grid_constant__ FinishParams F, const __grid_constant__ PeerPush X)
{
    __shared__ double part[8][32];
    pdl_wait();
    pdl_trigger();
    const unsigned tile = blockIdx.x;
    const int lane = threadIdx.x & 31;
    const unsigned wj = threadIdx.x >> 5;
    const unsigned b_lo = (tile * F.rpt) / F.rpc;
    const unsigned b_hi = ((tile + 1) * F.rpt - 1) / F.rpc;
    double sum = 0;
    for (unsigned b = b_lo + wj; b <= b_hi; b += 8) {
        const unsigned t_first = (b * F.rpc) / F.rpt;
        sum += F.slots[(static_cast<size_t>(b) * F.Tmax + (tile - t_first)) * 32 + lane];
    }
    part[wj][lane] = sum;
    __syncthreads();
    if (wj == 0) {
        double tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) tot += part[w][lane];
        const unsigned long long l = F.l_first + static_cast<unsigned long long>(tile) * F.TN + lane;
        if (static_cast<unsigned>(lane) < F.TN && l <= F.l_last) {
            F.rho_partial[l] = -F.dV * tot;
            if (F.rho_full) F.rho_full[l] = fma(-F.dV, tot, 1.0);
            for (int p = 0; p < X.world; ++p) peer_store_double(X.rho[p] + l, tot, X.flag);
        }
    }
}

// A rank without work in a multi-GPU step (more ranks than velocity nodes) still owes every GPU its (zero) sums of this epoch.
__global__ void __launch_bounds__(256) peer_noop_kernel(const __grid_constant__ PeerPush X, size_t n_nodes)
{
    pdl_wait();
    pdl_trigger();
    for (size_t l = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; l < n_nodes; l += static_cast<size_t>(gridDim.x) * blockDim.x)
        for (int p = 0; p < X.world; ++p) peer_store_double(X.rho[p] + l, 0.0, X.flag);
}

__global__ void finish_metrics_kernel(const double *mpartials, unsigned grid, double *metrics)
{
    if (threadIdx.x < 4) {
        double sum = 0;
        for (unsigned c = 0; c < grid; ++c) sum += mpartials[c * 4 + threadIdx.x];
        metrics[threadIdx.x] = sum;
    }
}

// ---------------------------------------------------------------- sampling at arbitrary points (plots, diagnostics)

struct FieldSampleParams
{
    int dim, order, Nx, Ny, Nz, der; // der: -1 value, 0/1/2 first derivative along x/y/z
    double x_min, y_min, z_min, Lx, Ly, Lz, Lx_inv, Ly_inv, Lz_inv, dx_inv, dy_inv, dz_inv;
    const double *level; // reference-format level: halo of order-1, row stride Nx+order-1
    const double *pts;   // [npts][dim]
    double *out;
    size_t npts;
};

// run-time-order flavour of basis_generic (not hot): values N[0..K) and first derivatives D[0..K) at t in [0,1]
__device__ void basis_rt(int K, double t, double *N, double *D)
{
    double b[8] = {1, 0, 0, 0, 0, 0, 0, 0};
    for (int p = 1; p < K; ++p) {
        if (p == K - 1) {
            D[0] = -b[0];
            for (int i = 1; i < K - 1; ++i) D[i] = b[i - 1] - b[i];
            D[K - 1] = b[K - 2];
        }
        const double inv = 1.0 / p;
        b[p] = (t * b[p - 1]) * inv;
        for (int i = p - 1; i >= 1; --i) b[i] = fma(t + (p - i), b[i - 1], ((1 + i) - t) * b[i]) * inv;
        b[0] = ((1.0 - t) * b[0]) * inv;
    }
    for (int i = 0; i < K; ++i) N[i] = b[i];
}

// phi_n or one first derivative at arbitrary points (nufi/fields.hpp eval<real,order,dx,dy,dz>).
__global__ void sample_field_kernel(const __grid_constant__ FieldSampleParams S)
{
    const int K = S.order;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < S.npts; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const double *q = S.pts + i * S.dim;
        int k[3] = {0, 0, 0};
        double tau[3] = {0, 0, 0};
        locate(q[0], S.x_min, S.Lx, S.Lx_inv, S.dx_inv, S.Nx, k[0], tau[0]);
        if (S.dim >= 2) locate(q[1], S.y_min, S.Ly, S.Ly_inv, S.dy_inv, S.Ny, k[1], tau[1]);
        if (S.dim >= 3) locate(q[2], S.z_min, S.Lz, S.Lz_inv, S.dz_inv, S.Nz, k[2], tau[2]);
        double W[3][8]; // per dimension: basis values (N_a) or derivatives (N'_a * dx_inv)
        const double inv[3] = {S.dx_inv, S.dy_inv, S.dz_inv};
        for (int d = 0; d < 3; ++d) {
            double N[8], D[8];
            basis_rt(K, 0.5 + tau[d], N, D);
            for (int a = 0; a < K; ++a) W[d][a] = d >= S.dim ? (a == 0 ? 1.0 : 0.0) : (S.der == d ? D[a] * inv[d] : N[a]);
        }
        const int sy = S.Nx + K - 1, sz = sy * (S.Ny + K - 1);
        const int nb = S.dim >= 2 ? K : 1, nc = S.dim >= 3 ? K : 1;
        double r = 0;
        for (int c = 0; c < nc; ++c)
            for (int b = 0; b < nb; ++b) {
                const double *row = S.level + static_cast<size_t>(k[2] + c) * sz + static_cast<size_t>(k[1] + b) * sy + k[0];
                double rv = row[0] * W[0][0];
                for (int a = 1; a < K; ++a) rv = fma(row[a], W[0][a], rv);
                r = fma(rv, W[1][b] * W[2][c], r);
            }
        S.out[i] = r;
    }
}

template <int DIM, int ILP, bool STAGED, bool POW2, bool XPP>
cudaError_t launch_variant(const BtParams &P, const EpilogueParams &E, unsigned grid, unsigned threads, size_t smem_bytes, cudaStream_t st, bool pdl)
{
    auto kern = backtrace_kernel<DIM, ILP, STAGED, POW2, XPP>;
    if (smem_bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes));
        if (e != cudaSuccess) return e;
    }
    // Programmatic dependent launch: the grid may become resident (barrier set-up, index arithmetic) while the kernel ahead of it
    // on the stream -- in a run of fused steps the previous step's one-CTA field tail -- is still running; the kernel executes
    // griddepcontrol.wait before it reads anything from global memory.
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    unsigned na = 0;
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (P.cluster == 2) { // CTAs 2k, 2k+1 on neighbouring SMs: every history chunk is fetched once per pair (multicast)
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kern, P, E);
}

template <int DIM, int ILP, bool XPP>
cudaError_t launch_fmt(const BtParams &P, const EpilogueParams &E, bool staged, bool pow2, unsigned grid, unsigned threads, size_t smem_bytes, cudaStream_t st, bool pdl)
{
    if (staged) {
        return pow2 ? launch_variant<DIM, ILP, true, true, XPP>(P, E, grid, threads, smem_bytes, st, pdl)
                    : launch_variant<DIM, ILP, true, false, XPP>(P, E, grid, threads, smem_bytes, st, pdl);
    }
    return pow2 ? launch_variant<DIM, ILP, false, true, XPP>(P, E, grid, threads, smem_bytes, st, pdl)
                : launch_variant<DIM, ILP, false, false, XPP>(P, E, grid, threads, smem_bytes, st, pdl);
}

template <int DIM, int ILP>
cudaError_t launch_ilp(const BtParams &P, const EpilogueParams &E, bool xpp, bool staged, bool pow2, unsigned grid, unsigned threads, size_t smem_bytes, cudaStream_t st, bool pdl)
{
    if constexpr (DIM >= 2) {
        if (xpp) return launch_fmt<DIM, ILP, true>(P, E, staged, pow2, grid, threads, smem_bytes, st, pdl);
    }
    return launch_fmt<DIM, ILP, false>(P, E, staged, pow2, grid, threads, smem_bytes, st, pdl);
}

template <int DIM>
cudaError_t launch_dim(const BtParams &P, const EpilogueParams &E, int ilp, bool xpp, bool staged, bool pow2, unsigned grid, unsigned threads, size_t smem_bytes,
                       cudaStream_t st, bool pdl)
{
    return ilp == 2 ? launch_ilp<DIM, 2>(P, E, xpp, staged, pow2, grid, threads, smem_bytes, st, pdl)
                    : launch_ilp<DIM, 1>(P, E, xpp, staged, pow2, grid, threads, smem_bytes, st, pdl);
}

int max_threads_for(int dim, int ilp, bool xpp)
{
    if (dim == 1) return ilp == 1 ? Tune<1, 1, false>::max_threads : Tune<1, 2, false>::max_threads;
    if (dim == 2) {
        if (xpp) return ilp == 1 ? Tune<2, 1, true>::max_threads : Tune<2, 2, true>::max_threads;
        return ilp == 1 ? Tune<2, 1, false>::max_threads : Tune<2, 2, false>::max_threads;
    }
    if (xpp) return ilp == 1 ? Tune<3, 1, true>::max_threads : Tune<3, 2, true>::max_threads;
    return ilp == 1 ? Tune<3, 1, false>::max_threads : Tune<3, 2, false>::max_threads;
}

bool is_pow2(size_t n) { return n && !(n & (n - 1)); }

int env_int(const char *name, int dflt)
{
    const char *v = std::getenv(name);
    return v && *v ? std::atoi(v) : dflt;
}

// How many CTAs of a one-CTA-per-SM grid can be resident as clusters of two (pairs need two SMs of one GPC): asked once per
// handle with a representative staged kernel at full shared-memory size; 0 = clusters cannot be used.
int query_pair_ctas(const Handle *h)
{
    auto kern = backtrace_kernel<1, 2, true, true, false>;
    const size_t smem = h->smem_optin > 2048 ? h->smem_optin - 1024 : 0;
    if (smem == 0 || cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(h->sm_count & ~1));
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return 2 * n;
}

// chains (warps x points per thread) per SM beyond which a CTA-round's time grows with its width
int saturation_chains(int dim) { return dim == 1 ? 32 : (dim == 2 ? 40 : 24); }

} // namespace

// Host side: decomposition + launch.  q is the reference's flat quadrature index (cuda_kernel.cu:40-41,
// 219-225, 402-412); [q_begin,q_end) may cut through a node's velocity range (masked per point).
int launch_backtrace(Handle *h, size_t n, size_t q_begin, size_t q_end, bool metrics, bool defer_finish)
{
    const nufi_b200_config3d &c = h->c;
    BtParams P{};
    P.dim = h->dim;
    P.Nx = static_cast<int>(c.Nx); P.Ny = static_cast<int>(c.Ny); P.Nz = static_cast<int>(c.Nz);
    P.Nu = static_cast<int>(c.Nu); P.Nv = static_cast<int>(c.Nv); P.Nw = static_cast<int>(c.Nw);
    P.sx = h->sx;
    P.sxy = h->sxy;
    P.hist = h->d_hist;
    P.first_level = metrics ? (n == 0 ? -1 : static_cast<int>(n)) : static_cast<int>(n) - 1;
    P.metrics = metrics ? 1 : 0;
    P.ncx = -(c.dt * c.dx_inv); P.ncy = -(c.dt * c.dy_inv); P.ncz = -(c.dt * c.dz_inv);
    // cubic steps: basis4 returns 6 N and 2 N', the factors go into the kick (2d: 6*2, 3d: 6*6*2); generic orders: true N, N'
    const double scale = h->order != 4 ? 1.0 : (h->dim == 2 ? 12.0 : 72.0);
    P.gx = -c.dt * c.dx_inv / scale / (h->xpp ? 3.0 : 1.0); P.gy = -c.dt * c.dy_inv / scale; P.gz = -c.dt * c.dz_inv / scale;
    P.x_min = c.x_min; P.y_min = c.y_min; P.z_min = c.z_min;
    P.dx = c.dx; P.dy = c.dy; P.dz = c.dz;
    // rho.hpp:136-137, 291-296, 441-447: du recomputed from the bounds, first node u_min + 0.5*du
    P.du = (c.u_max - c.u_min) / c.Nu; P.u0 = c.u_min + 0.5 * P.du;
    P.dv = (c.v_max - c.v_min) / c.Nv; P.v0 = c.v_min + 0.5 * P.dv;
    P.dw = (c.w_max - c.w_min) / c.Nw; P.w0 = c.w_min + 0.5 * P.dw;
    // cuda_kernel.cu:67, 259-260, 453-455: u_min + iu*du + du/2 with the stored conf.du
    P.dug = c.du; P.ug0 = c.u_min + c.du / 2;
    P.dvg = c.dv; P.vg0 = c.v_min + c.dv / 2;
    P.dwg = c.dw; P.wg0 = c.w_min + c.dw / 2;
    P.f0_kind = h->f0.kind;
    for (int i = 0; i < 4; ++i) P.f0p[i] = h->f0.p[i];
    // metric weights exactly as the reference writes them (cuda_kernel.cu:70, 262, 457)
    P.mweight = h->dim == 1 ? c.du * c.dx : (h->dim == 2 ? c.dx * c.dy * c.du * c.dv : c.du * c.dv * c.dw);

    P.Nvel = h->n_vel;
    if (metrics && h->mgrid_set) { // dim 1 only (nufi/cuda_kernel.cu:55-70): nodes and weights of the metrics grid, field of `conf`
        const nufi_b200_config1d &m = h->mconf;
        P.mgrid = 1;
        P.mx_min = m.x_min; P.mdx = m.dx;
        P.Lx = c.Lx; P.Lx_inv = c.Lx_inv; P.dx_inv = c.dx_inv;
        P.Nu = static_cast<int>(m.Nu);
        P.Nvel = m.Nu;
        P.dug = m.du; P.ug0 = m.u_min + m.du / 2;
        P.mweight = m.du * m.dx;
    }
    P.vstride = h->vstride > 0 ? h->vstride : 1;
    P.voff = h->voff;
    P.Nvel_loc = P.voff < P.Nvel ? (P.Nvel - P.voff + P.vstride - 1) / P.vstride : 0;
    if (P.Nvel_loc == 0) return fail(h, NUFI_B200_ERR_ARG, "velocity share of this GPU is empty");
    P.q_begin = q_begin; P.q_end = q_end;
    P.l_first = q_begin / P.Nvel;
    P.l_last = (q_end - 1) / P.Nvel;
    const unsigned long long n_nodes_range = P.l_last - P.l_first + 1;
    // nodes per tile: 32 = every lane its own node (rigid drift, conflict-free at any depth); fewer = 32/TN lanes per node with
    // neighbouring velocities, whose window loads coincide early in the history and are served as broadcasts (short 3d histories)
    unsigned TN = 32;
    {
        const int want = env_int("NUFI_B200_TN", 0);
        if (want == 1 || want == 2 || want == 4 || want == 8 || want == 16 || want == 32) TN = static_cast<unsigned>(want);
        else if (h->tn_force) TN = static_cast<unsigned>(h->tn_force);
    }
    P.TN = TN;
    P.TNlog2 = 0;
    while ((1u << P.TNlog2) < TN) ++P.TNlog2;
    const unsigned G = 32u / TN;
    const unsigned long long n_tiles64 = (n_nodes_range + TN - 1) / TN;
    if (n_tiles64 >= (1ull << 31)) return fail(h, NUFI_B200_ERR_RANGE, "too many tiles for one launch");
    P.n_tiles = static_cast<unsigned>(n_tiles64);
    P.level_bytes = static_cast<unsigned>(h->level_stride * 8);

    // ---- variant: stage the history through shared memory when at least two levels fit
    const size_t ring_budget = h->smem_optin > kSmemFixed + 1024 ? h->smem_optin - kSmemFixed - 1024 : 0;
    bool staged = P.first_level >= 0 && 2ull * P.level_bytes <= ring_budget;
    if (h->variant_force == 1 || h->order != 4) staged = false; // generic orders: window loads straight from global memory (L1/L2)
    if (h->variant_force == 2 && 2ull * P.level_bytes > ring_budget)
        return fail(h, NUFI_B200_ERR_ARG, "staged variant forced but two levels do not fit in shared memory");
    const bool pow2 = is_pow2(c.Nx) && is_pow2(c.Ny) && is_pow2(c.Nz);
    // Staged variant, optional (NUFI_B200_CLUSTER=2): the persistent grid runs as clusters of two CTAs that share every history
    // chunk (each fetches half, multicast to both), provided (nearly) all SMs can be paired.  Built to halve the L2 -> SM traffic
    // of the fill (every SM streams the same levels); measured on B200 it changes nothing -- C1 0.1023 / 0.1023 ms, C2 0.1472 /
    // 0.1491, C3 7.83 / 7.88, C4 0.238 / 0.237, C5-16 8.75 / 8.77 (off / on; profiles/r02_fill_path.md): the L2 already
    // de-duplicates concurrent requests of neighbouring SMs for one line, the fill is bounded by what ONE SM can take in -- so
    // it stays off by default and remains selectable (the parity suite runs it).
    const bool want_pairs = staged && env_int("NUFI_B200_CLUSTER", 1) == 2;
    if (want_pairs && h->pair_ctas == 0) {
        h->pair_ctas = query_pair_ctas(h);
        if (h->pair_ctas <= 0) h->pair_ctas = -1;
    }
    const bool paired = want_pairs && h->pair_ctas >= h->sm_count - 4 && h->pair_ctas >= 2;
    const unsigned grid = paired ? static_cast<unsigned>(std::min(h->pair_ctas, h->sm_count & ~1)) : static_cast<unsigned>(h->sm_count);
    P.cluster = paired ? 2 : 1;

    // ---- chunking of the staged history
    P.Lc = 1; P.stages = 0; P.stage_bytes = P.level_bytes;
    if (staged) {
        // as many levels per chunk as a two-stage ring allows (cap 16): the per-chunk bookkeeping (barrier wait/arrive, pointer
        // set-up) is paid once per Lc levels -- measured C1 0.117 -> 0.096 ms (Lc 5 -> 16), C3 8.65 -> 7.86 ms (Lc 1 -> 3)
        int Lc = static_cast<int>(ring_budget / (2ull * P.level_bytes));
        Lc = env_int("NUFI_B200_LC", Lc);
        Lc = Lc < 1 ? 1 : (Lc > 16 ? 16 : Lc);
        while (Lc > 1 && 2ull * Lc * P.level_bytes > ring_budget) --Lc;
        int stages = static_cast<int>(ring_budget / (static_cast<size_t>(Lc) * P.level_bytes));
        if (stages > kMaxStages) stages = kMaxStages;
        const int n_chunks = P.first_level / Lc + 1;
        if (stages > n_chunks) stages = n_chunks > 2 ? n_chunks : 2;
        P.Lc = Lc; P.stages = stages; P.stage_bytes = static_cast<unsigned>(Lc) * P.level_bytes;
    } else if (P.first_level >= 0) {
        P.Lc = P.first_level + 1; // global variant: the whole history is one "chunk" read in place
    }

    // ---- shape: points per thread (ILP) and consumer warps per CTA (W); a CTA-round = W warp-units of one tile
    int best_ilp = 1;
    unsigned best_W = 1;
    {
        double best_cost = 1e300;
        const int csat = saturation_chains(h->dim);
        const int force_ilp = env_int("NUFI_B200_ILP", 0), force_w = env_int("NUFI_B200_W", 0);
        for (int ilp = 1; ilp <= 2; ++ilp) {
            if (force_ilp && ilp != force_ilp) continue;
            // 3d: one point per thread at 128 registers (16 warps) beats two points at 255 (8 warps) -- measured
            if (!force_ilp && h->dim == 3 && ilp == 2) continue;
            if (h->order != 4 && ilp == 2) continue; // generic orders are instantiated with one point per thread
            const unsigned wmax = (h->order != 4 ? generic_max_threads(h->dim) : max_threads_for(h->dim, ilp, h->xpp)) / 32 - (staged ? 1 : 0);
            const unsigned long long upt = ((P.Nvel_loc + G - 1) / G + ilp - 1) / ilp;
            for (unsigned W = 1; W <= wmax; ++W) {
                if (force_w && static_cast<int>(W) != force_w) continue;
                const unsigned long long rpt = (upt + W - 1) / W;
                const unsigned long long R = rpt * P.n_tiles;
                const unsigned long long rpc = (R + grid - 1) / grid;
                const unsigned long long ctas = (R + rpc - 1) / rpc;
                const double chains = static_cast<double>(W) * ilp;
                // time of a CTA-round ~ max(latency floor, throughput term); ILP 2 shares the per-level bookkeeping (2d).  1d: the same
                // chains as twice the warps with one point each run as fast (C1) or faster (C2 0.1408 vs 0.1470 ms, W30 x 1 vs
                // W15 x 2, profiles/r02_sweep_shape_1d.txt): a replayed load stalls one chain, not two
                double cost = static_cast<double>(rpc) * (chains > csat ? chains : csat) * (ilp == 2 && h->dim != 1 ? 0.92 : 1.0);
                cost *= 1.0 + 1e-3 * (static_cast<double>(grid) - static_cast<double>(ctas)) / grid; // prefer more busy SMs
                cost *= 1.0 + 1e-4 * chains;                                                          // then narrower CTAs
                if (cost < best_cost) { best_cost = cost; best_ilp = ilp; best_W = W; }
            }
        }
        if (best_cost >= 1e300) return fail(h, NUFI_B200_ERR_ARG, "NUFI_B200_ILP / NUFI_B200_W override out of range");
    }
    const int ilp = best_ilp;
    P.W = best_W;
    P.upt = static_cast<unsigned>(((P.Nvel_loc + G - 1) / G + ilp - 1) / ilp);
    P.rpt = (P.upt + P.W - 1) / P.W;
    const unsigned long long R64 = static_cast<unsigned long long>(P.rpt) * P.n_tiles;
    if (R64 >= (1ull << 31)) return fail(h, NUFI_B200_ERR_RANGE, "quadrature range too large for one launch");
    P.R = static_cast<unsigned>(R64);
    P.rpc = (P.R + grid - 1) / grid;
    P.Tmax = (P.rpc - 1) / P.rpt + 2;
    P.interleave = env_int("NUFI_B200_INTERLEAVE", 1) ? 1 : 0;
    const unsigned threads = (P.W + (staged ? 1 : 0)) * 32;
    const size_t smem_bytes = kSmemFixed + (staged ? static_cast<size_t>(P.stages) * P.stage_bytes : 0);

    // ---- slots: one per (CTA, tile it touches); every slot the reduction reads is written by its CTA
    const bool push = !metrics && h->peer_push; // this GPU's share of a multi-GPU step: the sums go to every GPU
    const bool all_nodes = q_begin == 0 && q_end == h->n_nodes * h->n_vel; // every node's value gets written
    const bool whole = all_nodes && P.vstride == 1;                         // ... and it is the complete sum
    if (push && !all_nodes) return fail(h, NUFI_B200_ERR_ARG, "peer step: the backtrace must cover every spatial node");
    // who adds the slots: the fused one-CTA tail behind this launch (fused step on a small grid: it polls self-validating slots
    // instead of waiting for this grid to retire), finish_rho_kernel (many tiles, or a multi-GPU step on a large grid), or this
    // kernel's own last-CTA epilogue (compute_rho)
    const bool tail_reduces = !metrics && ((defer_finish && whole) || push) && tail_is_small(h);
    if (!metrics) {
        const size_t need = static_cast<size_t>(grid) * P.Tmax * 32;
        if (tail_reduces) {
            if (need > h->partials_ll_cap) {
                if (h->d_partials_ll) cudaFree(h->d_partials_ll);
                h->d_partials_ll = nullptr;
                h->partials_ll_cap = 0;
                if (cudaMalloc(&h->d_partials_ll, need * sizeof(uint4)) != cudaSuccess)
                    return fail(h, NUFI_B200_ERR_ALLOC, "cudaMalloc of the rho partial slots failed");
                NUFI_CUDA_CHECK(h, cudaMemsetAsync(h->d_partials_ll, 0, need * sizeof(uint4), h->stream)); // epoch 0 = "nothing yet"
                h->partials_ll_cap = need;
            }
            if (++h->slot_epoch == 0) h->slot_epoch = 1;
            P.slots_ll = h->d_partials_ll;
            P.slot_flag = h->slot_epoch;
        } else {
            if (need > h->partials_cap) {
                if (h->d_partials) cudaFree(h->d_partials);
                h->d_partials = nullptr;
                h->partials_cap = 0;
                if (cudaMalloc(&h->d_partials, need * sizeof(double)) != cudaSuccess)
                    return fail(h, NUFI_B200_ERR_ALLOC, "cudaMalloc of the rho partial slots failed");
                h->partials_cap = need;
            }
            P.slots = h->d_partials;
        }
    } else {
        P.mpartials = h->d_mpartials;
        NUFI_CUDA_CHECK(h, cudaMemsetAsync(h->d_mpartials, 0, sizeof(double) * 4 * grid, h->stream));
    }

    EpilogueParams E{};
    if (!metrics) {
        FinishParams F{};
        F.slots = P.slots;
        F.slots_ll = P.slots_ll;
        F.slot_flag = P.slot_flag;
        F.status = h->d_ll_status;
        F.rho_partial = h->d_rho_partial;
        F.rho_full = whole ? h->d_rho_full : nullptr;
        F.dV = h->dim == 1 ? P.du : (h->dim == 2 ? P.du * P.dv : P.du * P.dv * P.dw); // rho.hpp:145, 307, 459
        F.l_first = P.l_first; F.l_last = P.l_last;
        F.rpt = P.rpt; F.rpc = P.rpc; F.Tmax = P.Tmax;
        F.n_tiles = P.n_tiles;
        F.TN = P.TN;
        if (!all_nodes) NUFI_CUDA_CHECK(h, cudaMemsetAsync(h->d_rho_partial, 0, sizeof(double) * h->n_nodes, h->stream));
        h->fin = F;
        h->fin_pending = tail_reduces;
        // many tiles (TN < 32 on a large grid): the one-CTA epilogue would serialise them; use the multi-block finish kernel,
        // which is also the one that pushes to the peers
        const bool epilogue = P.n_tiles <= 256 && !push;
        if (!tail_reduces && !epilogue) {
            h->fin_pending = true; // launch_finish() below runs finish_rho_kernel
        } else if (!tail_reduces) {
            E.mode = 1;
            E.n_active = (P.R + P.rpc - 1) / P.rpc;
            E.done = h->d_done;
            E.F = F;
        }
    }

    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    if (h->kernel_timing) { // off by default: an event between two kernels keeps the second from launching programmatically
        int rc = ev_acquire(h, &ev_start, &ev_stop);
        if (rc) return rc;
        NUFI_CUDA_CHECK(h, cudaEventRecord(ev_start, h->stream));
    }
    cudaError_t e;
    if (h->order != 4) e = launch_backtrace_generic(h->order, h->dim, P, E, grid, threads, smem_bytes, h->stream);
    else if (h->dim == 1) e = launch_dim<1>(P, E, ilp, false, staged, pow2, grid, threads, smem_bytes, h->stream, h->pdl);
    else if (h->dim == 2) e = launch_dim<2>(P, E, ilp, h->xpp, staged, pow2, grid, threads, smem_bytes, h->stream, h->pdl);
    else e = launch_dim<3>(P, E, ilp, h->xpp, staged, pow2, grid, threads, smem_bytes, h->stream, h->pdl);
    NUFI_CUDA_CHECK(h, e);
    if (h->kernel_timing) {
        NUFI_CUDA_CHECK(h, cudaEventRecord(ev_stop, h->stream));
        h->ev_pending += 1;
    }
    h->launches += 1;
    const char *fmt = h->xpp ? "/xpp" : "";
    char tn[16] = "";
    if (P.TN != 32) std::snprintf(tn, sizeof(tn), "/tn%u", P.TN);
    char ord[16] = "";
    if (h->order != 4) std::snprintf(ord, sizeof(ord), "/order%d", h->order);
    if (staged) std::snprintf(h->variant_buf, sizeof(h->variant_buf), "smem-tma%s%s/ilp%d/W%u/Lc%dx%d%s", paired ? "-mc2" : "", fmt, ilp, P.W, P.Lc, P.stages, tn);
    else std::snprintf(h->variant_buf, sizeof(h->variant_buf), "global%s%s/ilp%d/W%u%s", fmt, ord, ilp, P.W, tn);
    h->last_variant = h->variant_buf;

    if (!metrics) {
        if (h->fin_pending && !tail_reduces) return launch_finish(h); // many tiles: multi-block slot reduction (+ push)
        return NUFI_B200_OK; // rho_partial (and rho_full) are complete when the kernel ends, or the tail reduces the slots
    } else {
        finish_metrics_kernel<<<1, 32, 0, h->stream>>>(h->d_mpartials, grid, h->d_metrics);
        NUFI_CUDA_CHECK(h, cudaGetLastError());
    }
    h->launches += 1;
    return NUFI_B200_OK;
}

// fills the geometry / f0 part of BtParams shared by every launch
static void fill_common(const Handle *h, BtParams &P)
{
    const nufi_b200_config3d &c = h->c;
    P.dim = h->dim;
    P.Nx = static_cast<int>(c.Nx); P.Ny = static_cast<int>(c.Ny); P.Nz = static_cast<int>(c.Nz);
    P.Nu = static_cast<int>(c.Nu); P.Nv = static_cast<int>(c.Nv); P.Nw = static_cast<int>(c.Nw);
    P.sx = h->sx; P.sxy = h->sxy;
    P.hist = h->d_hist;
    P.level_bytes = static_cast<unsigned>(h->level_stride * 8);
    P.ncx = -(c.dt * c.dx_inv); P.ncy = -(c.dt * c.dy_inv); P.ncz = -(c.dt * c.dz_inv);
    // cubic steps: basis4 returns 6 N and 2 N', the factors go into the kick (2d: 6*2, 3d: 6*6*2); generic orders: true N, N'
    const double scale = h->order != 4 ? 1.0 : (h->dim == 2 ? 12.0 : 72.0);
    P.gx = -c.dt * c.dx_inv / scale / (h->xpp ? 3.0 : 1.0); P.gy = -c.dt * c.dy_inv / scale; P.gz = -c.dt * c.dz_inv / scale;
    P.x_min = c.x_min; P.y_min = c.y_min; P.z_min = c.z_min;
    P.dx = c.dx; P.dy = c.dy; P.dz = c.dz;
    P.f0_kind = h->f0.kind;
    for (int i = 0; i < 4; ++i) P.f0p[i] = h->f0.p[i];
}

int launch_sample_f(Handle *h, size_t n, size_t npts, const double *d_pts, double *d_out, bool full, bool feet)
{
    SampleParams S{};
    S.feet = feet ? 1 : 0;
    fill_common(h, S.P);
    const nufi_b200_config3d &c = h->c;
    S.P.metrics = full ? 1 : 0;
    S.P.first_level = full ? (n == 0 ? -1 : static_cast<int>(n)) : static_cast<int>(n) - 1;
    S.Lx = c.Lx; S.Ly = c.Ly; S.Lz = c.Lz; S.Lx_inv = c.Lx_inv; S.Ly_inv = c.Ly_inv; S.Lz_inv = c.Lz_inv;
    S.dx_inv = c.dx_inv; S.dy_inv = c.dy_inv; S.dz_inv = c.dz_inv;
    S.pts = d_pts; S.out = d_out; S.npts = npts; S.with_first_half_kick = full ? 1 : 0;
    const unsigned blocks = static_cast<unsigned>(std::min<size_t>((npts + 127) / 128, 148 * 8));
    if (h->order != 4) NUFI_CUDA_CHECK(h, launch_sample_f_generic(h->order, h->dim, S, blocks, h->stream));
    else if (h->dim == 1) sample_f_kernel<1, false><<<blocks, 128, 0, h->stream>>>(S);
    else if (h->dim == 2) { if (h->xpp) sample_f_kernel<2, true><<<blocks, 128, 0, h->stream>>>(S); else sample_f_kernel<2, false><<<blocks, 128, 0, h->stream>>>(S); }
    else { if (h->xpp) sample_f_kernel<3, true><<<blocks, 128, 0, h->stream>>>(S); else sample_f_kernel<3, false><<<blocks, 128, 0, h->stream>>>(S); }
    NUFI_CUDA_CHECK(h, cudaGetLastError());
    h->launches += 1;
    return NUFI_B200_OK;
}

int launch_sample_field(Handle *h, const double *d_ref_level, int der, size_t npts, const double *d_pts, double *d_out)
{
    const nufi_b200_config3d &c = h->c;
    FieldSampleParams S{};
    S.dim = h->dim; S.order = h->order; S.Nx = static_cast<int>(c.Nx); S.Ny = static_cast<int>(c.Ny); S.Nz = static_cast<int>(c.Nz); S.der = der;
    S.x_min = c.x_min; S.y_min = c.y_min; S.z_min = c.z_min;
    S.Lx = c.Lx; S.Ly = c.Ly; S.Lz = c.Lz; S.Lx_inv = c.Lx_inv; S.Ly_inv = c.Ly_inv; S.Lz_inv = c.Lz_inv;
    S.dx_inv = c.dx_inv; S.dy_inv = c.dy_inv; S.dz_inv = c.dz_inv;
    S.level = d_ref_level; S.pts = d_pts; S.out = d_out; S.npts = npts;
    const unsigned blocks = static_cast<unsigned>(std::min<size_t>((npts + 127) / 128, 148 * 8));
    sample_field_kernel<<<blocks, 128, 0, h->stream>>>(S);
    NUFI_CUDA_CHECK(h, cudaGetLastError());
    h->launches += 1;
    return NUFI_B200_OK;
}

int launch_peer_noop(Handle *h)
{
    PeerPush X = h->px.push;
    size_t blocks = (h->n_nodes + 255) / 256;
    if (blocks > 148) blocks = 148;
    NUFI_CUDA_CHECK(h, launch_chained(h, peer_noop_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, X, h->n_nodes));
    h->launches += 1;
    return NUFI_B200_OK;
}

int launch_finish(Handle *h)
{
    if (!h->fin_pending) return NUFI_B200_OK;
    PeerPush X{}; // world 0 unless this is the large-grid leg of a multi-GPU step
    if (h->peer_push) X = h->px.push;
    NUFI_CUDA_CHECK(h, launch_chained(h, finish_rho_kernel, dim3(h->fin.n_tiles), dim3(256), 0, h->fin, X));
    NUFI_CUDA_CHECK(h, cudaGetLastError());
    h->fin_pending = false;
    h->launches += 1;
    return NUFI_B200_OK;
}

} // namespace nufi_b200
