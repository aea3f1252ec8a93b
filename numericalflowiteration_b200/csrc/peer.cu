// peer.cu -- the per-step exchange of the partial rho between the GPUs of one NVSwitch box, fused into the path's own kernels.
//
// Replaces the reference's fan-in (cuda_kernel::download_rho's blocking copy + host add per device, nufi/cuda_kernel.cu:135-145,
// nufi/cuda_scheduler.hpp:113-118, and MPI_Allreduce on host buffers, bin/test_nufi_gpu_3d.cpp:158) -- and the NCCL all-reduce
// this library used first -- by direct stores into peer memory (layout and wire encoding: internal.cuh):
//   * the backtrace kernel is the single-GPU fused step's: it writes its per-(CTA, tile) slots into LOCAL memory as self-validating
//     words (every 8-byte half carries the step's epoch beside 32 bits of payload) and ends when its CTAs end -- no reduction, no
//     fence, no flag, no serial section;
//   * the field tail (tail_small_kernel), resident beside that kernel since its start, adds the local slots in the fixed order as
//     they land, STORES the per-node sums into the exchange buffer of every other GPU (NVLink stores of the same kind of words),
//     and adds all ranks' sums in rank order, polling every word it loads until it carries this step's epoch -- so all replicas
//     compute bit-identical rho, phi and histories with no collective call, no extra launch and no host synchronisation.  Grids too
//     large for the one-CTA tail: finish_rho_kernel (one block per tile) pushes, peer_gather_kernel polls.
//     History of this exchange (profiles/r02_peer_exchange.md): r01 = the kernel's LAST CTA reduced all tiles serially and pushed
//     the result behind a system-scope fence (3.8 us alone) and a flag, ~9 us on a 150 us step; r02 v2/v3 = every CTA pushed its
//     raw slots and the one-CTA tail added world x (CTAs x tiles) of them -- no sender epilogue, but a tail that grew with the
//     GPU count (33 us at 8 GPUs); v4 = the last CTA to finish a tile reduced and pushed it (fence + atomic + dependent loads at
//     the end of every CTA: the kernel 10 us longer).
// Mapping of the peers' buffers: one process driving all GPUs (nufi_b200_group_*) enables direct peer access; one process per
// GPU (torchrun) exchanges cudaIpcMemHandle_t through the host layer (nufi_b200_peer_export / _attach).
#include "internal.cuh"

#include <cstring>

namespace nufi_b200
{

namespace
{

// exchange buffer -> d_rho_full for grids too large for the single-CTA tail (the cuFFT path reads rho from memory).  No
// pdl_wait: the words validate themselves, and nothing else the kernels ahead on the stream write is read here.
__global__ void __launch_bounds__(256) peer_gather_kernel(const __grid_constant__ PeerRecv X)
{
    pdl_trigger();
    for (size_t l = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; l < X.n_nodes; l += static_cast<size_t>(gridDim.x) * blockDim.x)
        X.rho_full[l] = fma(-X.dV, peer_rank_sum(X.rho + l, X.n_nodes, X.world, X.flag, X.status, -1, 0.0), 1.0);
}

size_t region_offset(const Handle *h, int parity, int rank)
{
    return (static_cast<size_t>(parity) * h->px.world + rank) * h->n_nodes * sizeof(uint4);
}

} // namespace

void peer_free(Handle *h)
{
    PeerState &px = h->px;
    if (px.ipc)
        for (int p = 0; p < px.world; ++p)
            if (p != px.rank && px.peer_xb[p]) cudaIpcCloseMemHandle(px.peer_xb[p]);
    cudaFree(px.xb);
    cudaFree(px.d_status);
    px = PeerState{};
}

int peer_alloc(Handle *h, int world)
{
    if (world < 1 || world > kMaxPeers) return fail(h, NUFI_B200_ERR_ARG, "peer exchange: world size must be 1.." + std::to_string(kMaxPeers));
    peer_free(h);
    PeerState &px = h->px;
    px.world = world;
    px.xb_bytes = 2 * static_cast<size_t>(world) * h->n_nodes * sizeof(uint4);
    NUFI_CUDA_CHECK(h, cudaMalloc(&px.xb, px.xb_bytes));
    NUFI_CUDA_CHECK(h, cudaMemset(px.xb, 0, px.xb_bytes));
    NUFI_CUDA_CHECK(h, cudaMalloc(&px.d_status, sizeof(int)));
    NUFI_CUDA_CHECK(h, cudaMemset(px.d_status, 0, sizeof(int)));
    NUFI_CUDA_CHECK(h, cudaDeviceSynchronize()); // the zeroed words (epoch 0 = "nothing yet") are in place before any peer learns the address
    return NUFI_B200_OK;
}

// Next epoch: pointers and the epoch word for the sender, the local regions for the receiver.
int peer_prepare_step(Handle *h)
{
    PeerState &px = h->px;
    if (px.world < 1 || px.rank < 0) return fail(h, NUFI_B200_ERR_ARG, "peer exchange not attached (nufi_b200_peer_attach / group)");
    const unsigned long long e = ++px.epoch;
    const int parity = static_cast<int>(e & 1);
    PeerPush P{};
    P.world = px.world;
    P.rank = px.rank;
    P.flag = static_cast<unsigned int>(e & 0xffffffffull);
    if (P.flag == 0) P.flag = 0x80000000u; // never the value the zeroed buffers hold (after 2^32 steps)
    for (int p = 0; p < px.world; ++p) P.rho[p] = reinterpret_cast<uint4 *>(px.peer_xb[p] + region_offset(h, parity, px.rank));
    PeerRecv R{};
    R.world = px.world;
    R.flag = P.flag;
    R.rho = reinterpret_cast<const uint4 *>(px.xb + region_offset(h, parity, 0));
    R.n_nodes = h->n_nodes;
    const nufi_b200_config3d &c = h->c; // rho.hpp:136-145, 291-307, 441-459: du recomputed from the bounds
    R.dV = (c.u_max - c.u_min) / c.Nu;
    if (h->dim >= 2) R.dV *= (c.v_max - c.v_min) / c.Nv;
    if (h->dim >= 3) R.dV *= (c.w_max - c.w_min) / c.Nw;
    R.rho_full = h->d_rho_full;
    R.status = px.d_status;
    px.push = P;
    px.recv = R;
    return NUFI_B200_OK;
}

int launch_peer_gather(Handle *h)
{
    size_t blocks = (h->n_nodes + 255) / 256;
    if (blocks > 592) blocks = 592;
    NUFI_CUDA_CHECK(h, launch_chained(h, peer_gather_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, h->px.recv));
    h->launches += 1;
    return NUFI_B200_OK;
}

} // namespace nufi_b200

using namespace nufi_b200;

static inline Handle *HH(nufi_b200_handle *h) { return reinterpret_cast<Handle *>(h); }

#define PEER_ENTER(h)                                                    \
    Handle *hh = HH(h);                                                  \
    if (!hh) return fail(nullptr, NUFI_B200_ERR_ARG, "handle is NULL");  \
    NUFI_CUDA_CHECK(hh, cudaSetDevice(hh->device))

extern "C" {

int nufi_b200_peer_export(nufi_b200_handle *h, int world, void *ipc_handle)
{
    PEER_ENTER(h);
    if (!ipc_handle) return fail(hh, NUFI_B200_ERR_ARG, "ipc_handle is NULL");
    static_assert(sizeof(cudaIpcMemHandle_t) == NUFI_B200_PEER_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
    int rc = peer_alloc(hh, world);
    if (rc) return rc;
    cudaIpcMemHandle_t mh;
    NUFI_CUDA_CHECK(hh, cudaIpcGetMemHandle(&mh, hh->px.xb));
    std::memcpy(ipc_handle, &mh, sizeof(mh));
    return NUFI_B200_OK;
}

int nufi_b200_peer_attach(nufi_b200_handle *h, int rank, int world, const void *ipc_handles)
{
    PEER_ENTER(h);
    PeerState &px = hh->px;
    if (!px.xb || px.world != world) return fail(hh, NUFI_B200_ERR_ARG, "peer_attach: call nufi_b200_peer_export with the same world size first");
    if (rank < 0 || rank >= world) return fail(hh, NUFI_B200_ERR_ARG, "peer_attach: rank out of range");
    if (!ipc_handles && world > 1) return fail(hh, NUFI_B200_ERR_ARG, "ipc_handles is NULL");
    px.rank = rank;
    px.ipc = true;
    for (int p = 0; p < world; ++p) {
        if (p == rank) { px.peer_xb[p] = px.xb; continue; }
        cudaIpcMemHandle_t mh;
        std::memcpy(&mh, static_cast<const unsigned char *>(ipc_handles) + static_cast<size_t>(p) * sizeof(mh), sizeof(mh));
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, mh, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int q = 0; q < p; ++q)
                if (q != rank && px.peer_xb[q]) { cudaIpcCloseMemHandle(px.peer_xb[q]); px.peer_xb[q] = nullptr; }
            px.rank = -1;
            return fail(hh, NUFI_B200_ERR_CUDA, std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(p) + "): " + cudaGetErrorString(e));
        }
        px.peer_xb[p] = static_cast<unsigned char *>(ptr);
    }
    return NUFI_B200_OK;
}

int nufi_b200_peer_detach(nufi_b200_handle *h)
{
    PEER_ENTER(h);
    NUFI_CUDA_CHECK(hh, cudaStreamSynchronize(hh->stream));
    peer_free(hh);
    return NUFI_B200_OK;
}

int nufi_b200_peer_step(nufi_b200_handle *h, size_t n)
{
    PEER_ENTER(h);
    if (n > hh->Nt) return fail(hh, NUFI_B200_ERR_RANGE, "Time-step out of range.");
    for (size_t m = 0; m < n; ++m) // validate BEFORE the epoch advances: a rank that bails out later would stall its peers
        if (!hh->level_valid[m])
            return fail(hh, NUFI_B200_ERR_RANGE, "peer_step: history level " + std::to_string(m) + " was never uploaded or computed");
    int rc = peer_prepare_step(hh);
    if (rc) return rc;
    const PeerState &px = hh->px;
    // Sharding of the fused multi-GPU step: rank r takes the velocity nodes r, r+world, r+2 world, ... of every spatial node.
    // (compute_rho keeps the reference's contiguous flat-q split, nufi/cuda_scheduler.hpp:88-111; that split hands each GPU
    // a different region of phase space, and regions differ in cost -- trapped orbits replay shared-memory loads -- so the
    // step would wait for the slowest GPU.  The interleaved split gives every GPU a statistically identical sample.)
    if (static_cast<size_t>(px.rank) < hh->n_vel) {
        hh->peer_push = true;
        hh->vstride = static_cast<unsigned long long>(px.world);
        hh->voff = static_cast<unsigned long long>(px.rank);
        rc = nufi_b200_compute_rho(h, n, 0, hh->n_nodes * hh->n_vel); // backtrace whose CTAs push the finished tiles to every GPU
        hh->peer_push = false;
        hh->vstride = 1;
        hh->voff = 0;
    } else {
        rc = launch_peer_noop(hh);
    }
    if (rc) return rc;
    return tail_run(hh, n, nullptr, /*from_peer=*/true);
}

int nufi_b200_peer_status(nufi_b200_handle *h, int *timed_out)
{
    PEER_ENTER(h);
    if (!timed_out) return fail(hh, NUFI_B200_ERR_ARG, "timed_out is NULL");
    *timed_out = 0;
    if (!hh->px.d_status) return NUFI_B200_OK;
    NUFI_CUDA_CHECK(hh, cudaStreamSynchronize(hh->stream));
    NUFI_CUDA_CHECK(hh, cudaMemcpy(timed_out, hh->px.d_status, sizeof(int), cudaMemcpyDeviceToHost));
    int slots = 0; // the tail's polling loads of this GPU's own slots
    NUFI_CUDA_CHECK(hh, cudaMemcpy(&slots, hh->d_ll_status, sizeof(int), cudaMemcpyDeviceToHost));
    *timed_out |= slots;
    return NUFI_B200_OK;
}

} // extern "C"
