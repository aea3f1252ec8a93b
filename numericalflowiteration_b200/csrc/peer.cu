// peer.cu -- the per-step exchange of the partial rho between the GPUs of one NVSwitch box, fused into the path's own kernels.
//
// Replaces the reference's fan-in (cuda_kernel::download_rho's blocking copy + host add per device, nufi/cuda_kernel.cu:135-145,
// nufi/cuda_scheduler.hpp:113-118, and MPI_Allreduce on host buffers, bin/test_nufi_gpu_3d.cpp:158) -- and the NCCL all-reduce
// this library used first -- by direct stores into peer memory:
//   * the backtrace kernel (epilogue mode 3, backtrace_kernel.cuh): every CTA stores each per-(CTA, tile) slot it finishes into
//     the exchange buffer of EVERY GPU (NVLink stores) and, when it is done, fences once and adds to a per-(rank, parity) counter
//     on every GPU.  There is no reduction and no serial section in the kernel: it ends when its CTAs end;
//   * the field tail (tail_small_kernel / peer_gather_kernel) acquires the `world` counters and adds the slots of all ranks in a
//     fixed order, so all replicas compute bit-identical rho, phi and histories with no collective call, no extra launch for
//     the reduction and no host synchronisation.  Two parities make the buffers safe to reuse: a rank can only write epoch e+2
//     after its tail of e+1 saw every peer's push of e+1, which those peers issued after their tail of e had read epoch e.
//     (r01: the kernel's LAST CTA reduced all tiles and pushed the reduced rho -- a serial epilogue of ~9 us on a 150 us step.)
// Mapping of the peers' buffers: one process driving all GPUs (nufi_b200_group_*) enables direct peer access; one process per
// GPU (torchrun) exchanges cudaIpcMemHandle_t through the host layer (nufi_b200_peer_export / _attach).
#include "internal.cuh"

#include <cstring>

namespace nufi_b200
{

namespace
{

// exchange buffer -> d_rho_full for grids too large for the single-CTA tail (the cuFFT path reads rho from memory)
__global__ void __launch_bounds__(256) peer_gather_kernel(const __grid_constant__ PeerRecv X, size_t n_nodes)
{
    pdl_wait();
    peer_wait_all(X);
    __syncthreads();
    for (size_t l = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; l < n_nodes; l += static_cast<size_t>(gridDim.x) * blockDim.x)
        X.rho_full[l] = peer_rho(X, l);
}

size_t slot_offset(const Handle *h, int parity, int rank)
{
    return kPeerCounterBytes + kPeerHeaderBytes + (static_cast<size_t>(parity) * h->px.world + rank) * h->px.slot_cap * sizeof(double);
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
    // slot space per (parity, rank): grid x Tmax x 32 doubles with Tmax = (rpc-1)/rpt + 2 <= n_tiles/grid + 3 (tiles of 32 nodes)
    const size_t n_tiles = (h->n_nodes + 31) / 32;
    px.slot_cap = static_cast<size_t>(h->sm_count) * (n_tiles / h->sm_count + 3) * 32;
    px.xb_bytes = kPeerCounterBytes + kPeerHeaderBytes + 2 * static_cast<size_t>(world) * px.slot_cap * sizeof(double);
    NUFI_CUDA_CHECK(h, cudaMalloc(&px.xb, px.xb_bytes));
    NUFI_CUDA_CHECK(h, cudaMemset(px.xb, 0, px.xb_bytes));
    NUFI_CUDA_CHECK(h, cudaMalloc(&px.d_status, sizeof(int)));
    NUFI_CUDA_CHECK(h, cudaMemset(px.d_status, 0, sizeof(int)));
    NUFI_CUDA_CHECK(h, cudaDeviceSynchronize()); // the zeroed counters are in place before any peer learns the address
    return NUFI_B200_OK;
}

// Next epoch: pointers for the sender, counter target and slot regions for the receiver.
int peer_prepare_step(Handle *h)
{
    PeerState &px = h->px;
    if (px.world < 1 || px.rank < 0) return fail(h, NUFI_B200_ERR_ARG, "peer exchange not attached (nufi_b200_peer_attach / group)");
    const unsigned long long e = ++px.epoch;
    const int parity = static_cast<int>(e & 1);
    PeerPush P{};
    P.world = px.world;
    for (int p = 0; p < px.world; ++p) {
        P.slots[p] = reinterpret_cast<double *>(px.peer_xb[p] + slot_offset(h, parity, px.rank));
        P.counter[p] = reinterpret_cast<unsigned long long *>(px.peer_xb[p]) + parity * kMaxPeers + px.rank;
        P.header[p] = reinterpret_cast<PeerHeader *>(px.peer_xb[p] + kPeerCounterBytes) + parity * kMaxPeers + px.rank;
    }
    PeerRecv R{};
    R.world = px.world;
    R.counters = reinterpret_cast<const unsigned long long *>(px.xb) + parity * kMaxPeers;
    R.target = ((e + (e & 1)) / 2) * kPeerUnit; // epochs of this parity so far, kPeerUnit from every rank in each
    R.headers = reinterpret_cast<const PeerHeader *>(px.xb + kPeerCounterBytes) + parity * kMaxPeers;
    R.slots = reinterpret_cast<const double *>(px.xb + slot_offset(h, parity, 0));
    R.slot_cap = px.slot_cap;
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
    NUFI_CUDA_CHECK(h, launch_chained(h, peer_gather_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, h->px.recv, h->n_nodes));
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
        rc = nufi_b200_compute_rho(h, n, 0, hh->n_nodes * hh->n_vel); // backtrace whose CTAs push their slots to every GPU
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
    return NUFI_B200_OK;
}

} // extern "C"
