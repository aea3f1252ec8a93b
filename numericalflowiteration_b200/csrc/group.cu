// group.cu -- several GPUs driven by one process: the partial rho of the per-device handles is exchanged through peer memory
// inside the path's own kernels (peer.cu; default when every device can map every other), or by an NCCL all-reduce.
//
// Replaces the host fan-in of the reference (cuda_kernel::download_rho's blocking copy + host add per device,
// nufi/cuda_kernel.cu:135-145, and the MPI_Allreduce on host buffers, bin/test_nufi_gpu_3d.cpp:158) by one
// ncclAllReduce(sum, double) per device on the device-resident vector, over NVLink/NVSwitch.  NCCL is loaded with
// dlopen on first use so the single-GPU library has no NCCL dependency.
#include "internal.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <new>

namespace nufi_b200
{

namespace
{

struct NcclApi
{
    void *lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};

NcclApi &nccl()
{
    static NcclApi api;
    if (api.lib || !api.err.empty()) return api;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
        api.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
        if (api.lib) break;
    }
    if (!api.lib) {
        api.err = std::string("cannot load NCCL: ") + dlerror();
        return api;
    }
    auto sym = [&](const char *n) {
        void *p = dlsym(api.lib, n);
        if (!p && api.err.empty()) api.err = std::string("NCCL symbol missing: ") + n;
        return p;
    };
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    return api;
}

} // namespace

struct Group
{
    std::vector<Handle *> hs;
    std::vector<ncclComm_t> comms;
    std::vector<size_t> q_edges; // contiguous near-equal split of [0, Nquad), first `rem` shares one longer
    bool peer_ok = false;        // every pair of devices can map each other's memory: exchange buffers are set up
    int exchange = 0;            // 0: peer-memory exchange fused into the kernels (when peer_ok), 1: NCCL all-reduce
    std::string peer_why;        // why peer_ok is false
    std::string err;
};

// direct peer access between all devices of the group + one exchange buffer per handle, mapped by plain pointers
static bool setup_peer_exchange(Group *g)
{
    const int parts = static_cast<int>(g->hs.size());
    if (parts > kMaxPeers) { g->peer_why = "more devices than kMaxPeers"; return false; }
    for (int i = 0; i < parts; ++i)
        for (int j = 0; j < parts; ++j) {
            if (i == j) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, g->hs[i]->device, g->hs[j]->device) != cudaSuccess || !can) {
                cudaGetLastError();
                g->peer_why = "device " + std::to_string(g->hs[i]->device) + " cannot access device " + std::to_string(g->hs[j]->device);
                return false;
            }
        }
    for (int i = 0; i < parts; ++i) {
        if (cudaSetDevice(g->hs[i]->device) != cudaSuccess) { g->peer_why = "cudaSetDevice failed"; return false; }
        for (int j = 0; j < parts; ++j) {
            if (i == j) continue;
            cudaError_t e = cudaDeviceEnablePeerAccess(g->hs[j]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                g->peer_why = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
                cudaGetLastError();
                return false;
            }
            cudaGetLastError();
        }
        if (peer_alloc(g->hs[i], parts) != NUFI_B200_OK) { g->peer_why = g->hs[i]->err; return false; }
    }
    for (int i = 0; i < parts; ++i) {
        PeerState &px = g->hs[i]->px;
        px.rank = i;
        px.ipc = false;
        for (int j = 0; j < parts; ++j) px.peer_xb[j] = g->hs[j]->px.xb;
    }
    return true;
}

static thread_local std::string g_group_create_error;

static int gfail(Group *g, int code, const std::string &msg)
{
    if (g) g->err = msg;
    else g_group_create_error = msg;
    return code;
}

static int ensure_nccl(Group *g)
{
    if (!g->comms.empty()) return NUFI_B200_OK;
    NcclApi &api = nccl();
    if (!api.err.empty()) return gfail(g, NUFI_B200_ERR_CUDA, api.err);
    std::vector<int> devs;
    for (Handle *h : g->hs) devs.push_back(h->device);
    g->comms.assign(g->hs.size(), nullptr);
    ncclResult_t r = api.CommInitAll(g->comms.data(), static_cast<int>(devs.size()), devs.data());
    if (r != ncclSuccess) {
        g->comms.clear();
        return gfail(g, NUFI_B200_ERR_CUDA, std::string("ncclCommInitAll: ") + api.GetErrorString(r));
    }
    return NUFI_B200_OK;
}

} // namespace nufi_b200

using namespace nufi_b200;

extern "C" {

int nufi_b200_group_create(nufi_b200_handle *const *handles, int n_handles, nufi_b200_group **out)
{
    if (!out) return gfail(nullptr, NUFI_B200_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (!handles || n_handles < 1) return gfail(nullptr, NUFI_B200_ERR_ARG, "need at least one handle");
    Group *g = new (std::nothrow) Group;
    if (!g) return gfail(nullptr, NUFI_B200_ERR_ALLOC, "out of host memory");
    std::vector<int> devs;
    for (int i = 0; i < n_handles; ++i) {
        Handle *h = reinterpret_cast<Handle *>(handles[i]);
        if (!h) { delete g; return gfail(nullptr, NUFI_B200_ERR_ARG, "NULL handle in group"); }
        if (i > 0 && (h->dim != g->hs[0]->dim || h->n_nodes != g->hs[0]->n_nodes || h->n_vel != g->hs[0]->n_vel || h->Nt != g->hs[0]->Nt)) {
            delete g;
            return gfail(nullptr, NUFI_B200_ERR_ARG, "handles of one group must share one configuration");
        }
        for (int d : devs)
            if (d == h->device) { delete g; return gfail(nullptr, NUFI_B200_ERR_ARG, "two handles of one group on the same device"); }
        g->hs.push_back(h);
        devs.push_back(h->device);
    }
    const size_t nq = g->hs[0]->n_nodes * g->hs[0]->n_vel, parts = static_cast<size_t>(n_handles);
    g->q_edges.assign(parts + 1, 0); // nufi/cuda_scheduler.hpp:88-111
    for (size_t i = 0; i < parts; ++i) g->q_edges[i + 1] = g->q_edges[i] + nq / parts + (i < nq % parts ? 1 : 0);
    if (n_handles > 1) g->peer_ok = setup_peer_exchange(g);
    if (n_handles > 1 && !g->peer_ok) { // NCCL only when the fused peer exchange is unavailable (or asked for: set_exchange)
        g->exchange = 1;
        int rc = ensure_nccl(g);
        if (rc) { std::string m = g->err; delete g; return gfail(nullptr, rc, m); }
    }
    *out = reinterpret_cast<nufi_b200_group *>(g);
    return NUFI_B200_OK;
}

void nufi_b200_group_destroy(nufi_b200_group *gg)
{
    Group *g = reinterpret_cast<Group *>(gg);
    if (!g) return;
    for (ncclComm_t c : g->comms)
        if (c) nccl().CommDestroy(c);
    delete g; // the handles stay owned by the caller
}

int nufi_b200_group_step(nufi_b200_group *gg, size_t n)
{
    Group *g = reinterpret_cast<Group *>(gg);
    if (!g) return gfail(nullptr, NUFI_B200_ERR_ARG, "group is NULL");
    const size_t parts = g->hs.size();
    if (parts == 1) {
        int rc = nufi_b200_step(reinterpret_cast<nufi_b200_handle *>(g->hs[0]), n);
        if (rc) g->err = g->hs[0]->err;
        return rc;
    }
    if (g->exchange == 0) { // fused: backtrace -> tail that adds this GPU's slots, pushes its sums to every GPU and polls the peers' (peer.cu)
        for (size_t i = 0; i < parts; ++i) {
            int rc = nufi_b200_peer_step(reinterpret_cast<nufi_b200_handle *>(g->hs[i]), n);
            if (rc) { g->err = g->hs[i]->err; return rc; }
        }
        return NUFI_B200_OK;
    }
    for (size_t i = 0; i < parts; ++i) {
        int rc = nufi_b200_compute_rho(reinterpret_cast<nufi_b200_handle *>(g->hs[i]), n, g->q_edges[i], g->q_edges[i + 1]);
        if (rc) { g->err = g->hs[i]->err; return rc; }
    }
    NcclApi &api = nccl();
    ncclResult_t r = api.GroupStart();
    for (size_t i = 0; i < parts && r == ncclSuccess; ++i) {
        Handle *h = g->hs[i];
        r = api.AllReduce(h->d_rho_partial, h->d_rho_partial, h->n_nodes, ncclDouble, ncclSum, g->comms[i], h->stream);
    }
    ncclResult_t r2 = api.GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) return gfail(g, NUFI_B200_ERR_CUDA, std::string("ncclAllReduce: ") + api.GetErrorString(r));
    for (size_t i = 0; i < parts; ++i) {
        Handle *h = g->hs[i];
        int rc = nufi_b200_field_tail_device(reinterpret_cast<nufi_b200_handle *>(h), n, h->d_rho_partial);
        if (rc) { g->err = h->err; return rc; }
    }
    return NUFI_B200_OK;
}

int nufi_b200_group_sync(nufi_b200_group *gg)
{
    Group *g = reinterpret_cast<Group *>(gg);
    if (!g) return gfail(nullptr, NUFI_B200_ERR_ARG, "group is NULL");
    for (Handle *h : g->hs) {
        int rc = nufi_b200_sync(reinterpret_cast<nufi_b200_handle *>(h));
        if (rc) { g->err = h->err; return rc; }
    }
    return NUFI_B200_OK;
}

int nufi_b200_group_set_exchange(nufi_b200_group *gg, int mode)
{
    Group *g = reinterpret_cast<Group *>(gg);
    if (!g) return gfail(nullptr, NUFI_B200_ERR_ARG, "group is NULL");
    if (mode != 0 && mode != 1) return gfail(g, NUFI_B200_ERR_ARG, "exchange mode must be 0 (peer memory) or 1 (NCCL)");
    if (g->hs.size() == 1) return NUFI_B200_OK;
    if (mode == 0 && !g->peer_ok) return gfail(g, NUFI_B200_ERR_ARG, "peer-memory exchange unavailable: " + g->peer_why);
    if (mode == 1) {
        int rc = ensure_nccl(g);
        if (rc) return rc;
    }
    int rc = nufi_b200_group_sync(gg);
    if (rc) return rc;
    g->exchange = mode;
    return NUFI_B200_OK;
}

const char *nufi_b200_group_exchange(const nufi_b200_group *gg)
{
    const Group *g = reinterpret_cast<const Group *>(gg);
    if (!g) return "none";
    if (g->hs.size() == 1) return "single";
    return g->exchange == 0 ? "peer-memory" : "nccl";
}

const char *nufi_b200_group_last_error(const nufi_b200_group *g)
{
    return g ? reinterpret_cast<const Group *>(g)->err.c_str() : g_group_create_error.c_str();
}

} // extern "C"
