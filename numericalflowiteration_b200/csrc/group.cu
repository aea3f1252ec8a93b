// group.cu -- several GPUs driven by one process: NCCL all-reduce of the partial rho between the per-device handles.
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
    std::string err;
};

static thread_local std::string g_group_create_error;

static int gfail(Group *g, int code, const std::string &msg)
{
    if (g) g->err = msg;
    else g_group_create_error = msg;
    return code;
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
    if (n_handles > 1) {
        NcclApi &api = nccl();
        if (!api.err.empty()) { std::string m = api.err; delete g; return gfail(nullptr, NUFI_B200_ERR_CUDA, m); }
        g->comms.assign(parts, nullptr);
        ncclResult_t r = api.CommInitAll(g->comms.data(), n_handles, devs.data());
        if (r != ncclSuccess) {
            std::string m = std::string("ncclCommInitAll: ") + api.GetErrorString(r);
            delete g;
            return gfail(nullptr, NUFI_B200_ERR_CUDA, m);
        }
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

const char *nufi_b200_group_last_error(const nufi_b200_group *g)
{
    return g ? reinterpret_cast<const Group *>(g)->err.c_str() : g_group_create_error.c_str();
}

} // extern "C"
