// nufi/rho.hpp -- eval_rho with the reference's signature (nufi/rho.hpp:133-134, 283-284, 428-429), computed on the GPU.
//
// The reference's CPU drivers call  rho[l] = eval_rho<real,order>(n, l, coeffs, conf)  for every spatial node l inside an
// OpenMP loop (bin/test_nufi_cpu_1d.cpp:65-70, _2d.cpp:68-72, _3d.cpp:68-72).  Here the FIRST call for a step n runs the
// whole sweep on the device in one launch (after pushing the levels of `coeffs` the device does not hold yet) and every
// call, from any thread, returns its node from that result -- so the driver loop compiles and runs unchanged.
// eval_rho_all is the same thing without the per-node detour.
#ifndef NUFI_B200_NUFI_RHO_HPP
#define NUFI_B200_NUFI_RHO_HPP

#include "device_context.hpp"

namespace nufi
{

#define NUFI_B200_DEFINE_RHO(DIM)                                                                                       \
    namespace DIM                                                                                                       \
    {                                                                                                                   \
    template <typename real, size_t order> real eval_rho(size_t n, size_t l, const real *coeffs, const config_t<real> &conf) \
    {                                                                                                                   \
        static_assert(std::is_same<real, double>::value, "libnufi_b200 computes in FP64");                              \
        return detail::context<config_t<real>, order>(conf)->rho_at(n, l, coeffs);                                      \
    }                                                                                                                   \
    template <typename real, size_t order> void eval_rho_all(size_t n, real *rho, const real *coeffs, const config_t<real> &conf) \
    {                                                                                                                   \
        static_assert(std::is_same<real, double>::value, "libnufi_b200 computes in FP64");                              \
        detail::context<config_t<real>, order>(conf)->rho_all(n, coeffs, rho);                                          \
    }                                                                                                                   \
    }

// eval_phase_flow (nufi/rho.hpp:98-131; dim1 in the reference, bin/test_nufi_gpu_1d.cpp:155,190,339): the reference's
// signature (one point per call, x and u updated in place -- one device call per point) and a batched form for plot grids.
namespace dim1
{
template <typename real, size_t order> void eval_phase_flow(size_t n, real &x, real &u, const real *coeffs, const config_t<real> &conf)
{
    static_assert(std::is_same<real, double>::value, "libnufi_b200 computes in FP64");
    double in[2] = {x, u}, out[2];
    detail::context<config_t<real>, order>(conf)->phase_flow(n, 1, in, out, coeffs);
    x = out[0];
    u = out[1];
}
template <typename real, size_t order>
void eval_phase_flow_all(size_t n, size_t npts, const real *points /*[npts][2]: x, u*/, real *feet, const real *coeffs, const config_t<real> &conf)
{
    static_assert(std::is_same<real, double>::value, "libnufi_b200 computes in FP64");
    detail::context<config_t<real>, order>(conf)->phase_flow(n, npts, points, feet, coeffs);
}
} // namespace dim1

NUFI_B200_DEFINE_RHO(dim1)
NUFI_B200_DEFINE_RHO(dim2)
NUFI_B200_DEFINE_RHO(dim3)
#undef NUFI_B200_DEFINE_RHO

} // namespace nufi

#endif
