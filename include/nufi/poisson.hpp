// nufi/poisson.hpp -- nufi::dim{1,2,3}::poisson<real> with the reference's interface (nufi/poisson.hpp:33-60, 100-128,
// 164-192): construct from a config, `real solve(real *data)` overwrites rho at the nodes with phi and returns the
// electric energy, `alignment` for the caller's aligned_alloc.  The solve runs on the device (cuFFT D2Z -> symbol ->
// Z2D) instead of two FFTW DHTs (nufi/poisson.cpp:66-89, 190-219, 328-362).
#ifndef NUFI_B200_NUFI_POISSON_HPP
#define NUFI_B200_NUFI_POISSON_HPP

#include "cuda_scheduler.hpp"

namespace nufi
{

#define NUFI_B200_DEFINE_POISSON(DIM)                                                                                   \
    namespace DIM                                                                                                       \
    {                                                                                                                   \
    template <typename real> class poisson : detail::require_double<real>                                               \
    {                                                                                                                   \
    public:                                                                                                             \
        static constexpr size_t alignment{64};                                                                          \
        poisson() = delete;                                                                                             \
        explicit poisson(const config_t<real> &param) : kern{without_history(param), -1}, param_{param} {}              \
        config_t<real> conf() const noexcept { return param_; }                                                         \
        /* nufi/poisson.hpp:49, poisson.cpp:51-64: re-plan for a new grid */                                            \
        void conf(const config_t<real> &new_param)                                                                      \
        {                                                                                                               \
            kern = detail::kernel_impl<config_t<real>, 4>{without_history(new_param), -1};                              \
            param_ = new_param;                                                                                         \
        }                                                                                                               \
        real solve(real *data)                                                                                          \
        {                                                                                                               \
            double e = 0;                                                                                               \
            cuda::check(nufi_b200_poisson_solve(kern.handle(), data, &e), nufi_b200_last_error(kern.handle()));         \
            return e;                                                                                                   \
        }                                                                                                               \
                                                                                                                        \
    private:                                                                                                            \
        static config_t<real> without_history(config_t<real> c) { c.Nt = 0; return c; }                                 \
        detail::kernel_impl<config_t<real>, 4> kern;                                                                    \
        config_t<real> param_;                                                                                          \
    };                                                                                                                  \
    }

NUFI_B200_DEFINE_POISSON(dim1)
NUFI_B200_DEFINE_POISSON(dim2)
NUFI_B200_DEFINE_POISSON(dim3)
#undef NUFI_B200_DEFINE_POISSON

} // namespace nufi

#endif
