// nufi/device_context.hpp -- a cached device handle per configuration, for the reference's FREE functions and
// handle-less classes (eval_rho, interpolate, poisson<real>), which take only (coeffs, conf).  Not in the reference:
// there these run on the host and need no state.
#ifndef NUFI_B200_NUFI_DEVICE_CONTEXT_HPP
#define NUFI_B200_NUFI_DEVICE_CONTEXT_HPP

#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

#include "cuda_scheduler.hpp"

namespace nufi
{

namespace detail
{

// Device-side mirror of a host coefficient history: remembers a sampled fingerprint of every level it uploaded and
// re-uploads a level only when the host copy changed (in the time loop: exactly the one new level per step).
template <typename Conf, size_t order> class mirror
{
    using tr = conf_traits<Conf>;

public:
    explicit mirror(const Conf &c) : conf{c}, kern{c, -1}, stride{tr::stride_t(c, order)}, prints(c.Nt + 1, 0) {}

    bool same_config(const Conf &c) const { return std::memcmp(&c, &conf, sizeof(Conf)) == 0; }

    // make device levels [0, n) equal to coeffs[0 .. n*stride_t)
    void sync_levels(size_t n, const double *coeffs)
    {
        for (size_t m = 0; m < n && m < prints.size(); ++m) {
            const unsigned long long fp = fingerprint(coeffs + m * stride);
            if (fp != prints[m]) {
                kern.upload_phi(m, coeffs);
                prints[m] = fp;
            }
        }
    }

    // rho of step n for all nodes (CPU convention, with the leading 1); cached per (n, history fingerprint)
    const std::vector<double> &rho(size_t n, const double *coeffs)
    {
        std::lock_guard<std::mutex> lock(mtx);
        const unsigned long long key = n == 0 ? 1 : fingerprint(coeffs + (n - 1) * stride) ^ (0x9e3779b97f4a7c15ull * (n + 1));
        if (!(have && cached_n == n && cached_key == key && cached_ptr == coeffs)) {
            sync_levels(n, coeffs);
            cache.resize(tr::nodes(conf));
            kern.eval_rho_all(n, cache.data());
            have = true; cached_n = n; cached_key = key; cached_ptr = coeffs;
        }
        return cache;
    }

    // feet (x.., v..) at t = 0 of the characteristics through `npts` phase-space points at t_n (needs levels 0..n of coeffs)
    void phase_flow(size_t n, size_t npts, const double *points, double *feet, const double *coeffs)
    {
        std::lock_guard<std::mutex> lock(mtx);
        if (n > 1) sync_levels(n + 1, coeffs);
        kern.eval_phase_flow(n, npts, points, feet);
    }

    kernel_impl<Conf, order> &kernel() { return kern; }

private:
    unsigned long long fingerprint(const double *level) const
    {
        unsigned long long h = 0xcbf29ce484222325ull;
        const size_t step = stride > 64 ? stride / 61 : 1;
        for (size_t i = 0; i < stride; i += step) {
            unsigned long long b;
            std::memcpy(&b, level + i, 8);
            h = (h ^ b) * 0x100000001b3ull;
        }
        return h | 1ull; // never 0 (= "not uploaded")
    }

    Conf conf;
    kernel_impl<Conf, order> kern;
    size_t stride;
    std::vector<unsigned long long> prints;
    std::mutex mtx;
    std::vector<double> cache;
    bool have = false;
    size_t cached_n = 0;
    unsigned long long cached_key = 0;
    const double *cached_ptr = nullptr;
};

// one mirror per configuration type; re-created when a different configuration shows up
template <typename Conf, size_t order> mirror<Conf, order> &context(const Conf &conf)
{
    static std::mutex mtx;
    static std::unique_ptr<mirror<Conf, order>> ctx;
    std::lock_guard<std::mutex> lock(mtx);
    if (!ctx || !ctx->same_config(conf)) ctx.reset(new mirror<Conf, order>(conf));
    return *ctx;
}

} // namespace detail

} // namespace nufi

#endif
