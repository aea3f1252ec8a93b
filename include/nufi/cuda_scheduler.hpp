// nufi/cuda_scheduler.hpp -- nufi::dim{1,2,3}::cuda_kernel / cuda_scheduler over libnufi_b200.so.
//
// Drop-in for the reference's nufi/cuda_kernel.hpp:33-56, 77-98, 119-140 and nufi/cuda_scheduler.hpp:33-164, 173-279,
// 286-392: same class names, constructors, the five methods (compute_rho, download_rho, upload_phi, compute_metrics,
// download_metrics) with the same argument meaning, the same split of [q_begin,q_end) over the cards
// (cuda_scheduler.hpp:88-111), the same accumulate-on-download convention and the same exception types.  New, beyond
// the reference: step(n) -- backtrace + NCCL all-reduce + Poisson + interpolation entirely on the devices, no host
// round trip (what bin/test_nufi_gpu_3d.cpp:154-162 does through the host) -- plus electric_energy(), download_phi().
//
// real = double; order = 3..8 as the reference instantiates (order 4, what every reference driver runs, is the specialised
// fast path; the others run the generic Cox-de Boor kernel).
#ifndef NUFI_B200_NUFI_CUDA_SCHEDULER_HPP
#define NUFI_B200_NUFI_CUDA_SCHEDULER_HPP

#include <cstddef>
#include <cstring>
#include <type_traits>
#include <utility>
#include <vector>

#include "config.hpp"
#include "cuda_runtime.hpp"

namespace nufi
{

namespace detail
{

template <typename Conf> struct conf_traits;
template <> struct conf_traits<dim1::config_t<double>>
{
    static constexpr int dim = 1;
    using pod = nufi_b200_config1d;
    static int create(const pod *c, int order, const nufi_b200_f0 *f, int dev, nufi_b200_handle **h) { return nufi_b200_create_1d(c, order, f, dev, h); }
    static size_t nodes(const dim1::config_t<double> &c) { return c.Nx; }
    static size_t quad(const dim1::config_t<double> &c) { return c.Nx * c.Nu; }
    static size_t stride_t(const dim1::config_t<double> &c, size_t o) { return c.Nx + o - 1; }
};
template <> struct conf_traits<dim2::config_t<double>>
{
    static constexpr int dim = 2;
    using pod = nufi_b200_config2d;
    static int create(const pod *c, int order, const nufi_b200_f0 *f, int dev, nufi_b200_handle **h) { return nufi_b200_create_2d(c, order, f, dev, h); }
    static size_t nodes(const dim2::config_t<double> &c) { return c.Nx * c.Ny; }
    static size_t quad(const dim2::config_t<double> &c) { return c.Nx * c.Ny * c.Nu * c.Nv; }
    static size_t stride_t(const dim2::config_t<double> &c, size_t o) { return (c.Nx + o - 1) * (c.Ny + o - 1); }
};
template <> struct conf_traits<dim3::config_t<double>>
{
    static constexpr int dim = 3;
    using pod = nufi_b200_config3d;
    static int create(const pod *c, int order, const nufi_b200_f0 *f, int dev, nufi_b200_handle **h) { return nufi_b200_create_3d(c, order, f, dev, h); }
    static size_t nodes(const dim3::config_t<double> &c) { return c.Nx * c.Ny * c.Nz; }
    static size_t quad(const dim3::config_t<double> &c) { return c.Nx * c.Ny * c.Nz * c.Nu * c.Nv * c.Nw; }
    static size_t stride_t(const dim3::config_t<double> &c, size_t o) { return (c.Nx + o - 1) * (c.Ny + o - 1) * (c.Nz + o - 1); }
};

// One device.  Owns a nufi_b200_handle (move-only, like the reference's cuda::autoptr members).
template <typename Conf, size_t order> class kernel_impl
{
    using tr = conf_traits<Conf>;
    static_assert(sizeof(Conf) == sizeof(typename tr::pod), "config_t<double> must be layout-identical to the C ABI struct");
    static_assert(order >= 3 && order <= 8, "libnufi_b200 implements the spline orders the reference instantiates: 3..8");

public:
    kernel_impl(const Conf &conf, int dev) : conf_{conf}
    {
        const int rc = tr::create(reinterpret_cast<const typename tr::pod *>(&conf_), static_cast<int>(order), &Conf::f0_sel, dev, &h_);
        cuda::check(rc, nufi_b200_last_error(nullptr));
    }
    kernel_impl(const kernel_impl &) = delete;
    kernel_impl &operator=(const kernel_impl &) = delete;
    kernel_impl(kernel_impl &&rhs) noexcept : conf_{rhs.conf_}, h_{std::exchange(rhs.h_, nullptr)} {}
    kernel_impl &operator=(kernel_impl &&rhs) noexcept
    {
        if (this != &rhs) { reset(); conf_ = rhs.conf_; h_ = std::exchange(rhs.h_, nullptr); }
        return *this;
    }
    ~kernel_impl() { reset(); }

    void compute_rho(size_t n, size_t q_min, size_t q_max) { ck(nufi_b200_compute_rho(h_, n, q_min, q_max)); }
    void download_rho(double *rho) { ck(nufi_b200_download_rho(h_, rho)); }
    void upload_phi(size_t n, const double *coeffs) { ck(nufi_b200_upload_phi(h_, n, coeffs)); }
    void compute_metrics(size_t n, size_t q_min, size_t q_max) { ck(nufi_b200_compute_metrics(h_, n, q_min, q_max)); }
    void download_metrics(double *metrics) { ck(nufi_b200_download_metrics(h_, metrics)); }
    // dim 1: metrics on a grid of their own (the reference's cuda_kernel(conf, conf_metrics, dev), nufi/cuda_kernel.cu:97-110)
    void set_metrics_grid(const Conf &conf_metrics)
    {
        static_assert(tr::dim == 1, "a separate metrics grid exists for dim1 only");
        ck(nufi_b200_set_metrics_grid_1d(h_, reinterpret_cast<const nufi_b200_config1d *>(&conf_metrics)));
    }

    // beyond the reference
    void step(size_t n) { ck(nufi_b200_step(h_, n)); }
    // the reference drivers' loop body in one call, history on the host: returns the electric energy of step n
    double step_host(size_t n, double *coeffs, double *rho = nullptr) { double e = 0; ck(nufi_b200_step_host(h_, n, coeffs, rho, &e, 0)); return e; }
    void eval_rho_all(size_t n, double *rho) { ck(nufi_b200_eval_rho_all(h_, n, rho)); }
    double solve_interpolate(size_t n) { double e = 0; ck(nufi_b200_solve_interpolate(h_, n, &e)); return e; }
    double electric_energy(size_t n) { double e = 0; ck(nufi_b200_download_energy(h_, n, n + 1, &e)); return e; }
    void download_phi(size_t n, double *level) { ck(nufi_b200_download_phi(h_, n, level)); }
    void eval_phase_flow(size_t n, size_t npts, const double *points, double *feet) { ck(nufi_b200_eval_phase_flow(h_, n, npts, points, feet)); }
    void sync() { ck(nufi_b200_sync(h_)); }
    nufi_b200_handle *handle() const noexcept { return h_; }
    const Conf &config() const noexcept { return conf_; }

private:
    void ck(int rc) { cuda::check(rc, nufi_b200_last_error(h_)); }
    void reset() noexcept { if (h_) nufi_b200_destroy(h_); h_ = nullptr; }
    Conf conf_;
    nufi_b200_handle *h_ = nullptr;
};

// All visible devices, one kernel per device that could be set up (cuda_scheduler.hpp:43-63).
template <typename Conf, size_t order> class scheduler_impl
{
    using tr = conf_traits<Conf>;

public:
    scheduler_impl() = delete;
    scheduler_impl(const scheduler_impl &) = delete;
    scheduler_impl &operator=(const scheduler_impl &) = delete;
    scheduler_impl(scheduler_impl &&rhs) noexcept : conf{rhs.conf}, kernels{std::move(rhs.kernels)}, group{std::exchange(rhs.group, nullptr)} {}
    scheduler_impl &operator=(scheduler_impl &&rhs) noexcept
    {
        if (this != &rhs) { drop_group(); conf = rhs.conf; kernels = std::move(rhs.kernels); group = std::exchange(rhs.group, nullptr); }
        return *this;
    }
    ~scheduler_impl() { drop_group(); }

    // max_devices = 0: every visible device (the reference's behaviour)
    explicit scheduler_impl(const Conf &p_conf, size_t max_devices = 0) : conf{p_conf}
    {
        size_t n_dev = static_cast<size_t>(cuda::device_count());
        if (max_devices && max_devices < n_dev) n_dev = max_devices;
        kernels.reserve(n_dev);
        for (size_t i = 0; i < n_dev; ++i) {
            try {
                kernels.emplace_back(conf, static_cast<int>(i));
            } catch (cuda::exception &) {
                // Do not use this device.
            } catch (std::bad_alloc &) {
                // Do not use this device.
            }
        }
        if (kernels.size() == 0) throw cuda::exception("cuda_scheduler: Failed to create kernels.");
    }

    void compute_rho(size_t n, size_t q_begin, size_t q_end)
    {
        if (q_begin == q_end) return;
        split(q_begin, q_end, [&](size_t i, size_t a, size_t b) { kernels[i].compute_rho(n, a, b); });
    }
    void download_rho(double *rho)
    {
        for (auto &k : kernels) k.download_rho(rho);
    }
    void upload_phi(size_t n, const double *coeffs)
    {
        for (auto &k : kernels) k.upload_phi(n, coeffs);
    }
    void compute_metrics(size_t n, size_t q_begin, size_t q_end)
    {
        if (q_begin == q_end) return;
        split(q_begin, q_end, [&](size_t i, size_t a, size_t b) { kernels[i].compute_metrics(n, a, b); });
    }
    void download_metrics(double *metrics)
    {
        for (auto &k : kernels) k.download_metrics(metrics);
    }

    // ---- beyond the reference: the whole time step on the devices
    void step(size_t n)
    {
        if (kernels.size() == 1) { kernels[0].step(n); return; }
        if (!group) {
            std::vector<nufi_b200_handle *> hs;
            for (auto &k : kernels) hs.push_back(k.handle());
            cuda::check(nufi_b200_group_create(hs.data(), static_cast<int>(hs.size()), &group), nufi_b200_group_last_error(nullptr));
        }
        cuda::check(nufi_b200_group_step(group, n), nufi_b200_group_last_error(group));
    }
    double electric_energy(size_t n) { return kernels[0].electric_energy(n); }
    void download_phi(size_t n, double *level) { kernels[0].download_phi(n, level); }
    void sync() { for (auto &k : kernels) k.sync(); }
    size_t device_count() const noexcept { return kernels.size(); }
    kernel_impl<Conf, order> &kernel(size_t i) { return kernels[i]; }
    void set_metrics_grid(const Conf &conf_metrics) { for (auto &k : kernels) k.set_metrics_grid(conf_metrics); }

private:
    template <typename F> void split(size_t q_begin, size_t q_end, F &&f)
    {
        const size_t n_cards = kernels.size(), N = q_end - q_begin;
        const size_t chunk_size = N / n_cards, remainder = N % n_cards;
        size_t current = q_begin;
        for (size_t i = 0; i < n_cards; ++i) {
            const size_t len = chunk_size + (i < remainder ? 1 : 0);
            f(i, current, current + len);
            current += len;
        }
    }
    void drop_group() noexcept { if (group) nufi_b200_group_destroy(group); group = nullptr; }

    Conf conf;
    std::vector<kernel_impl<Conf, order>> kernels;
    nufi_b200_group *group = nullptr;
};

template <typename real> struct require_double
{
    static_assert(std::is_same<real, double>::value, "libnufi_b200 computes in FP64: instantiate with real = double");
};

} // namespace detail

#define NUFI_B200_DEFINE_SCHEDULER(DIM)                                                                                 \
    namespace DIM                                                                                                       \
    {                                                                                                                   \
    template <typename real, size_t order> class cuda_kernel : detail::require_double<real>, public detail::kernel_impl<config_t<real>, order> \
    {                                                                                                                   \
    public:                                                                                                             \
        cuda_kernel(const config_t<real> &conf, int dev = -1) : detail::kernel_impl<config_t<real>, order>(conf, dev) {} \
    };                                                                                                                  \
    template <typename real, size_t order> class cuda_scheduler : detail::require_double<real>, public detail::scheduler_impl<config_t<real>, order> \
    {                                                                                                                   \
    public:                                                                                                             \
        explicit cuda_scheduler(const config_t<real> &conf, size_t max_devices = 0) : detail::scheduler_impl<config_t<real>, order>(conf, max_devices) {} \
    };                                                                                                                  \
    }

NUFI_B200_DEFINE_SCHEDULER(dim2)
NUFI_B200_DEFINE_SCHEDULER(dim3)
#undef NUFI_B200_DEFINE_SCHEDULER

// dim1 has a second constructor with a separate metrics grid (nufi/cuda_kernel.hpp:36-37, nufi/cuda_scheduler.hpp:65-85):
// compute_metrics then integrates over the (x,u) nodes of conf_metrics while eval_f uses the field grid of conf
// (nufi/cuda_kernel.cu:55-70).
namespace dim1
{
template <typename real, size_t order> class cuda_kernel : detail::require_double<real>, public detail::kernel_impl<config_t<real>, order>
{
public:
    cuda_kernel(const config_t<real> &conf, int dev = -1) : detail::kernel_impl<config_t<real>, order>(conf, dev) {}
    cuda_kernel(const config_t<real> &conf, const config_t<real> &conf_metrics, int dev = -1) : detail::kernel_impl<config_t<real>, order>(conf, dev)
    {
        this->set_metrics_grid(conf_metrics);
    }
};
template <typename real, size_t order> class cuda_scheduler : detail::require_double<real>, public detail::scheduler_impl<config_t<real>, order>
{
public:
    explicit cuda_scheduler(const config_t<real> &conf, size_t max_devices = 0) : detail::scheduler_impl<config_t<real>, order>(conf, max_devices) {}
    cuda_scheduler(const config_t<real> &conf, const config_t<real> &conf_metrics) : detail::scheduler_impl<config_t<real>, order>(conf, 0)
    {
        this->set_metrics_grid(conf_metrics);
    }
};
} // namespace dim1


} // namespace nufi

#endif
