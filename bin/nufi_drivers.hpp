// bin/nufi_drivers.hpp -- the time loops of the reference's drivers, written against the reference's own entry points
// (nufi::dimN::config_t, cuda_scheduler, poisson, interpolate, eval_rho) as provided by include/nufi/*.hpp over
// libnufi_b200.so.
//
//   gpu_main<DIM>: the loop of bin/test_nufi_gpu_{1,2,3}d.cpp (reference bin/test_nufi_gpu_3d.cpp:150-178, do_stats
//                  :187-224): compute_rho -> download_rho -> poisson.solve -> interpolate -> upload_phi, statistics.csv.
//                  --fused replaces the loop body by sched.step(n): the same step entirely on the device(s).
//   cpu_main<DIM>: the loop of bin/test_nufi_cpu_{1,2,3}d.cpp (reference bin/test_nufi_cpu_2d.cpp:63-107):
//                  "#pragma omp parallel for: rho[l] = eval_rho(n,l,coeffs,conf)" -> solve -> interpolate.
// The reference's drivers take no arguments; these accept a few so the same binaries serve as tests and benchmarks:
//   --steps N   run N time steps instead of conf.Nt      --fused       device-resident step (gpu drivers)
//   --stats K   metrics every K steps (0: never)         --gpus N      use at most N devices (0: all visible)
//   --landau    weak Landau damping f0 (1d: alpha 0.01 k 0.5; 2d: 0.05, 0.5; 3d: 0.001, 0.2 on a 10 pi box, |v| <= 6)
//   --quiet     no per-step line                          --energy FILE write "t energy" per step
//   --metrics-grid NX NU   (gpu 1d) integrate the statistics on an NX x NU grid of their own: the reference's two-config
//               scheduler, bin/test_nufi_gpu_1d.cpp:216-230
// The 1d CPU driver also writes the reference's stats.txt (t, max|E|, ||E||^2 on 256 plot points, every second step) and
// E_<t>.txt (every 160th step) into the working directory, bin/test_nufi_cpu_1d.cpp:82-119.
#ifndef NUFI_B200_BIN_DRIVERS_HPP
#define NUFI_B200_BIN_DRIVERS_HPP

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>
#include <string>

#include <nufi/config.hpp>
#include <nufi/cuda_scheduler.hpp>
#include <nufi/fields.hpp>
#include <nufi/poisson.hpp>
#include <nufi/rho.hpp>
#include <nufi/stopwatch.hpp>

namespace nufi_drivers
{

struct options
{
    size_t steps = 0, stats_every = 1, gpus = 0, metrics_nx = 0, metrics_nu = 0;
    bool fused = false, landau = false, quiet = false;
    std::string energy_file;
};

inline options parse(int argc, char **argv)
{
    options o;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> const char * {
            if (i + 1 >= argc) { std::cerr << "missing value after " << a << std::endl; std::exit(2); }
            return argv[++i];
        };
        if (a == "--steps") o.steps = std::strtoull(next(), nullptr, 10);
        else if (a == "--stats") o.stats_every = std::strtoull(next(), nullptr, 10);
        else if (a == "--gpus") o.gpus = std::strtoull(next(), nullptr, 10);
        else if (a == "--energy") o.energy_file = next();
        else if (a == "--metrics-grid") { o.metrics_nx = std::strtoull(next(), nullptr, 10); o.metrics_nu = std::strtoull(next(), nullptr, 10); }
        else if (a == "--fused") o.fused = true;
        else if (a == "--landau") o.landau = true;
        else if (a == "--quiet") o.quiet = true;
        else { std::cerr << "unknown option " << a << std::endl; std::exit(2); }
    }
    return o;
}

template <int DIM> struct api;
template <> struct api<1>
{
    using conf_t = nufi::dim1::config_t<double>;
    using sched_t = nufi::dim1::cuda_scheduler<double, 4>;
    using poisson_t = nufi::dim1::poisson<double>;
    static void interpolate(double *c, const double *v, const conf_t &cf) { nufi::dim1::interpolate<double, 4>(c, v, cf); }
    static double eval_rho(size_t n, size_t l, const double *c, const conf_t &cf) { return nufi::dim1::eval_rho<double, 4>(n, l, c, cf); }
    static void landau(conf_t &) { conf_t::f0_sel = nufi_b200_f0{0, {0.01, 0.5, 0, 0}}; }
};
template <> struct api<2>
{
    using conf_t = nufi::dim2::config_t<double>;
    using sched_t = nufi::dim2::cuda_scheduler<double, 4>;
    using poisson_t = nufi::dim2::poisson<double>;
    static void interpolate(double *c, const double *v, const conf_t &cf) { nufi::dim2::interpolate<double, 4>(c, v, cf); }
    static double eval_rho(size_t n, size_t l, const double *c, const conf_t &cf) { return nufi::dim2::eval_rho<double, 4>(n, l, c, cf); }
    static void landau(conf_t &) { conf_t::f0_sel = nufi_b200_f0{0, {0.05, 0.5, 0, 0}}; }
};
template <> struct api<3>
{
    using conf_t = nufi::dim3::config_t<double>;
    using sched_t = nufi::dim3::cuda_scheduler<double, 4>;
    using poisson_t = nufi::dim3::poisson<double>;
    static void interpolate(double *c, const double *v, const conf_t &cf) { nufi::dim3::interpolate<double, 4>(c, v, cf); }
    static double eval_rho(size_t n, size_t l, const double *c, const conf_t &cf) { return nufi::dim3::eval_rho<double, 4>(n, l, c, cf); }
    static void landau(conf_t &c)
    {
        conf_t::f0_sel = nufi_b200_f0{0, {0.001, 0.2, 0, 0}};
        c.x_max = c.y_max = c.z_max = 10 * M_PI; // the box the reference's comment gives for this f0 (config.hpp:206)
        c.u_min = c.v_min = c.w_min = -6; c.u_max = c.v_max = c.w_max = 6;
        c.derive();
    }
};

template <int DIM> typename api<DIM>::conf_t make_conf(const options &o)
{
    typename api<DIM>::conf_t conf;
    if (o.landau) api<DIM>::landau(conf);
    if (o.steps) conf.Nt = o.steps;
    return conf;
}

template <int DIM> typename api<DIM>::sched_t make_scheduler(const typename api<DIM>::conf_t &conf, const typename api<DIM>::conf_t &conf_metrics, const options &o)
{
    if constexpr (DIM == 1) {
        if (o.metrics_nx && o.metrics_nu) return typename api<1>::sched_t{conf, conf_metrics};
    }
    (void)conf_metrics;
    return typename api<DIM>::sched_t{conf, o.gpus};
}

struct host_buffers
{
    std::unique_ptr<double[]> coeffs;
    std::unique_ptr<double, decltype(std::free) *> rho{nullptr, std::free};
    host_buffers(size_t n_coeffs, size_t n_nodes, size_t alignment) : coeffs{new double[n_coeffs]{}}
    {
        void *tmp = std::aligned_alloc(alignment, (sizeof(double) * n_nodes + alignment - 1) / alignment * alignment);
        if (tmp == nullptr) throw std::bad_alloc{};
        rho.reset(reinterpret_cast<double *>(tmp));
    }
};

// ------------------------------------------------------------------------------------------------ GPU-driver loop
template <int DIM> int gpu_main(int argc, char **argv)
{
    using A = api<DIM>;
    using tr = nufi::detail::conf_traits<typename A::conf_t>;
    const options opt = parse(argc, argv);
    typename A::conf_t conf = make_conf<DIM>(opt);
    const size_t order = 4, stride_t = tr::stride_t(conf, order), n_nodes = tr::nodes(conf), Nquad = tr::quad(conf);

    typename A::conf_t conf_metrics = conf; // 1d: the grid the statistics are integrated on (bin/test_nufi_gpu_1d.cpp:216-221)
    if (opt.metrics_nx && opt.metrics_nu) {
        if (DIM != 1) { std::cerr << "--metrics-grid: dim 1 only" << std::endl; return 2; }
        conf_metrics.Nx = opt.metrics_nx;
        conf_metrics.Nu = opt.metrics_nu;
        conf_metrics.derive();
    }
    typename A::sched_t sched = make_scheduler<DIM>(conf, conf_metrics, opt);
    typename A::poisson_t poiss{conf};
    const size_t metrics_quad = tr::quad(conf_metrics);
    // one process drives all visible devices (no MPI here): this rank owns the whole quadrature range
    const size_t my_begin = 0, my_end = Nquad;
    std::cout << "Running NuFI on " << sched.device_count() << " GPU(s), " << Nquad << " quadrature points, "
              << (opt.fused ? "device-resident step" : "reference scheduler loop") << std::endl;

    std::ofstream statistics_file("statistics.csv");
    statistics_file << R"("Time"; "L1-Norm"; "L2-Norm"; "Electric Energy"; "Kinetic Energy"; "Total Energy"; "Entropy")" << std::endl;
    statistics_file << std::scientific;
    if (DIM == 1) statistics_file << std::setprecision(16); // bin/test_nufi_gpu_1d.cpp:282; the 2d/3d drivers keep the default 6
    std::cout << std::scientific;
    std::ofstream energy_file;
    if (!opt.energy_file.empty()) { energy_file.open(opt.energy_file); energy_file << std::setprecision(17); }

    host_buffers buf((conf.Nt + 1) * stride_t, n_nodes, A::poisson_t::alignment);
    double *coeffs = buf.coeffs.get(), *rho = buf.rho.get();

    double compute_time_total = 0;
    for (size_t n = 0; n <= conf.Nt; ++n) {
        nufi::stopwatch<double> timer;
        double electric_energy;
        if (opt.fused) {
            sched.step(n);
            electric_energy = sched.electric_energy(n); // blocks until the step is done
        } else {
            std::memset(rho, 0, sizeof(double) * n_nodes);
            sched.compute_rho(n, my_begin, my_end);
            sched.download_rho(rho);
            electric_energy = poiss.solve(rho);
            A::interpolate(coeffs + n * stride_t, rho, conf);
            sched.upload_phi(n, coeffs);
        }
        const double compute_time_step = timer.elapsed();
        compute_time_total += compute_time_step;
        if (!opt.quiet) std::cout << n * conf.dt << " " << compute_time_step << " " << compute_time_total << std::endl;
        if (energy_file.is_open()) energy_file << n * conf.dt << " " << electric_energy << "\n";

        if (opt.stats_every && n % opt.stats_every == 0) {
            double metrics[4]{0, 0, 0, 0};
            sched.compute_metrics(n, 0, metrics_quad);
            sched.download_metrics(metrics);
            metrics[1] = std::sqrt(metrics[1]); // square root for the L2 norm
            const double kinetic_energy = metrics[2], total_energy = kinetic_energy + electric_energy;
            statistics_file << conf.dt * n << "; " << metrics[0] << "; " << metrics[1] << "; " << electric_energy << "; "
                            << kinetic_energy << "; " << total_energy << "; " << metrics[3] << std::endl;
        }
    }
    std::cout << "Total compute time: " << compute_time_total << std::endl;
    return 0;
}

// stats.txt / E_<t>.txt of the reference's 1d CPU driver (bin/test_nufi_cpu_1d.cpp:82-119): every second step the field
// E = -phi' is sampled on 256 equispaced points; one row "t  max|E|  sum E^2 * conf.dx" (widths 20, 8 digits, scientific)
// goes to stats.txt, and every 160th step the samples go to E_<t>.txt as "x E" lines.
inline void write_stats_1d(size_t n, const double *level, const nufi::dim1::config_t<double> &conf, std::ofstream &stats_file)
{
    if (n % 2 != 0) return;
    const size_t plot_x = 256;
    const double dx_plot = conf.Lx / plot_x, t = n * conf.dt;
    std::ofstream file_E;
    if (n % (10 * 16) == 0) file_E.open("E_" + std::to_string(t) + ".txt");
    double Emax = 0, E_l2 = 0;
    for (size_t i = 0; i < plot_x; ++i) {
        const double x = conf.x_min + i * dx_plot;
        const double E = -nufi::dim1::eval<double, 4, 1>(x, level, conf);
        Emax = std::max(Emax, std::abs(E));
        E_l2 += E * E;
        if (file_E.is_open()) file_E << x << " " << E << std::endl;
    }
    E_l2 *= conf.dx; // sic: the reference scales by the grid's dx, not by dx_plot
    stats_file << std::setw(20) << t << std::setw(20) << std::setprecision(8) << std::scientific << Emax << std::setw(20)
               << std::setprecision(8) << std::scientific << E_l2 << std::endl;
}

// ------------------------------------------------------------------------------------------------ CPU-driver loop
template <int DIM> int cpu_main(int argc, char **argv)
{
    using A = api<DIM>;
    using tr = nufi::detail::conf_traits<typename A::conf_t>;
    const options opt = parse(argc, argv);
    typename A::conf_t conf = make_conf<DIM>(opt);
    const size_t order = 4, stride_t = tr::stride_t(conf, order), n_nodes = tr::nodes(conf);

    host_buffers buf((conf.Nt + 1) * stride_t, n_nodes, A::poisson_t::alignment);
    double *coeffs = buf.coeffs.get(), *rho = buf.rho.get();
    typename A::poisson_t poiss(conf);
    std::ofstream energy_file;
    if (!opt.energy_file.empty()) { energy_file.open(opt.energy_file); energy_file << std::setprecision(17); }

    std::ofstream stats_file;
    if (DIM == 1) stats_file.open("stats.txt");
    double total_time = 0;
    // the reference's 1d driver runs n = 0..Nt inclusive (bin/test_nufi_cpu_1d.cpp:60), its 2d/3d drivers n < Nt (_2d.cpp:63)
    const size_t n_end = DIM == 1 ? conf.Nt + 1 : conf.Nt;
    for (size_t n = 0; n < n_end; ++n) {
        nufi::stopwatch<double> timer;

        // Compute rho: the reference's loop, verbatim in shape; the first call of a step runs the sweep on the device.
#pragma omp parallel for
        for (size_t l = 0; l < n_nodes; l++) rho[l] = A::eval_rho(n, l, coeffs, conf);

        const double E_energy = poiss.solve(rho);
        A::interpolate(coeffs + n * stride_t, rho, conf);

        const double timer_elapsed = timer.elapsed();
        total_time += timer_elapsed;
        if (energy_file.is_open()) energy_file << n * conf.dt << " " << E_energy << "\n";
        if constexpr (DIM == 1) write_stats_1d(n, coeffs + n * stride_t, conf, stats_file);
        if (!opt.quiet)
            std::cout << "n = " << n << " t = " << n * conf.dt << " Comp-time: " << timer_elapsed << ". Total time s.f.: " << total_time << std::endl;
    }
    std::cout << "Total time: " << total_time << std::endl;
    return 0;
}

template <typename F> int guarded(F &&f)
{
    try {
        return f();
    } catch (const std::exception &ex) {
        std::cerr << "error: " << ex.what() << std::endl;
        return 1;
    }
}

} // namespace nufi_drivers

#endif
