// nufi/history_io.hpp -- coefficient-history files: checkpoint / restart and exchange with the reference's dumps.
//
// The history of spline coefficients is the complete simulation state.  Formats:
//  * binary (ours): 64-byte header {magic "NUFIB200", u32 version, u32 dim, u32 order, u32 reserved, u64 Nx, Ny, Nz,
//    u64 n_levels, f64 dt} then n_levels*stride_t doubles in the reference layout.  Exact; restart is bit-identical.
//  * text, "isolated-step" format of bin/test_nufi_cpu_3d_isolated.cpp:64-73, 160-163 (written), :190-211 (read): seven
//    header lines "Nt = ..", "dt = ..", "Nx = ..", "Ny = ..", "Nz = ..", "order = ..", blank; then one coefficient per
//    line with 16 significant digits ((Nt+1)*stride_t values).  (16 digits do not round-trip every double; the writer here
//    uses 17 unless `reference_precision` is set.)
//  * text, plain: one coefficient per line, no header -- bin/test_nufi_gpu_1d.cpp:239, 364-366 (written), :109-123 (read).
#ifndef NUFI_B200_NUFI_HISTORY_IO_HPP
#define NUFI_B200_NUFI_HISTORY_IO_HPP

#include <cstdint>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <stdexcept>
#include <string>
#include <vector>

namespace nufi
{

namespace history_io
{

struct header
{
    char magic[8];
    std::uint32_t version, dim, order, reserved;
    std::uint64_t Nx, Ny, Nz, n_levels;
    double dt;
};
static_assert(sizeof(header) == 64, "history header is 64 bytes");

inline size_t stride_t(const header &h)
{
    const size_t o = h.order - 1;
    return (h.Nx + o) * (h.dim >= 2 ? h.Ny + o : 1) * (h.dim >= 3 ? h.Nz + o : 1);
}

// does a header describe histories of the running configuration?  (dim, order and grid must agree before upload_history)
inline bool matches(const header &h, unsigned dim, unsigned order, size_t Nx, size_t Ny = 1, size_t Nz = 1)
{
    return h.dim == dim && h.order == order && h.Nx == Nx && (dim < 2 || h.Ny == Ny) && (dim < 3 || h.Nz == Nz);
}

inline void write_binary(const std::string &path, const header &h_in, const double *coeffs)
{
    header h = h_in;
    std::memcpy(h.magic, "NUFIB200", 8);
    h.version = 1;
    std::ofstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("history_io: cannot open " + path);
    f.write(reinterpret_cast<const char *>(&h), sizeof(h));
    f.write(reinterpret_cast<const char *>(coeffs), static_cast<std::streamsize>(sizeof(double) * h.n_levels * stride_t(h)));
    if (!f) throw std::runtime_error("history_io: write failed: " + path);
}

inline header read_binary(const std::string &path, std::vector<double> &coeffs)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("history_io: cannot open " + path);
    header h;
    f.read(reinterpret_cast<char *>(&h), sizeof(h));
    if (!f || std::memcmp(h.magic, "NUFIB200", 8) != 0 || h.version != 1) throw std::runtime_error("history_io: not a nufi-b200 history: " + path);
    // do not trust the header: every field bounded before it sizes an allocation (a corrupt or foreign file must not wrap the
    // product or ask for petabytes); callers compare the header with their running config (matches) before upload_history
    const size_t lim = size_t(1) << 20;
    if (h.dim < 1 || h.dim > 3 || h.order < 1 || h.order > 8 || h.Nx < 1 || h.Nx > lim || (h.dim >= 2 && (h.Ny < 1 || h.Ny > lim)) ||
        (h.dim >= 3 && (h.Nz < 1 || h.Nz > lim)) || h.n_levels > lim)
        throw std::runtime_error("history_io: implausible header in " + path);
    const long double total = static_cast<long double>(h.n_levels) * static_cast<long double>(stride_t(h));
    if (total > static_cast<long double>(size_t(1) << 37)) throw std::runtime_error("history_io: header of " + path + " asks for more than 1 TiB");
    f.seekg(0, std::ios::end);
    const std::streamoff have = f.tellg();
    if (have < 0 || static_cast<long double>(have) < static_cast<long double>(sizeof(h)) + total * sizeof(double))
        throw std::runtime_error("history_io: truncated history: " + path);
    f.seekg(sizeof(h), std::ios::beg);
    coeffs.resize(h.n_levels * stride_t(h));
    f.read(reinterpret_cast<char *>(coeffs.data()), static_cast<std::streamsize>(sizeof(double) * coeffs.size()));
    if (!f) throw std::runtime_error("history_io: truncated history: " + path);
    return h;
}

// the isolated-step text format; n_levels = Nt + 1 in the reference
inline void write_text_isolated(const std::string &path, const header &h, const double *coeffs, bool reference_precision = false)
{
    std::ofstream f(path);
    if (!f) throw std::runtime_error("history_io: cannot open " + path);
    f << "Nt = " << std::to_string(h.n_levels - 1) << std::endl;
    f << "dt = " << std::to_string(h.dt) << std::endl;
    f << "Nx = " << std::to_string(h.Nx) << std::endl;
    f << "Ny = " << std::to_string(h.Ny) << std::endl;
    f << "Nz = " << std::to_string(h.Nz) << std::endl;
    f << "order = " << std::to_string(h.order) << std::endl;
    f << std::endl;
    f << std::setprecision(reference_precision ? 16 : 17);
    const size_t n = h.n_levels * stride_t(h);
    for (size_t i = 0; i < n; ++i) f << coeffs[i] << "\n";
}

// reads (n_levels * stride) values after skipping `header_lines` lines (7: isolated-step format, 0: plain format)
inline void read_text(const std::string &path, size_t header_lines, size_t n_values, std::vector<double> &coeffs)
{
    std::ifstream f(path);
    if (!f) throw std::runtime_error("history_io: cannot open " + path);
    std::string dump;
    for (size_t i = 0; i < header_lines; ++i) std::getline(f, dump);
    coeffs.resize(n_values);
    for (size_t i = 0; i < n_values; ++i)
        if (!(f >> coeffs[i])) throw std::runtime_error("history_io: too few coefficients in " + path);
}

inline void write_text_plain(const std::string &path, const double *coeffs, size_t n_values, int precision = 17)
{
    std::ofstream f(path);
    if (!f) throw std::runtime_error("history_io: cannot open " + path);
    f << std::setprecision(precision);
    for (size_t i = 0; i < n_values; ++i) f << coeffs[i] << "\n";
}

} // namespace history_io

} // namespace nufi

#endif
