// nufi/cuda_runtime.hpp -- error type and device helpers of the reference's nufi/cuda_runtime.hpp:46-51, 93-98, 196-218,
// on top of the C ABI (no CUDA headers are needed to compile a driver).
#ifndef NUFI_B200_NUFI_CUDA_RUNTIME_HPP
#define NUFI_B200_NUFI_CUDA_RUNTIME_HPP

#include <new>
#include <stdexcept>
#include <string>

#include "../nufi_b200.h"

namespace nufi
{

namespace cuda
{

// The reference throws cuda::exception (a std::runtime_error carrying "cudaGetErrorName: cudaGetErrorString").
class exception : public std::runtime_error
{
public:
    explicit exception(const std::string &msg) : std::runtime_error{msg} {}
};

inline int device_count()
{
    int n = 0;
    if (nufi_b200_device_count(&n) != NUFI_B200_OK) throw exception{nufi_b200_last_error(nullptr)};
    return n;
}

// C ABI status -> the exception type the reference would have thrown
inline void check(int rc, const char *msg)
{
    switch (rc) {
    case NUFI_B200_OK: return;
    case NUFI_B200_ERR_RANGE: throw std::range_error{msg};       // nufi/cuda_kernel.cu:115-116, 306-307
    case NUFI_B200_ERR_ALLOC: throw std::bad_alloc{};            // new[] / aligned_alloc in the drivers
    case NUFI_B200_ERR_ARG: throw std::invalid_argument{msg};
    default: throw exception{msg};                               // nufi/cuda_runtime.hpp:93-98
    }
}

} // namespace cuda

} // namespace nufi

#endif
