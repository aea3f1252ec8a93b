// nufi/random.hpp -- nufi::random_real<real> / nufi::random_int<Int> with the reference's interface
// (nufi/random.hpp:30-80): callable objects returning uniformly distributed numbers from a DEFAULT-SEEDED
// std::default_random_engine, so every run of a driver draws the same sequence (bin/test_fields.cpp:113-126 relies
// on that).  Every reference driver includes this header (bin/test_nufi_cpu_2d.cpp:30, bin/test_nufi_gpu_3d.cpp:28);
// it is not on the hot path.
#ifndef NUFI_B200_NUFI_RANDOM_HPP
#define NUFI_B200_NUFI_RANDOM_HPP

#include <cstddef>
#include <random>

namespace nufi
{

// uniform reals in [min, max); operator() is const as in the reference, hence the mutable generator state
template <typename real> class random_real
{
public:
    random_real(real min, real max) : dist_(min, max) {}
    real operator()() const { return static_cast<real>(dist_(engine_)); }

private:
    // the reference binds std::uniform_real_distribution<> (i.e. <double>) for every `real`
    mutable std::uniform_real_distribution<> dist_;
    mutable std::default_random_engine engine_;
};

// uniform integers in [min, max]
template <typename Int = int> class random_int
{
public:
    random_int(Int min, Int max) : dist_(min, max) {}
    Int operator()() const { return dist_(engine_); }

private:
    mutable std::uniform_int_distribution<Int> dist_;
    mutable std::default_random_engine engine_;
};

} // namespace nufi

#endif
