// nufi/stopwatch.hpp -- wall-clock timer with the reference's interface (nufi/stopwatch.hpp:28-57): starts at
// construction, reset(), elapsed() in seconds.  Uses the steady clock.
#ifndef NUFI_B200_NUFI_STOPWATCH_HPP
#define NUFI_B200_NUFI_STOPWATCH_HPP

#include <chrono>

namespace nufi
{

template <typename real> class stopwatch
{
public:
    void reset() { t0 = clock::now(); }
    real elapsed() const { return std::chrono::duration<real>(clock::now() - t0).count(); }

private:
    using clock = std::chrono::steady_clock;
    clock::time_point t0{clock::now()};
};

} // namespace nufi

#endif
