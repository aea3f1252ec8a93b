// bin/test_nufi_cpu_1d -- the reference's CPU driver loop (bin/test_nufi_cpu_1d.cpp) with eval_rho / solve / interpolate
// served by libnufi_b200; see nufi_drivers.hpp.
#include "nufi_drivers.hpp"

int main(int argc, char *argv[])
{
    return nufi_drivers::guarded([&] { return nufi_drivers::cpu_main<1>(argc, argv); });
}
