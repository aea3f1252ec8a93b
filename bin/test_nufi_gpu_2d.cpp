// bin/test_nufi_gpu_2d -- the reference's GPU driver loop (bin/test_nufi_gpu_2d.cpp) on libnufi_b200; see nufi_drivers.hpp.
#include "nufi_drivers.hpp"

int main(int argc, char *argv[])
{
    return nufi_drivers::guarded([&] { return nufi_drivers::gpu_main<2>(argc, argv); });
}
