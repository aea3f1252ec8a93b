// backtrace_generic.cu -- the backtrace / sampling kernels for spline orders other than 4.
//
// The reference is generic in the spline order (nufi/splines.hpp:39-110) and instantiates its CUDA kernels for orders 3..8
// (nufi/cuda_kernel.cu:191-203, 373-385, 575-587); every driver runs order 4, which backtrace.cu specialises (per-cell
// polynomial level formats, shared-memory staging).  Here the same persistent kernel (work layout, deterministic slot
// reduction, epilogue, metrics, peer push -- backtrace_kernel.cuh) is instantiated with step_generic<DIM, ORDER>: Cox-de Boor
// basis in registers, window of ORDER^DIM coefficients read from the reference-format level in global memory (L1/L2: the
// history is L2-resident), one point per thread, general (non-power-of-two) periodic wrap.
#include "backtrace_kernel.cuh"

namespace nufi_b200
{

namespace
{

template <int DIM, int ORDER>
cudaError_t launch_one(const BtParams &P, const EpilogueParams &E, unsigned grid, unsigned threads, size_t smem_bytes, cudaStream_t st)
{
    auto kern = backtrace_kernel<DIM, 1, false, false, false, ORDER>;
    if (smem_bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes));
        if (e != cudaSuccess) return e;
    }
    kern<<<grid, threads, smem_bytes, st>>>(P, E);
    return cudaGetLastError();
}

template <int ORDER>
cudaError_t launch_order(int dim, const BtParams &P, const EpilogueParams &E, unsigned grid, unsigned threads, size_t smem_bytes, cudaStream_t st)
{
    if (dim == 1) return launch_one<1, ORDER>(P, E, grid, threads, smem_bytes, st);
    if (dim == 2) return launch_one<2, ORDER>(P, E, grid, threads, smem_bytes, st);
    return launch_one<3, ORDER>(P, E, grid, threads, smem_bytes, st);
}

template <int ORDER> cudaError_t sample_order(int dim, const SampleParams &S, unsigned blocks, cudaStream_t st)
{
    if (dim == 1) sample_f_kernel<1, false, ORDER><<<blocks, 128, 0, st>>>(S);
    else if (dim == 2) sample_f_kernel<2, false, ORDER><<<blocks, 128, 0, st>>>(S);
    else sample_f_kernel<3, false, ORDER><<<blocks, 128, 0, st>>>(S);
    return cudaGetLastError();
}

} // namespace

int generic_max_threads(int dim) { return dim == 1 ? Tune<1, 1, false, 3>::max_threads : Tune<3, 1, false, 3>::max_threads; }

cudaError_t launch_backtrace_generic(int order, int dim, const BtParams &P, const EpilogueParams &E, unsigned grid, unsigned threads,
                                     size_t smem_bytes, cudaStream_t st)
{
    switch (order) {
    case 3: return launch_order<3>(dim, P, E, grid, threads, smem_bytes, st);
    case 5: return launch_order<5>(dim, P, E, grid, threads, smem_bytes, st);
    case 6: return launch_order<6>(dim, P, E, grid, threads, smem_bytes, st);
    case 7: return launch_order<7>(dim, P, E, grid, threads, smem_bytes, st);
    case 8: return launch_order<8>(dim, P, E, grid, threads, smem_bytes, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_sample_f_generic(int order, int dim, const SampleParams &S, unsigned blocks, cudaStream_t st)
{
    switch (order) {
    case 3: return sample_order<3>(dim, S, blocks, st);
    case 5: return sample_order<5>(dim, S, blocks, st);
    case 6: return sample_order<6>(dim, S, blocks, st);
    case 7: return sample_order<7>(dim, S, blocks, st);
    case 8: return sample_order<8>(dim, S, blocks, st);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace nufi_b200
