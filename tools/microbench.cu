// Dependent-chain latencies on sm_100a of the FP64 / conversion / shared-memory instructions the 1d backtrace step is built
// from (one warp, one CTA; cycles per dependent op from clock64).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
//   -o tools/build/microbench tools/microbench.cu ; run on the GPU box.  Evidence for DESIGN.md's latency-chain analysis.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 4096;

template <int OP> __global__ void chain(double *out, double a, double b, long long *cycles)
{
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1e-9 * i;
    __syncthreads();
    double x = a + threadIdx.x * 1e-3;
    int k = threadIdx.x;
    const long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = fma(x, b, a);                          // DFMA
        if (OP == 1) x = x + b;                                 // DADD
        if (OP == 2) { double y = x + 6755399441055744.0; x = x - (y - 6755399441055744.0) + b; } // magic round: 3 DADD (+1)
        if (OP == 3) x = x - rint(x) + b;                       // FRND + 2 DADD
        if (OP == 4) { long long q = __double2ll_rn(x * 4503599627370496.0); x = static_cast<double>(q & 0xfffffffffffffll) * 2.220446049250313e-16 + b; } // F2I + I2F + mul/fma
        if (OP == 5) { k = (k + __double2loint(x + 6755399441055744.0)) & 1023; x = x + sm[k]; } // DADD, IADD, LOP, LDS, DADD
        if (OP == 6) { k = (k * 3 + 1) & 1023; k = k + __double2loint(sm[k]); }                 // LDS -> int chain
        if (OP == 7) x = fma(x, fma(x, b, a), a);               // 2 DFMA (Horner)
    }
    const long long t1 = clock64();
    out[threadIdx.x] = x + k;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

// Shared-memory load throughput of one SM: `warps` warps issue independent conflict-free loads of WIDTH bytes per lane
// (STRIDE doubles between neighbouring lanes: 1 = dense, 3 = the 1d level layout [p0 p1 p2] per cell) back to back.
template <int WIDTH, int STRIDE> __global__ void lds_tp(double *out, long long *cycles, int iters)
{
    extern __shared__ __align__(16) double smd[];
    for (int i = threadIdx.x; i < 12288; i += blockDim.x) smd[i] = 1e-9 * i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    int base = (warp * 97) & 1023;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            // STRIDE > 0: lane i at i*STRIDE doubles; 0: all lanes one address; < 0: groups of -STRIDE lanes share an address,
            // consecutive groups one cell (2 doubles for the 128-bit case) apart
            const int idx = ((base + u * 40) & 2047) + (STRIDE >= 0 ? lane * STRIDE : (lane / (STRIDE < 0 ? -STRIDE : 1)) * (WIDTH == 16 ? 2 : 1));
            if (WIDTH == 4) {
                acc0 += reinterpret_cast<const float *>(smd)[idx];
            } else if (WIDTH == 8) {
                acc0 += smd[idx];
            } else {
                const double2 v = *reinterpret_cast<const double2 *>(smd + 2 * (idx >> 1) + 0);
                acc0 += v.x;
                acc1 += v.y;
            }
        }
        base += 8;
    }
    const long long t1 = clock64();
    out[threadIdx.x] = acc0 + acc1 + acc2 + acc3;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}


// FP64 issue throughput of one SM against the number of DISTINCT register operands per DFMA (register-file read bandwidth:
// can the pipe sustain one warp-DFMA per 2 cycles per SMSP when none of its three 64-bit operands comes from the operand-reuse
// cache?), and with 64-bit shared-memory loads mixed in.  MODE 0: x = fma(x, a, b) (two loop-invariant operands, the shape of
// the roofline's peak loop); 1: x_i = fma(y_i, c, x_i) (one shared); 2: x_i = fma(y_i, z_i, x_i) (three distinct, the shape of
// a spline contraction); 3: MODE 2 plus one LDS.64 per 4 DFMA (the 3d step's mix).
template <int MODE> __global__ void dfma_tp(double *out, long long *cycles, int iters, double a, double b)
{
    __shared__ double sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = 1e-9 * i;
    __syncthreads();
    double x[8], y[8], z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = a * (i + 1) + threadIdx.x * 1e-6; y[i] = b + i * 1e-3 + threadIdx.x * 1e-9; z[i] = a - i * 1e-3 - threadIdx.x * 1e-9; }
    int k = threadIdx.x & 1023;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) x[i] = fma(x[i], a, b);
                if (MODE == 1) x[i] = fma(y[i], a, x[i]);
                if (MODE >= 2) x[i] = fma(y[i], z[i], x[i]);
            }
            if (MODE == 3) { y[u] += sm[k]; y[u + 4] += sm[k + 32]; k = (k + 64) & 1023; }
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + y[i] + z[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <int MODE> void run_dfma(const char *name, double *d_out, long long *d_cyc, int warps)
{
    const int iters = 2000;
    dfma_tp<MODE><<<1, warps * 32>>>(d_out, d_cyc, iters, 0.999, 1e-3);
    dfma_tp<MODE><<<1, warps * 32>>>(d_out, d_cyc, iters, 0.999, 1e-3);
    long long c = 0;
    cudaMemcpy(&c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double dfma_per_smsp = double(iters) * 32 * warps / 4;
    printf("%-52s %2d warps: %5.2f cycles per warp-DFMA per SMSP (pipe: 2.00)\n", name, warps, c / dfma_per_smsp);
}

template <int WIDTH, int STRIDE> void run_tp(const char *name, double *d_out, long long *d_cyc, int warps)
{
    const int iters = 2000;
    cudaFuncSetAttribute(lds_tp<WIDTH, STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12288 * 8);
    lds_tp<WIDTH, STRIDE><<<1, warps * 32, 12288 * 8>>>(d_out, d_cyc, iters);
    lds_tp<WIDTH, STRIDE><<<1, warps * 32, 12288 * 8>>>(d_out, d_cyc, iters);
    long long c = 0;
    cudaMemcpy(&c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double instr = double(iters) * 8 * warps;
    printf("%-34s %2d warps: %6.2f cycles per warp-load, %6.1f B/clk/SM (includes the dependent FP64 adds)\n", name, warps, c / instr,
           instr * 32 * WIDTH / c);
}

template <int OP> void run(const char *name, double *d_out, long long *d_cyc, double b, int ops)
{
    chain<OP><<<1, 32>>>(d_out, 0.3, b, d_cyc);
    chain<OP><<<1, 32>>>(d_out, 0.3, b, d_cyc);
    long long c = 0;
    cudaMemcpy(&c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
    printf("%-44s %8.2f cycles per iteration  (%d dependent ops -> %.2f per op)\n", name, double(c) / N, ops, double(c) / N / ops);
}

int main()
{
    double *d_out;
    long long *d_cyc;
    cudaMalloc(&d_out, 32 * sizeof(double));
    cudaMalloc(&d_cyc, sizeof(long long));
    run<0>("DFMA chain", d_out, d_cyc, 0.999, 1);
    run<1>("DADD chain", d_out, d_cyc, 1e-3, 1);
    run<2>("magic round (DADD,DADD,DADD,DADD)", d_out, d_cyc, 0.37, 4);
    run<3>("rint + DADD + DADD", d_out, d_cyc, 0.37, 3);
    run<4>("DMUL,F2I.S64,LOP,I2F.F64,DFMA", d_out, d_cyc, 0.37, 5);
    run<5>("DADD,IADD,LOP,LDS.64,DADD", d_out, d_cyc, 0.37, 5);
    run<6>("IMAD,LOP,LDS,F2I-lo,IADD", d_out, d_cyc, 0.37, 5);
    run<7>("DFMA,DFMA (Horner)", d_out, d_cyc, 0.999, 2);
    double *d_big;
    cudaMalloc(&d_big, 1024 * sizeof(double));
    for (int w : {4, 8, 16}) {
        run_dfma<0>("DFMA x=fma(x,a,b): 1 varying operand", d_big, d_cyc, w);
        run_dfma<1>("DFMA x_i=fma(y_i,a,x_i): 2 varying operands", d_big, d_cyc, w);
        run_dfma<2>("DFMA x_i=fma(y_i,z_i,x_i): 3 varying operands", d_big, d_cyc, w);
        run_dfma<3>("  ... + 2 LDS.64 per 8 DFMA", d_big, d_cyc, w);
    }
    for (int w : {4, 8, 16, 32}) {
        run_tp<4, 1>("LDS.32 dense", d_big, d_cyc, w);
        run_tp<8, 1>("LDS.64 dense", d_big, d_cyc, w);
        run_tp<8, 3>("LDS.64 stride 3 doubles", d_big, d_cyc, w);
        run_tp<16, 2>("LDS.128 dense", d_big, d_cyc, w);
        if (w == 16) { // lanes of a warp (nearly) co-located: what a velocity-fastest lane layout reads at shallow history depth
            run_tp<8, 0>("LDS.64 all lanes same address", d_big, d_cyc, w);
            run_tp<16, 0>("LDS.128 all lanes same address", d_big, d_cyc, w);
            run_tp<8, -4>("LDS.64 lanes in groups of 4", d_big, d_cyc, w);
            run_tp<16, -4>("LDS.128 lanes in groups of 4", d_big, d_cyc, w);
            run_tp<8, -2>("LDS.64 lanes in groups of 2", d_big, d_cyc, w);
            run_tp<16, -2>("LDS.128 lanes in groups of 2", d_big, d_cyc, w);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
