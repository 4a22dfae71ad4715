// developer microbenchmark: how fast can ONE warp issue FP64 instructions, as a function of the number of independent
// dependency chains it carries (1, 2, 4, 8) and of the number of warps resident on its scheduler (1..4)?
// Answers whether instruction-level parallelism inside a warp can shorten a trajectory's critical path.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_warp_ilp ubench_warp_ilp.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void k(double *out, long long *clk, int iters, double a, double b)
{
    double x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = threadIdx.x * 1e-3 + c;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < CH; ++c) x[c] = fma(x[c], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int CH>
void run(int warps, double *d_out, long long *d_clk)
{
    const int iters = 4000;
    k<CH><<<148, 32 * warps>>>(d_out, d_clk, iters, 0.999999, 1e-9);
    k<CH><<<148, 32 * warps>>>(d_out, d_clk, iters, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d_clk, sizeof(h), cudaMemcpyDeviceToHost);
    double m = 0;
    for (int i = 0; i < 148; ++i) m += h[i];
    m /= 148;
    const double per_instr_warp = m / (iters * 8.0 * CH);                       // cycles per DFMA of one warp
    const int per_sched = (warps + 3) / 4;
    printf("warps/SM %2d (per scheduler %d)  chains %d : %.2f cycles per DFMA per warp, %.2f per scheduler\n", warps,
           per_sched, CH, per_instr_warp, per_instr_warp / per_sched);
}

int main()
{
    double *d_out;
    long long *d_clk;
    cudaMalloc(&d_out, 148 * 1024 * 8);
    cudaMalloc(&d_clk, 148 * 8);
    for (int w : {1, 4, 8, 16}) {
        run<1>(w, d_out, d_clk);
        run<2>(w, d_out, d_clk);
        run<4>(w, d_out, d_clk);
        run<8>(w, d_out, d_clk);
    }
    return 0;
}
