// developer microbenchmark: what an FP64 instruction costs at the issue port of an sm_100a scheduler, and whether the
// FP64 tensor instruction (DMMA m8n8k4) relieves it.  16 warps per SM (as the stepper), register-resident operands.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/ubench/issue tools/ubench_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int kMode>
__global__ void __launch_bounds__(512, 1) k(double *out, long long *cyc, int iters)
{
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    int i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3;
    double d0 = 0, d1 = 0, d2 = 0, d3 = 0;   // DMMA accumulators: two independent chains
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (kMode == 0 || kMode == 1 || kMode == 2 || kMode == 4) {   // 8 independent DFMA
                a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
                a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
            }
            if (kMode == 1) {   // + 4 integer multiply-adds
                i0 = i0 * 3 + i1; i1 = i1 * 5 + i2; i2 = i2 * 7 + i3; i3 = i3 * 9 + i0;
            }
            if (kMode == 2) {   // + 8 integer multiply-adds
                i0 = i0 * 3 + i1; i1 = i1 * 5 + i2; i2 = i2 * 7 + i3; i3 = i3 * 9 + i0;
                i0 = i0 * 11 + i2; i1 = i1 * 13 + i3; i2 = i2 * 15 + i0; i3 = i3 * 17 + i1;
            }
            if (kMode == 3 || kMode == 4) {   // 2 DMMA m8n8k4 (256 FMA each per warp)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(d0), "+d"(d1) : "d"(b), "d"(c));
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(d2), "+d"(d3) : "d"(c), "d"(b));
            }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + i0 + i1 + i2 + i3 + d0 + d1 + d2 + d3;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int kMode>
void run(const char *name, double dfma_per_it, double other_per_it, double dmma_per_it)
{
    double *out;
    long long *cyc, h[1];
    cudaMalloc(&out, 148 * 512 * 8);
    cudaMalloc(&cyc, 148 * 8);
    const int iters = 20000;
    k<kMode><<<148, 512>>>(out, cyc, 100);
    k<kMode><<<148, 512>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    const double clk = (double)h[0] / iters / 8.0;   // cycles per unrolled group, 16 warps = 4 per scheduler
    printf("%-28s %7.2f clk per group per SM; per scheduler: %.2f clk per warp-group (DFMA %g, other %g, DMMA %g)\n", name, clk,
           clk / 4.0, dfma_per_it, other_per_it, dmma_per_it);
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    run<0>("8 DFMA", 8, 0, 0);
    run<1>("8 DFMA + 4 IMAD", 8, 4, 0);
    run<2>("8 DFMA + 8 IMAD", 8, 8, 0);
    run<3>("2 DMMA", 0, 0, 2);
    run<4>("8 DFMA + 2 DMMA", 8, 0, 2);
    return 0;
}
