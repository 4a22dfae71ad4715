// toss_b200/csrc/sharding.cu -- multi-GPU plumbing around the population-fitness path.
//
// The reference spreads udp.fitness calls over a process pool (pg.mp_bfe, toss.py:110) and its scaling script shrinks
// every worker's share as workers are added (scripts/scaling_performance_benchmark.py:126-137).  Here one process per
// GPU takes a share of ONE population.  Trajectory lengths spread 4x, so the shares are dealt by predicted cost:
// all ranks derive the same longest-first order from the all-gathered proxy costs (stepper.cu: cost_order_kernel) and
// rank g takes positions g, g + G, g + 2G, ... of it -- equal work for every GPU, and every shard is already in
// longest-first order for its own work queue.  These kernels move rows and results between the population's order
// and a rank's dealt order; the exchange itself is one NCCL all-gather issued by the host (toss_b200/parallel.py).
#include "common.cuh"

// x_local[j][:] = x[order[rank + j * world]][:]
__global__ void deal_rows_kernel(const double *__restrict__ x, int dim, const int32_t *__restrict__ order, int n,
                                 int rank, int world, double *__restrict__ x_local)
{
    const int n_local = (n - rank + world - 1) / world;
    const long long total = (long long)n_local * dim;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i / dim), c = (int)(i % dim);
        x_local[i] = x[(size_t)order[rank + (long long)j * world] * dim + c];
    }
}

// send = [fitness of the shard, padded with NaN to `cap` | #status 3 | #status 1 or 2]  (one block)
__global__ void __launch_bounds__(256) pack_result_kernel(const double *__restrict__ fit,
                                                          const toss_counters *__restrict__ cnt, int n_local, int cap,
                                                          double *__restrict__ send)
{
    __shared__ int s_short, s_failed;
    if (threadIdx.x == 0) s_short = s_failed = 0;
    __syncthreads();
    int n_short = 0, n_failed = 0;
    for (int i = threadIdx.x; i < cap; i += blockDim.x) {
        send[i] = i < n_local ? fit[i] : __longlong_as_double(0x7ff8000000000000LL);
        if (i < n_local && cnt) {
            const int st = cnt[i].status;
            n_short += st == 3;
            n_failed += (st == 1) | (st == 2);
        }
    }
    if (n_short) atomicAdd(&s_short, n_short);
    if (n_failed) atomicAdd(&s_failed, n_failed);
    __syncthreads();
    if (threadIdx.x == 0) {
        send[cap] = (double)s_short;
        send[cap + 1] = (double)s_failed;
    }
}

// fitness[order[p]] = recv[p % world][p / world];  flags = sums of the two counters over the ranks
__global__ void unpack_result_kernel(const double *__restrict__ recv, const int32_t *__restrict__ order, int n, int world,
                                     int cap, double *__restrict__ fitness, double *__restrict__ flags)
{
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x)
        fitness[order[p]] = recv[(size_t)(p % world) * (cap + 2) + p / world];
    if (blockIdx.x == 0 && threadIdx.x == 0 && flags) {
        double a = 0.0, b = 0.0;
        for (int r = 0; r < world; ++r) {
            a += recv[(size_t)r * (cap + 2) + cap];
            b += recv[(size_t)r * (cap + 2) + cap + 1];
        }
        flags[0] = a;
        flags[1] = b;
    }
}

int launch_deal_rows(toss_ctx *ctx, const double *d_x, int n, int dim, const int32_t *d_order, int rank, int world,
                     double *d_x_local, cudaStream_t st)
{
    TOSS_REQUIRE(d_x && d_order && d_x_local && n > 0 && dim > 0, "toss_deal_rows: bad argument");
    TOSS_REQUIRE(world >= 1 && rank >= 0 && rank < world, "toss_deal_rows: bad rank / world");
    const int n_local = (n - rank + world - 1) / world;
    if (n_local <= 0) return 0;
    const long long total = (long long)n_local * dim;
    const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    deal_rows_kernel<<<blocks, 256, 0, st>>>(d_x, dim, d_order, n, rank, world, d_x_local);
    ctx->launches++;
    TOSS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_pack_result(toss_ctx *ctx, const double *d_fit, const toss_counters *d_cnt, int n_local, int cap,
                       double *d_send, cudaStream_t st)
{
    TOSS_REQUIRE(d_send && cap >= n_local && n_local >= 0 && (n_local == 0 || d_fit), "toss_pack_result: bad argument");
    pack_result_kernel<<<1, 256, 0, st>>>(d_fit, d_cnt, n_local, cap, d_send);
    ctx->launches++;
    TOSS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_unpack_result(toss_ctx *ctx, const double *d_recv, const int32_t *d_order, int n, int world, int cap,
                         double *d_fitness, double *d_flags, cudaStream_t st)
{
    TOSS_REQUIRE(d_recv && d_order && d_fitness && n > 0 && world >= 1 && (long long)cap * world >= n,
                 "toss_unpack_result: bad argument");
    const int blocks = (n + 255) / 256 < 148 ? (n + 255) / 256 : 148;
    unpack_result_kernel<<<blocks, 256, 0, st>>>(d_recv, d_order, n, world, cap, d_fitness, d_flags);
    ctx->launches++;
    TOSS_CHECK_CUDA(cudaGetLastError());
    return 0;
}
