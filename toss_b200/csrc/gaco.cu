// toss_b200/csrc/gaco.cu -- candidate generation of the generational optimiser on the device.
//
// The reference evolves every spacecraft's population with pygmo's `gaco` (toss.py:39-68, default_cfg.toml:72-82):
// each generation samples pop_size new decision vectors from a Gaussian kernel mixture centred on the solution archive
// (Schlueter, Egea, Banga 2009) and hands them to the batch fitness evaluator.  Here that hand-over never leaves the
// GPU: this kernel writes the candidates straight into the chromosome matrix the stepper reads (pop 32768 x 12 genes =
// 3.1 MB that would otherwise cross PCIe twice per generation).  One thread per ant:
//   kernel k   = roulette over the cumulative rank weights with one uniform number,
//   gene h     = archive[k][h] + sigma[h] * N(0,1), drawn again while it falls outside [lb[h], ub[h]] (at most 64 draws,
//                then clamped).
// Random numbers are counter-based (splitmix64 of seed / generation / ant / gene / draw): the stream of an ant does not
// depend on the grid, the population size or the GPU count, and toss_b200/optimization/gaco.py holds the numpy mirror
// that the CPU tests and the host path use.
#include <string.h>

#include "common.cuh"

__host__ __device__ inline uint64_t gaco_mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ inline double gaco_u01(uint64_t h) { return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

__global__ void __launch_bounds__(256) gaco_sample_kernel(const double *__restrict__ tab, int ker, int dim, uint64_t key,
                                                          int first_ant, int n, double *__restrict__ out)
{
    // tab = archive[ker][dim] | cumulative weights[ker] | sigma[dim] | lb[dim] | ub[dim]
    extern __shared__ double s_tab[];
    const int words = ker * dim + ker + 3 * dim;
    for (int i = threadIdx.x; i < words; i += blockDim.x) s_tab[i] = tab[i];
    __syncthreads();
    const double *arch = s_tab, *cum = arch + ker * dim, *sigma = cum + ker, *lb = sigma + dim, *ub = lb + dim;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t ka = gaco_mix64(key + (uint64_t)(first_ant + j));
    const double u = gaco_u01(gaco_mix64(ka));
    int k = ker - 1;
    for (int i = 0; i < ker; ++i)
        if (u <= cum[i]) {
            k = i;
            break;
        }
    for (int h = 0; h < dim; ++h) {
        double g = 0.0;
        int t = 0;
        for (; t < 64; ++t) {
            const uint64_t c = (uint64_t)(h * 64 + t) * 2u;
            const double u1 = gaco_u01(gaco_mix64(ka + 1u + c)), u2 = gaco_u01(gaco_mix64(ka + 2u + c));
            g = arch[k * dim + h] + sigma[h] * (sqrt(-2.0 * log(u1)) * cospi(2.0 * u2));
            if (g >= lb[h] && g <= ub[h]) break;
        }
        if (t == 64) g = g < lb[h] ? lb[h] : (g > ub[h] ? ub[h] : g);
        out[(size_t)j * dim + h] = g;
    }
}

int launch_gaco_sample(toss_ctx *ctx, const double *h_archive_x, int ker, int dim, const double *h_cum,
                       const double *h_sigma, const double *h_lb, const double *h_ub, uint64_t seed, uint64_t generation,
                       int first_ant, int n, double *d_x_out, cudaStream_t st)
{
    TOSS_REQUIRE(h_archive_x && h_cum && h_sigma && h_lb && h_ub && d_x_out, "toss_gaco_sample: NULL argument");
    TOSS_REQUIRE(ker > 0 && dim > 0 && n > 0 && first_ant >= 0, "toss_gaco_sample: bad shape");
    const size_t words = (size_t)ker * dim + ker + 3 * (size_t)dim;
    TOSS_REQUIRE(words * 8 <= 160 * 1024, "toss_gaco_sample: archive too large for shared memory");
    if (ctx->gaco_tab_words < words) {
        if (ctx->d_gaco_tab) cudaFree(ctx->d_gaco_tab);
        ctx->d_gaco_tab = nullptr;
        TOSS_CHECK_CUDA(cudaMalloc(&ctx->d_gaco_tab, words * 8));
        ctx->gaco_tab_words = words;
    }
    std::vector<double> tab(words);
    double *w = tab.data();
    memcpy(w, h_archive_x, sizeof(double) * ker * dim), w += (size_t)ker * dim;
    memcpy(w, h_cum, sizeof(double) * ker), w += ker;
    memcpy(w, h_sigma, sizeof(double) * dim), w += dim;
    memcpy(w, h_lb, sizeof(double) * dim), w += dim;
    memcpy(w, h_ub, sizeof(double) * dim);
    // (pageable source: the copy has left the host buffer when the call returns)
    TOSS_CHECK_CUDA(cudaMemcpyAsync(ctx->d_gaco_tab, tab.data(), words * 8, cudaMemcpyHostToDevice, st));
    const uint64_t key = gaco_mix64(seed ^ gaco_mix64(generation));
    if (words * 8 > 48 * 1024)
        TOSS_CHECK_CUDA(cudaFuncSetAttribute(gaco_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(words * 8)));
    gaco_sample_kernel<<<(n + 255) / 256, 256, words * 8, st>>>(ctx->d_gaco_tab, ker, dim, key, first_ant, n, d_x_out);
    ctx->launches++;
    TOSS_CHECK_CUDA(cudaGetLastError());
    return 0;
}
