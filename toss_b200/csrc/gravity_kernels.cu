// toss_b200/csrc/gravity_kernels.cu -- K1: batched polyhedral gravity; batched RHS; point-in-polyhedron; DFMA peak.
#include <stdlib.h>

#include "gravity.cuh"

// ------------------------------------------------------------------------------------------------
// K1a  thread-per-point.  One CTA = 128 field points.  The pair records (AoS, 352 B) stream through a
// 4-stage shared-memory ring filled by TMA bulk copies (one 11 KB tile = 32 pairs per copy); every
// lane reads the SAME record (LDS.128 broadcast), so there is no reduction at all: each thread owns
// its accumulator.  FP64-pipe bound: ~1 smem wavefront per ~10 DFMA-pipe instructions.
// ------------------------------------------------------------------------------------------------
constexpr int kK1Threads = 128;
constexpr int kK1Stages = 4;
constexpr int kPairTileBytes = kPairAosTile * kPairAosStride * 8;

#ifndef K1_MINBLOCKS
#define K1_MINBLOCKS 4
#endif
#ifndef K1_UNROLL
#define K1_UNROLL 1
#endif
constexpr int kK1Unroll = K1_UNROLL;

template <bool kPot, int kMode = kPairLane>
__global__ void __launch_bounds__(kK1Threads, K1_MINBLOCKS)
gravity_points_kernel(const double *__restrict__ pair_aos, int n_tiles, int last_tile_pairs,
                           const double *__restrict__ pts, int64_t n,
                           double *__restrict__ acc, double *__restrict__ pot, double Grho)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *tiles = reinterpret_cast<double *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)kK1Stages * kPairTileBytes);
    uint64_t *empty = full + kK1Stages;
    const int tid = threadIdx.x, lane = tid & 31;
    constexpr int kTileDoubles = kPairAosTile * kPairAosStride;

    if (tid == 0) {
        for (int s = 0; s < kK1Stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kK1Threads / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        const int pre = n_tiles < kK1Stages ? n_tiles : kK1Stages;
        for (int s = 0; s < pre; ++s) {
            mbar_arrive_expect_tx(&full[s], kPairTileBytes);
            tma_load_1d(tiles + (size_t)s * kTileDoubles, pair_aos + (size_t)s * kTileDoubles, kPairTileBytes, &full[s]);
        }
    }

    const int64_t i = (int64_t)blockIdx.x * kK1Threads + tid;
    const int64_t ic = i < n ? i : n - 1;   // tail threads shadow the last point (keeps the warp converged)
    const double Px = pts[3 * ic], Py = pts[3 * ic + 1], Pz = pts[3 * ic + 2];
    double ax = 0.0, ay = 0.0, az = 0.0, pv = 0.0;

    for (int it = 0; it < n_tiles; ++it) {
        const int s = it % kK1Stages;
        const uint32_t ph = (uint32_t)(it / kK1Stages) & 1u;
        if (tid == 0 && it >= 1) {
            const int prev = it - 1, nxt = prev + kK1Stages;
            if (nxt < n_tiles) {
                const int ps = prev % kK1Stages;
                mbar_wait(&empty[ps], (uint32_t)(prev / kK1Stages) & 1u);
                mbar_arrive_expect_tx(&full[ps], kPairTileBytes);
                tma_load_1d(tiles + (size_t)ps * kTileDoubles, pair_aos + (size_t)nxt * kTileDoubles, kPairTileBytes,
                            &full[ps]);
            }
        }
        mbar_wait(&full[s], ph);
        const double *tile = tiles + (size_t)s * kTileDoubles;
        const int p_end = (it == n_tiles - 1) ? last_tile_pairs : kPairAosTile;
#pragma unroll kK1Unroll
        for (int f = 0; f < p_end; ++f) {
            const double *rec = tile + f * kPairAosStride;
            pair_term<kPot, kMode>([rec](int k) { return rec[k]; }, Px, Py, Pz, ax, ay, az, pv);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (i < n) {
        const double G2 = Grho + Grho;   // pair_term accumulates exact halves
        acc[3 * i] = G2 * ax;
        acc[3 * i + 1] = G2 * ay;
        acc[3 * i + 2] = G2 * az;
        if (kPot) pot[i] = Grho * pv;
    }
}

// ------------------------------------------------------------------------------------------------
// K1b  warp-per-point (few points: the single-point `evaluable(x)` call, compute_motion batches):
// lanes stride the pair records of the flat slot table (coalesced 512 B loads), butterfly-shuffle reduction.
// Same lane->face map and reduction tree as the stepper, so values are identical to the RHS it sees.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_gravity_global(const double *__restrict__ soa, int Ppad32, double Px, double Py,
                                                    double Pz, int lane, double &ax, double &ay, double &az,
                                                    double &pv)
{
    ax = ay = az = pv = 0.0;
    for (int f = lane; f < Ppad32; f += 32) {
        const PairCols<true> cols{reinterpret_cast<const double2 *>(soa) + f, Ppad32};
        pair_term<true, kPairRow>(cols, Px, Py, Pz, ax, ay, az, pv);
    }
    ax = warp_sum(ax);
    ay = warp_sum(ay);
    az = warp_sum(az);
    pv = warp_sum(pv);
}

__global__ void __launch_bounds__(256) gravity_warp_kernel(const double *__restrict__ soa, int Ppad32,
                                                           const double *__restrict__ pts, int64_t n,
                                                           double *__restrict__ acc, double *__restrict__ pot,
                                                           double Grho)
{
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n) return;
    double ax, ay, az, pv;
    warp_gravity_global(soa, Ppad32, pts[3 * w], pts[3 * w + 1], pts[3 * w + 2], lane, ax, ay, az, pv);
    if (lane == 0) {
        const double G2 = Grho + Grho;   // pair_term accumulates exact halves
        acc[3 * w] = G2 * ax;
        acc[3 * w + 1] = G2 * ay;
        acc[3 * w + 2] = G2 * az;
        if (pot) pot[w] = Grho * pv;
    }
}

// equations_of_motion.py:62-95 for a batch of (t, y): rotate r by -w t, gravity there, NOT rotated back.
__global__ void __launch_bounds__(256) compute_motion_kernel(const double *__restrict__ soa, int Ppad32, double Grho,
                                                             UnitAxis k, double w, int rotate,
                                                             const double *__restrict__ ts,
                                                             const double *__restrict__ ys, int64_t n,
                                                             double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const double *y = ys + 6 * i;
    double px = y[0], py = y[1], pz = y[2];
    if (rotate) rotate_point_dev(ts[i], y[0], y[1], y[2], k.x, k.y, k.z, w, px, py, pz);
    double ax, ay, az, pv;
    warp_gravity_global(soa, Ppad32, px, py, pz, lane, ax, ay, az, pv);
    if (lane == 0) {
        double *o = out + 6 * i;
        o[0] = y[3];
        o[1] = y[4];
        o[2] = y[5];
        const double G2 = Grho + Grho;   // pair_term accumulates exact halves
        o[3] = -(G2 * ax);
        o[4] = -(G2 * ay);
        o[5] = -(G2 * az);
    }
}

// equations_of_motion.py:98-117 for a batch
__global__ void rotate_points_kernel(UnitAxis k, double w, const double *__restrict__ ts, const double *__restrict__ xs,
                                     int64_t n, double *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double ox, oy, oz;
    rotate_point_dev(ts[i], xs[3 * i], xs[3 * i + 1], xs[3 * i + 2], k.x, k.y, k.z, w, ox, oy, oz);
    out[3 * i] = ox;
    out[3 * i + 1] = oy;
    out[3 * i + 2] = oz;
}

// K5 standalone: mesh_utility.py:70-89.  Warp per point, lanes stride the faces, integer hit count.
__global__ void __launch_bounds__(256) is_outside_kernel(const double *__restrict__ soa, int Fpad32,
                                                         const double *__restrict__ pts, int64_t n,
                                                         uint8_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const double ox = pts[3 * i], oy = pts[3 * i + 1], oz = pts[3 * i + 2];
    int hits = 0;
    for (int f = lane; f < Fpad32; f += 32) {
        const double *c = soa + f;
        const size_t S = (size_t)Fpad32;
        hits += ray_hits_dev(ox, oy, oz, c[0], c[S], c[2 * S], c[3 * S], c[4 * S], c[5 * S], c[6 * S], c[7 * S],
                             c[8 * S]);
    }
    hits = __reduce_add_sync(0xffffffffu, hits);
    if (lane == 0) out[i] = (uint8_t)((hits & 1) == 0);
}

// Register-resident DFMA chains on every SM: the FP64 roofline denominator measured on the box.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c);
            a1 = fma(a1, m, c);
            a2 = fma(a2, m, c);
            a3 = fma(a3, m, c);
            a4 = fma(a4, m, c);
            a5 = fma(a5, m, c);
            a6 = fma(a6, m, c);
            a7 = fma(a7, m, c);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// tmath.cuh functions by id, for the bit-parity tests:
// 0 two_atanh 1 log 2 atan2(a,b) 3 sincos 4 root8 5 root4 6 atan series (8 terms) 7 a/b 8 sqrt
// 9 2 atanh minimax (z*z <= 0.04) 10 2 atan minimax (t*t <= 0.01)
// 11 / 12 far-tier 2 atanh / 2 atan (<= 0.0025) 13 middle-tier 2 atanh (<= 0.01) 14 wide 2 atanh (<= 0.16): twice the
// half values the pair sum works with (the doubling is exact)
__global__ void math_probe_kernel(int fn, const double *__restrict__ a, const double *__restrict__ b, int64_t n,
                                  double *__restrict__ out, double *__restrict__ out2)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r = 0.0, r2 = 0.0;
    switch (fn) {
    case 0: r = tm_two_atanh(a[i]); break;
    case 1: r = tm_log(a[i]); break;
    case 2: r = tm_atan2(a[i], b[i]); break;
    case 3: tm_sincos(a[i], r, r2); break;
    case 4: r = tm_root8(a[i]); break;
    case 5: r = tm_root4(a[i]); break;
    case 6: r = tm_atan_series<8>(a[i]); break;
    case 7: r = tm_div(a[i], b[i]); break;
    case 8: r = tm_sqrt(a[i]); break;
    case 9: r = tm_two_atanh_mm(a[i], a[i] * a[i]); break;
    case 10: r = tm_two_atan_mm(a[i], a[i] * a[i]); break;
    case 11: r = 2.0 * tm_atanh_far(a[i], a[i] * a[i]); break;
    case 12: r = 2.0 * tm_atan_far(a[i], a[i] * a[i]); break;
    case 13: r = 2.0 * tm_atanh_mid(a[i], a[i] * a[i]); break;
    case 14: r = 2.0 * tm_atanh_wide(a[i], a[i] * a[i]); break;
    default: r = __longlong_as_double(0x7ff8000000000000LL);
    }
    out[i] = r;
    if (out2) out2[i] = r2;
}

// tm_div / tm_sqrt against the guarded IEEE intrinsics on pseudo-random operands spanning 2^-200 .. 2^200
__global__ void div_sqrt_selftest_kernel(unsigned long long seed, int iters, unsigned long long *mismatch)
{
    unsigned long long x = seed + 0x9E3779B97F4A7C15ULL * ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x + 1);
    unsigned long long bad_div = 0, bad_sqrt = 0;
    for (int i = 0; i < iters; ++i) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        unsigned long long m1 = x & 0x000fffffffffffffULL;
        int e1 = (int)((x >> 52) % 401) - 200;
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        unsigned long long m2 = x & 0x000fffffffffffffULL;
        int e2 = (int)((x >> 52) % 401) - 200;
        const double a = __longlong_as_double((long long)(m1 | ((unsigned long long)(e1 + 1023) << 52)));
        const double b = __longlong_as_double((long long)(m2 | ((unsigned long long)(e2 + 1023) << 52)));
        if (__double_as_longlong(tm_div(a, b)) != __double_as_longlong(__ddiv_rn(a, b))) ++bad_div;
        if (__double_as_longlong(tm_sqrt(a)) != __double_as_longlong(__dsqrt_rn(a))) ++bad_sqrt;
    }
    if (bad_div) atomicAdd(&mismatch[0], bad_div);
    if (bad_sqrt) atomicAdd(&mismatch[1], bad_sqrt);
}

// ---------------------------------------- launchers ---------------------------------------------
int launch_math_probe(toss_ctx *ctx, int fn, const double *a, const double *b, int64_t n, double *out, double *out2,
                      cudaStream_t st)
{
    if (n <= 0) return 0;
    math_probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(fn, a, b, n, out, out2);
    ctx->launches++;
    TOSS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_div_sqrt_selftest(toss_ctx *ctx, int64_t samples, int64_t *bad_div, int64_t *bad_sqrt)
{
    unsigned long long *d = nullptr, h[2] = {0, 0};
    TOSS_CHECK_CUDA(cudaMalloc(&d, 16));
    TOSS_CHECK_CUDA(cudaMemset(d, 0, 16));
    const int blocks = ctx->sm_count * 8, threads = 256;
    const int iters = (int)((samples + (int64_t)blocks * threads - 1) / ((int64_t)blocks * threads));
    div_sqrt_selftest_kernel<<<blocks, threads>>>(0x1234567ULL, iters, d);
    ctx->launches++;
    TOSS_CHECK_CUDA(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    cudaFree(d);
    *bad_div = (int64_t)h[0];
    *bad_sqrt = (int64_t)h[1];
    return 0;
}

int launch_gravity(toss_ctx *ctx, const double *d_pts, int64_t n, double *d_acc, double *d_pot, cudaStream_t st)
{
    TOSS_REQUIRE(ctx && ctx->d_pair_aos, "toss_gravity_eval: no mesh uploaded");
    if (n <= 0) return 0;
    if (n < 4096) {   // not enough points to fill the machine with one thread each
        const int64_t threads = n * 32;
        const int64_t blocks = (threads + 255) / 256;
        gravity_warp_kernel<<<(unsigned)blocks, 256, 0, st>>>(ctx->d_pair_soa, ctx->Ppad32, d_pts, n, d_acc, d_pot,
                                                            ctx->Grho);
    } else {
        const size_t smem = (size_t)kK1Stages * kPairTileBytes + 2 * kK1Stages * sizeof(uint64_t);
        const int64_t blocks = (n + kK1Threads - 1) / kK1Threads;
        TOSS_REQUIRE(blocks < 2147483647LL, "toss_gravity_eval: too many points for one launch");
        const int last = ctx->n_pairs - (ctx->n_pair_aos_tiles - 1) * kPairAosTile;   // skip the padding records
#define TOSS_LAUNCH_K1(POT, MODE)                                                                                  \
    do {                                                                                                           \
        TOSS_CHECK_CUDA(cudaFuncSetAttribute(gravity_points_kernel<POT, MODE>,                                     \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
        gravity_points_kernel<POT, MODE><<<(unsigned)blocks, kK1Threads, smem, st>>>(                              \
            ctx->d_pair_aos, ctx->n_pair_aos_tiles, last, d_pts, n, d_acc, POT ? d_pot : nullptr, ctx->Grho);      \
    } while (0)
        // per-lane far tier when the faces are small (longest edge below kLaneFarEdge metres): see gravity.cuh
        const bool lane_far = ctx->max_edge > 0.0 && ctx->max_edge < kLaneFarEdge;
        if (d_pot) {
            if (lane_far) TOSS_LAUNCH_K1(true, kPairLaneFar);
            else TOSS_LAUNCH_K1(true, kPairLane);
        } else {
            if (lane_far) TOSS_LAUNCH_K1(false, kPairLaneFar);
            else TOSS_LAUNCH_K1(false, kPairLane);
        }
#undef TOSS_LAUNCH_K1
    }
    ctx->launches++;
    TOSS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_compute_motion(toss_ctx *ctx, const toss_params *p, const double *d_t, const double *d_y, int64_t n,
                          double *d_out, cudaStream_t st)
{
    TOSS_REQUIRE(ctx && ctx->d_pair_soa, "toss_compute_motion: no mesh uploaded");
    if (n <= 0) return 0;
    const int64_t blocks = (n * 32 + 255) / 256;
    compute_motion_kernel<<<(unsigned)blocks, 256, 0, st>>>(ctx->d_pair_soa, ctx->Ppad32, ctx->Grho,
                                                          unit_axis_host(p->spin_axis), p->spin_velocity,
                                                          p->activate_rotation, d_t, d_y, n, d_out);
    ctx->launches++;
    TOSS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_is_outside(toss_ctx *ctx, const double *d_pts, int64_t n, uint8_t *d_out, cudaStream_t st)
{
    TOSS_REQUIRE(ctx && ctx->d_ray_soa, "toss_is_outside: no mesh uploaded");
    if (n <= 0) return 0;
    const int64_t blocks = (n * 32 + 255) / 256;
    is_outside_kernel<<<(unsigned)blocks, 256, 0, st>>>(ctx->d_ray_soa, ctx->Fpad32, d_pts, n, d_out);
    ctx->launches++;
    TOSS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_rotate_points(toss_ctx *ctx, const double *h_axis3, double w, const double *d_t, const double *d_x,
                         int64_t n, double *d_out, cudaStream_t st)
{
    if (n <= 0) return 0;
    rotate_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(unit_axis_host(h_axis3), w, d_t, d_x, n, d_out);
    ctx->launches++;
    TOSS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_fp64_peak(toss_ctx *ctx, double *tflops)
{
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 4096;
    double *d_out = nullptr;
    TOSS_CHECK_CUDA(cudaMalloc(&d_out, sizeof(double) * (size_t)blocks * threads));
    cudaEvent_t e0, e1;
    TOSS_CHECK_CUDA(cudaEventCreate(&e0));
    TOSS_CHECK_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        TOSS_CHECK_CUDA(cudaEventRecord(e0));
        dfma_peak_kernel<<<blocks, threads>>>(d_out, iters, 1.0 + rep);
        TOSS_CHECK_CUDA(cudaEventRecord(e1));
        TOSS_CHECK_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        TOSS_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep >= 1 && tf > best) best = tf;
        ctx->launches++;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    *tflops = best;
    return 0;
}
