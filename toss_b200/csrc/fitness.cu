// toss_b200/csrc/fitness.cu -- K3 fixed-step dense-output resampling, K4 covered-space / penalty fitness,
// K5 collision verdict over the event points.
//
// Replaces (reference file:line)
//   get_trajectory_fixed_step            toss/trajectory/trajectory_tools.py:28-76   (+ desolver's cubic-Hermite
//                                        dense output behind ode_object._OdeSystem__sol, :67)
//   get_fitness + the nine fitness fns   toss/fitness/fitness_functions.py:7-227
//   compute_space_coverage / update_spherical_tensor_grid   toss/fitness/fitness_function_utils.py:54-228
//   point_is_outside_mesh / is_outside   toss/trajectory/compute_trajectory.py:141-159, toss/mesh/mesh_utility.py:70-166
//
// Everything that decides an INTEGER (sample -> segment -> interval, voxel index, hit parity, the
// inclusive radius filters) is computed with un-fused IEEE operations in the reference's order, so the
// integer outputs are bit-exact given identical inputs; the floating-point reductions use a fixed tree.
#include "gravity.cuh"

constexpr int kFitThreads = 256;
constexpr double kPi = 3.14159265358979323846;

int toss_workspace_layout_impl(int pop, int n_man, int knot_cap, toss_ws_layout *out);

// ---------------------------------------------------------------------------------------------
// dense output: desolver's cubic Hermite between consecutive accepted steps.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int bisect_knots(const double *__restrict__ kt, int n, double t)
{
    int lo = 0, hi = n - 1;   // interval lo with kt[lo] < t <= kt[lo+1] (t == kt[0] -> 0)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (kt[mid] < t) lo = mid;
        else hi = mid;
    }
    return lo;
}

struct HermiteW {
    double h00, h10h, h01, h11h;
};
__device__ __forceinline__ HermiteW hermite_weights(double t0, double t1, double t)
{
    const double h = sub_(t1, t0);
    const double s = __ddiv_rn(sub_(t, t0), h), s2 = mul_(s, s), s3 = mul_(s2, s);
    HermiteW w;
    w.h00 = add_(sub_(mul_(2.0, s3), mul_(3.0, s2)), 1.0);
    w.h10h = mul_(add_(sub_(s3, mul_(2.0, s2)), s), h);
    w.h01 = add_(mul_(-2.0, s3), mul_(3.0, s2));
    w.h11h = mul_(sub_(s3, s2), h);
    return w;
}
__device__ __forceinline__ double hermite_eval(const HermiteW &w, double p0, double m0, double p1, double m1)
{
    return add_(add_(add_(mul_(w.h00, p0), mul_(w.h10h, m0)), mul_(w.h01, p1)), mul_(w.h11h, m1));
}

// one segment's interpolant at time t, components [c0, c1)
__device__ __forceinline__ void dense_eval_dev(const double *__restrict__ kt, const double *__restrict__ ky,
                                               const double *__restrict__ kf, int n, double t, int c0, int c1,
                                               double *out)
{
    if (n < 2) {
        for (int c = c0; c < c1; ++c) out[c - c0] = ky[c];
        return;
    }
    const int lo = bisect_knots(kt, n, t);
    const HermiteW w = hermite_weights(kt[lo], kt[lo + 1], t);
    const double *p0 = ky + 6 * (size_t)lo, *p1 = p0 + 6, *m0 = kf + 6 * (size_t)lo, *m1 = m0 + 6;
    for (int c = c0; c < c1; ++c) out[c - c0] = hermite_eval(w, p0[c], m0[c], p1[c], m1[c]);
}

__global__ void dense_eval_kernel(const double *__restrict__ kt, const double *__restrict__ ky,
                                  const double *__restrict__ kf, int n, const double *__restrict__ t, int64_t nt,
                                  double *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    double o[6];
    dense_eval_dev(kt, ky, kf, n, t[i], 0, 6, o);
    for (int c = 0; c < 6; ++c) out[6 * i + c] = o[c];
}

// trajectory_tools.py:55-58: index of the last sample with t <= t_end ("nearest, minus one if later"),
// np.argmin first-minimum semantics; |ts - t_end| is V-shaped so a local scan around the analytic
// guess finds the same index.
__device__ __forceinline__ long long last_sample_index(double t_end, double start, double period, long long N)
{
    long long g = (long long)floor((t_end - start) / period);
    long long lo = g - 1, hi = g + 2;
    lo = lo < 0 ? 0 : (lo > N - 1 ? N - 1 : lo);
    hi = hi < 0 ? 0 : (hi > N - 1 ? N - 1 : hi);
    long long e = lo;
    double best = fabs(sub_(add_(start, mul_((double)lo, period)), t_end));
    for (long long i = lo + 1; i <= hi; ++i) {
        const double d = fabs(sub_(add_(start, mul_((double)i, period)), t_end));
        if (d < best) {
            best = d;
            e = i;
        }
    }
    if (t_end < add_(start, mul_((double)e, period))) e -= 1;
    return e;
}

// per-trajectory segment table in shared memory: seg_e[s] = last sample served by segment s
struct SegTable {
    int n_seg;
    const int32_t *seg_start;
};
__device__ __forceinline__ void build_seg_table(long long *seg_e, int n_seg, const int32_t *seg_start,
                                                const double *kt, double start, double period, long long N)
{
    // sequential, as the reference loop: start_idx advances; a segment may serve no sample
    long long start_idx = 0;
    bool stopped = false;
    for (int s = 0; s < n_seg; ++s) {
        const int k0 = seg_start[s], k1 = seg_start[s + 1];
        if (stopped || k1 <= k0) {
            seg_e[s] = start_idx - 1;
            continue;
        }
        const long long e = last_sample_index(kt[k1 - 1], start, period, N);
        seg_e[s] = e;
        start_idx = e + 1;
        if (N - 1 < start_idx) stopped = true;   // trajectory_tools.py:73-74
    }
}
__device__ __forceinline__ int find_segment(const long long *seg_e, int n_seg, long long i)
{
    for (int s = 0; s < n_seg; ++s)
        if (i <= seg_e[s]) return s;
    return n_seg - 1;   // beyond the last knot (reference leaves np.empty garbage there): last segment
}

// K3 materialising variant: positions / velocities (pop x 3 x N) and timesteps (N)
__global__ void __launch_bounds__(kFitThreads)
resample_kernel(toss_params p, const double *__restrict__ knots_t, const double *__restrict__ knots_y,
                const double *__restrict__ knots_f, const int32_t *__restrict__ seg_start_all,
                const int32_t *__restrict__ n_seg_all, int n_man, int knot_cap, long long N,
                double *__restrict__ pos, double *__restrict__ vel, double *__restrict__ ts)
{
    extern __shared__ long long seg_e[];
    const int traj = blockIdx.x;
    const int n_seg = n_seg_all[traj];
    const int32_t *seg_start = seg_start_all + (size_t)traj * (n_man + 2);
    const double *kt = knots_t + (size_t)traj * knot_cap;
    const double *ky = knots_y + (size_t)traj * knot_cap * 6;
    const double *kf = knots_f + (size_t)traj * knot_cap * 6;
    if (threadIdx.x == 0) build_seg_table(seg_e, n_seg, seg_start, kt, p.start_time, p.measurement_period, N);
    __syncthreads();
    for (long long i = threadIdx.x; i < N; i += blockDim.x) {
        const double t = add_(p.start_time, mul_((double)i, p.measurement_period));
        if (traj == 0 && ts) ts[i] = t;
        const int s = find_segment(seg_e, n_seg, i);
        const int k0 = seg_start[s], k1 = seg_start[s + 1];
        double o[6];
        dense_eval_dev(kt + k0, ky + 6 * (size_t)k0, kf + 6 * (size_t)k0, k1 - k0, t, 0, 6, o);
        for (int c = 0; c < 3; ++c) {
            pos[((size_t)traj * 3 + c) * N + i] = o[c];
            if (vel) vel[((size_t)traj * 3 + c) * N + i] = o[3 + c];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K4: fitness.  One CTA per trajectory; the voxel bitmap lives in shared memory.
// ---------------------------------------------------------------------------------------------
struct GridArgs {
    const double *grid;          // r | theta | phi
    int nr, nth, nph;
    const uint32_t *prev_bits;   // packed bool_tensor
    int n_words;
};

__device__ __forceinline__ int argmin_abs_dev(const double *__restrict__ g, int n, double v)
{   // np.argmin(|grid - v|): first minimum wins (fitness_function_utils.py:137-139)
    int best = 0;
    double bd = fabs(sub_(g[0], v));
    for (int i = 1; i < n; ++i) {
        const double d = fabs(sub_(g[i], v));
        if (d < bd) {
            bd = d;
            best = i;
        }
    }
    return best;
}

// np.array_split(positions, k, axis=1) + timesteps[col] with col restarting in every chunk
// (fitness_function_utils.py:96-102)
__device__ __forceinline__ long long chunk_col_dev(long long n, long long N, int k)
{
    if (k <= 1) return n;
    const long long base = N / k, rem = N % k, big = rem * (base + 1);
    if (n < big) return n % (base + 1);
    if (base == 0) return 0;
    return (n - big) % base;
}

// marks the voxel of one sample (rotated by -w * t_clock) in the shared bitmap
__device__ __forceinline__ void mark_voxel(uint32_t *bits, const GridArgs &g, const UnitAxis &k, double w,
                                           double t_clock, double x, double y, double z, double rmin, double rmax)
{
    double rx, ry, rz;
    rotate_point_dev(t_clock, x, y, z, k.x, k.y, k.z, w, rx, ry, rz);
    // cart2sphere (fitness_function_utils.py:282-314)
    const double xy2 = add_(mul_(rx, rx), mul_(ry, ry));
    const double r = __dsqrt_rn(add_(xy2, mul_(rz, rz)));
    if (!(r <= rmax) || !(r >= rmin)) return;   // :113-122 inclusive bounds
    const double th = tm_atan2(rz, __dsqrt_rn(xy2));
    const double ph = tm_atan2(ry, rx);
    const int i = argmin_abs_dev(g.grid, g.nr, r);
    const int j = argmin_abs_dev(g.grid + g.nr, g.nth, th);
    const int kk = argmin_abs_dev(g.grid + g.nr + g.nth, g.nph, ph);
    const unsigned b = (unsigned)((i * g.nth + j) * g.nph + kk);
    atomicOr(&bits[b >> 5], 1u << (b & 31));
}

struct BlockRed {
    double dsum[kFitThreads / 32];
    double dmin[kFitThreads / 32];
    long long cnt[kFitThreads / 32];
};
// fixed-tree block reductions (deterministic): returns the result in thread 0
__device__ __forceinline__ double block_sum(double v, double *scratch)
{
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kFitThreads / 32; ++w) r += scratch[w];
    __syncthreads();
    return r;
}
__device__ __forceinline__ double block_min(double v, double *scratch)
{
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = v;
    if (threadIdx.x == 0) {
        r = scratch[0];
        for (int w = 1; w < kFitThreads / 32; ++w) r = fmin(r, scratch[w]);
    }
    __syncthreads();
    return r;
}
__device__ __forceinline__ long long block_count(long long v, long long *scratch)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    long long r = 0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kFitThreads / 32; ++w) r += scratch[w];
    __syncthreads();
    return r;
}

// coverage value from the shared bitmap (fitness_function_utils.py:150-161; the "empty" branch :125-133
// is the same expression on the previous tensor): n_true / size + sum_i c_i / r_i.  thread 0 returns it.
__device__ __forceinline__ double coverage_from_bits(const uint32_t *bits, const GridArgs &g, int *shell,
                                                     long long *n_visited)
{
    const int S = g.nth * g.nph;
    for (int i = threadIdx.x; i < g.nr; i += blockDim.x) shell[i] = 0;
    __syncthreads();
    for (int w = threadIdx.x; w < g.n_words; w += blockDim.x) {
        const uint32_t word = bits[w];
        if (!word) continue;
        const long long b0 = 32LL * w;
        int i_lo = (int)(b0 / S), i_hi = (int)((b0 + 31) / S);
        if (i_hi > g.nr - 1) i_hi = g.nr - 1;
        for (int i = i_lo; i <= i_hi; ++i) {
            long long lo = (long long)i * S - b0, hi = (long long)(i + 1) * S - b0;
            if (lo < 0) lo = 0;
            if (hi > 32) hi = 32;
            if (hi <= lo) continue;
            const uint32_t mask = (hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
            const int c = __popc(word & mask);
            if (c) atomicAdd(&shell[i], c);
        }
    }
    __syncthreads();
    double cov = 0.0;
    if (threadIdx.x == 0) {
        long long total = 0;
        double wsum = 0.0;
        for (int i = 0; i < g.nr; ++i) {
            total += shell[i];
            wsum = add_(wsum, __ddiv_rn((double)shell[i], g.grid[i]));
        }
        const double nvox = (double)g.nr * (double)g.nth * (double)g.nph;
        cov = add_(__ddiv_rn((double)total, nvox), wsum);
        if (n_visited) *n_visited = total;
    }
    return cov;
}

// combine the reductions into the requested fitness value (fitness_functions.py:44-227); thread 0 only
struct FitParts {
    double close_min;     // min over samples of (|p|^2 - R_in^2) where negative (+inf if none)
    double far_sum;       // sum of (|p|^2 - R_out^2) where positive
    long long far_cnt;
    double alt_sum;       // sum |(|p|^2 - target)|
    double vol;           // estimate_covered_volume
    double cov;           // compute_space_coverage
};
__device__ __forceinline__ double combine_fitness(int fn, const toss_params &p, const FitParts &f, long long N)
{
    const double rin = p.radius_inner, rout = p.radius_outer, sc = p.penalty_scaling_factor;
    double close = 0.0, far = 0.0;
    if (f.close_min < 0.0) close = mul_(tm_root4(__ddiv_rn(fabs(f.close_min), mul_(rin, rin))), sc);
    if (f.far_cnt > 0) far = mul_(tm_root4(__ddiv_rn(__ddiv_rn(f.far_sum, (double)f.far_cnt), mul_(rout, rout))), sc);
    const double volN = -__ddiv_rn(f.vol, mul_(p.maximal_measurement_sphere_volume, (double)N));
    switch (fn) {
    case 1: return tm_root4(__ddiv_rn(__ddiv_rn(f.alt_sum, (double)N), p.target_squared_altitude));
    case 2: return close;
    case 3: return far;
    case 4: return volN;
    case 5: return -__ddiv_rn(f.vol, p.total_measurable_volume);
    case 6: return add_(volN, far);
    case 7: return add_(add_(volN, close), far);
    case 8: return f.cov;
    case 9: return sub_(add_(close, far), f.cov);
    default: return __longlong_as_double(0x7ff8000000000000LL);
    }
}

__device__ __forceinline__ double sqdist_dev(double x, double y, double z, double c)
{   // fitness_function_utils.py:41-51
    return sub_(add_(add_(mul_(x, x), mul_(y, y)), mul_(z, z)), mul_(c, c));
}

// K4 from materialised positions (batch x 3 x N) -- the get_fitness drop-in.
__global__ void __launch_bounds__(kFitThreads)
fitness_positions_kernel(int fn, toss_params p, UnitAxis axis, GridArgs g, const double *__restrict__ pos_all,
                         long long N, const double *__restrict__ ts, double *__restrict__ fit,
                         int64_t *__restrict__ n_vis, uint32_t *__restrict__ bits_out)
{
    extern __shared__ __align__(16) unsigned char fsm[];
    uint32_t *bits = reinterpret_cast<uint32_t *>(fsm);
    int *shell = reinterpret_cast<int *>(bits + ((g.n_words + 1) & ~1));   // keeps the 8-byte data behind it aligned
    __shared__ double sc_d[kFitThreads / 32];
    __shared__ long long sc_l[kFitThreads / 32];
    const int b = blockIdx.x;
    const double *px = pos_all + (size_t)b * 3 * N, *py = px + N, *pz = py + N;
    const bool need_cov = (fn == 8 || fn == 9);
    if (need_cov) {
        for (int w = threadIdx.x; w < g.n_words; w += blockDim.x) bits[w] = g.prev_bits[w];
        __syncthreads();
    }
    const double c_alt = __dsqrt_rn(p.target_squared_altitude);
    double cmin = __longlong_as_double(0x7ff0000000000000LL), fsum = 0.0, asum = 0.0, vsum = 0.0;
    long long fcnt = 0;
    for (long long n = threadIdx.x; n < N; n += blockDim.x) {
        const double x = px[n], y = py[n], z = pz[n];
        if (fn == 1) asum += fabs(sqdist_dev(x, y, z, c_alt));
        if (fn == 2 || fn == 7 || fn == 9) {
            const double d = sqdist_dev(x, y, z, p.radius_inner);
            if (d < 0.0) cmin = fmin(cmin, d);
        }
        if (fn == 3 || fn == 6 || fn == 7 || fn == 9) {
            const double d = sqdist_dev(x, y, z, p.radius_outer);
            if (d > 0.0) {
                fsum += d;
                ++fcnt;
            }
        }
        if (fn >= 4 && fn <= 7 && N >= 2) {   // fitness_function_utils.py:8-38
            long long j = n;
            if (j == 0) j = 1;
            if (n == N - 1) j = (N - 2 >= 1) ? N - 2 : 1;
            const double dx = sub_(px[j], px[j - 1]), dy = sub_(py[j], py[j - 1]), dz = sub_(pz[j], pz[j - 1]);
            const double rad = __ddiv_rn(__dsqrt_rn(add_(add_(mul_(dx, dx), mul_(dy, dy)), mul_(dz, dz))), 2.0);
            vsum += mul_((4.0 / 3.0) * kPi, mul_(mul_(rad, rad), rad));
        }
        if (need_cov)
            mark_voxel(bits, g, axis, p.spin_velocity, ts[chunk_col_dev(n, N, p.number_of_spacecrafts)], x, y, z,
                       p.radius_inner, p.radius_outer);
    }
    FitParts f;
    f.close_min = block_min(cmin, sc_d);
    f.far_sum = block_sum(fsum, sc_d);
    f.far_cnt = block_count(fcnt, sc_l);
    f.alt_sum = block_sum(asum, sc_d);
    f.vol = block_sum(vsum, sc_d);
    f.cov = 0.0;
    __syncthreads();
    long long nv = -1;
    if (need_cov) f.cov = coverage_from_bits(bits, g, shell, &nv);
    if (threadIdx.x == 0) {
        fit[b] = combine_fitness(fn, p, f, N);
        if (n_vis) n_vis[b] = nv;
    }
    if (bits_out && need_cov)
        for (int w = threadIdx.x; w < g.n_words; w += blockDim.x) bits_out[(size_t)b * g.n_words + w] = bits[w];
}

// K3+K4 fused: udp_initial_condition.py:123-137 after the integration -- resample straight from the knots
// (positions only; velocities are never used by any fitness function) and score with enum 9.
__global__ void __launch_bounds__(kFitThreads)
fitness_knots_kernel(toss_params p, UnitAxis axis, GridArgs g, const double *__restrict__ knots_t,
                     const double *__restrict__ knots_y, const double *__restrict__ knots_f,
                     const int32_t *__restrict__ seg_start_all, const int32_t *__restrict__ n_seg_all,
                     const int32_t *__restrict__ collision, const toss_counters *__restrict__ counters, int n_man,
                     int knot_cap, long long N, double *__restrict__ fit, int64_t *__restrict__ n_vis)
{
    extern __shared__ __align__(16) unsigned char fsm[];
    uint32_t *bits = reinterpret_cast<uint32_t *>(fsm);
    int *shell = reinterpret_cast<int *>(bits + ((g.n_words + 1) & ~1));   // keeps the 8-byte data behind it aligned
    long long *seg_e = reinterpret_cast<long long *>(shell + ((g.nr + 1) & ~1));
    __shared__ double sc_d[kFitThreads / 32];
    __shared__ long long sc_l[kFitThreads / 32];
    const int traj = blockIdx.x;
    if (collision[traj] || counters[traj].status != 0) {   // udp_initial_condition.py:126-128
        if (threadIdx.x == 0) {
            fit[traj] = 1e30;
            if (n_vis) n_vis[traj] = -1;
        }
        return;
    }
    const int n_seg = n_seg_all[traj];
    const int32_t *seg_start = seg_start_all + (size_t)traj * (n_man + 2);
    const double *kt = knots_t + (size_t)traj * knot_cap;
    const double *ky = knots_y + (size_t)traj * knot_cap * 6;
    const double *kf = knots_f + (size_t)traj * knot_cap * 6;
    for (int w = threadIdx.x; w < g.n_words; w += blockDim.x) bits[w] = g.prev_bits[w];
    if (threadIdx.x == 0) build_seg_table(seg_e, n_seg, seg_start, kt, p.start_time, p.measurement_period, N);
    __syncthreads();
    double cmin = __longlong_as_double(0x7ff0000000000000LL), fsum = 0.0;
    long long fcnt = 0;
    for (long long n = threadIdx.x; n < N; n += blockDim.x) {
        const double t = add_(p.start_time, mul_((double)n, p.measurement_period));
        const int s = find_segment(seg_e, n_seg, n);
        const int k0 = seg_start[s], k1 = seg_start[s + 1];
        double o[3];
        dense_eval_dev(kt + k0, ky + 6 * (size_t)k0, kf + 6 * (size_t)k0, k1 - k0, t, 0, 3, o);
        const double din = sqdist_dev(o[0], o[1], o[2], p.radius_inner);
        if (din < 0.0) cmin = fmin(cmin, din);
        const double dout = sqdist_dev(o[0], o[1], o[2], p.radius_outer);
        if (dout > 0.0) {
            fsum += dout;
            ++fcnt;
        }
        const long long col = chunk_col_dev(n, N, p.number_of_spacecrafts);
        mark_voxel(bits, g, axis, p.spin_velocity, add_(p.start_time, mul_((double)col, p.measurement_period)), o[0],
                   o[1], o[2], p.radius_inner, p.radius_outer);
    }
    FitParts f;
    f.close_min = block_min(cmin, sc_d);
    f.far_sum = block_sum(fsum, sc_d);
    f.far_cnt = block_count(fcnt, sc_l);
    f.alt_sum = f.vol = 0.0;
    __syncthreads();
    long long nv = -1;
    f.cov = coverage_from_bits(bits, g, shell, &nv);
    if (threadIdx.x == 0) {
        fit[traj] = combine_fitness(9, p, f, N);
        if (n_vis) n_vis[traj] = nv;
    }
}

// ---------------------------------------------------------------------------------------------
// K5 over the event points: every accepted state with |r|^2 <= R_in^2 (desolver records the step-function
// event at each accepted step inside the risk sphere) is ray-cast against the mesh;
// any point inside -> collision_detected (compute_trajectory.py:141-159).  One warp per trajectory.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
collision_kernel(toss_params p, const double *__restrict__ soa, int Fpad32, const double *__restrict__ knots_y,
                 const int32_t *__restrict__ seg_start_all, const int32_t *__restrict__ n_seg_all, int n_man,
                 int knot_cap, int pop, int32_t *__restrict__ collision, toss_counters *__restrict__ counters)
{
    const int lane = threadIdx.x & 31;
    const int traj = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (traj >= pop) return;
    const int n_seg = n_seg_all[traj];
    const int32_t *seg_start = seg_start_all + (size_t)traj * (n_man + 2);
    const double *ky = knots_y + (size_t)traj * knot_cap * 6;
    const double R2 = mul_(p.radius_inner, p.radius_inner);
    int collided = 0, n_ev = 0;
    if (p.activate_event) {
        for (int s = 0; s < n_seg; ++s) {
            const int k0 = seg_start[s], k1 = seg_start[s + 1];
            for (int k = k0 + 1; k < k1; ++k) {   // the first knot of a segment is its initial state, not an accepted step
                const double x = ky[6 * (size_t)k], y = ky[6 * (size_t)k + 1], z = ky[6 * (size_t)k + 2];
                if (!(add_(add_(mul_(x, x), mul_(y, y)), mul_(z, z)) <= R2)) continue;
                ++n_ev;
                int hits = 0;
                for (int f = lane; f < Fpad32; f += 32) {
                    const double *c = soa + f;
                    const size_t S = (size_t)Fpad32;
                    hits += ray_hits_dev(x, y, z, c[0], c[S], c[2 * S], c[3 * S], c[4 * S], c[5 * S], c[6 * S],
                                         c[7 * S], c[8 * S]);
                }
                hits = __reduce_add_sync(0xffffffffu, hits);
                if (hits & 1) {
                    collided = 1;
                    if (p.stop_on_collision) break;
                }
            }
            if (collided && p.stop_on_collision) break;
        }
    }
    if (lane == 0) {
        collision[traj] = collided;
        counters[traj].n_event_points = n_ev;
    }
}

// ---------------------------------------- launchers ---------------------------------------------
static inline GridArgs grid_args(const toss_ctx *ctx)
{
    GridArgs g;
    g.grid = ctx->d_grid;
    g.nr = ctx->nr;
    g.nth = ctx->nth;
    g.nph = ctx->nph;
    g.prev_bits = ctx->d_bool_bits;
    g.n_words = (int)(((long long)ctx->nr * ctx->nth * ctx->nph + 31) / 32);
    return g;
}

struct WsPtrs {
    double *knots_t, *knots_y, *knots_f, *intervals;
    int32_t *seg_start, *n_seg, *collision;
    toss_counters *counters;
};
static inline WsPtrs ws_ptrs(void *d_ws, int pop, int n_man, int knot_cap)
{
    toss_ws_layout L;
    toss_workspace_layout_impl(pop, n_man, knot_cap, &L);
    unsigned char *ws = static_cast<unsigned char *>(d_ws);
    WsPtrs w;
    w.knots_t = reinterpret_cast<double *>(ws + L.knots_t);
    w.knots_y = reinterpret_cast<double *>(ws + L.knots_y);
    w.knots_f = reinterpret_cast<double *>(ws + L.knots_f);
    w.intervals = reinterpret_cast<double *>(ws + L.intervals);
    w.seg_start = reinterpret_cast<int32_t *>(ws + L.seg_start);
    w.n_seg = reinterpret_cast<int32_t *>(ws + L.n_seg);
    w.collision = reinterpret_cast<int32_t *>(ws + L.collision);
    w.counters = reinterpret_cast<toss_counters *>(ws + L.counters);
    return w;
}

int launch_collision(toss_ctx *ctx, const toss_params *p, void *d_ws, int pop, int n_man, int knot_cap,
                     cudaStream_t st)
{
    WsPtrs w = ws_ptrs(d_ws, pop, n_man, knot_cap);
    const int blocks = (pop * 32 + 255) / 256;
    collision_kernel<<<blocks, 256, 0, st>>>(*p, ctx->d_ray_soa, ctx->Fpad32, w.knots_y, w.seg_start, w.n_seg, n_man,
                                             knot_cap, pop, w.collision, w.counters);
    ctx->launches++;
    TOSS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_resample(toss_ctx *ctx, const toss_params *p, const void *d_ws, int pop, int n_man, int knot_cap,
                    double *d_pos, double *d_vel, double *d_ts, cudaStream_t st)
{
    WsPtrs w = ws_ptrs(const_cast<void *>(d_ws), pop, n_man, knot_cap);
    const long long N = toss_num_samples(p);
    if (N <= 0) return 0;
    resample_kernel<<<pop, kFitThreads, sizeof(long long) * (n_man + 2), st>>>(
        *p, w.knots_t, w.knots_y, w.knots_f, w.seg_start, w.n_seg, n_man, knot_cap, N, d_pos, d_vel, d_ts);
    ctx->launches++;
    TOSS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_dense_eval(toss_ctx *ctx, const double *kt, const double *ky, const double *kf, int n, const double *t,
                      int64_t nt, double *out, cudaStream_t st)
{
    if (nt <= 0) return 0;
    TOSS_REQUIRE(n >= 1, "toss_dense_eval: a segment needs at least one knot");
    dense_eval_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(kt, ky, kf, n, t, nt, out);
    ctx->launches++;
    TOSS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static int fitness_smem_bytes(const toss_ctx *ctx, const GridArgs &g, int n_man, size_t *out)
{
    const size_t b = (size_t)((g.n_words + 1) & ~1) * 4 + (size_t)((g.nr + 1) & ~1) * 4 + sizeof(long long) * (n_man + 2) + 16;
    TOSS_REQUIRE(b <= (size_t)ctx->max_smem_optin - 1024,
                 "voxel grid too fine for the shared-memory bitmap (nr*ntheta*nphi bits must fit in ~220 KB)");
    *out = b;
    return 0;
}

int launch_fitness_positions(toss_ctx *ctx, int fn, const toss_params *p, const double *d_pos, int batch, int64_t N,
                             const double *d_ts, double *d_fit, int64_t *d_nvis, uint32_t *d_bits_out,
                             cudaStream_t st)
{
    TOSS_REQUIRE(fn >= 1 && fn <= 9, "toss_fitness_eval: unknown FitnessFunctions value");
    TOSS_REQUIRE(ctx->d_grid != nullptr, "toss_fitness_eval: no voxel grid uploaded");
    TOSS_REQUIRE(N >= 1 && batch >= 1, "toss_fitness_eval: empty input");
    GridArgs g = grid_args(ctx);
    size_t smem = 0;
    if (int rc = fitness_smem_bytes(ctx, g, 0, &smem)) return rc;
    TOSS_CHECK_CUDA(cudaFuncSetAttribute(fitness_positions_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
    fitness_positions_kernel<<<batch, kFitThreads, smem, st>>>(fn, *p, unit_axis_host(p->spin_axis), g, d_pos, N, d_ts,
                                                               d_fit, d_nvis, d_bits_out);
    ctx->launches++;
    TOSS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_fitness_knots(toss_ctx *ctx, const toss_params *p, const void *d_ws, int pop, int n_man, int knot_cap,
                         double *d_fit, int32_t *d_coll, toss_counters *d_cnt, int64_t *d_nvis, cudaStream_t st)
{
    TOSS_REQUIRE(ctx->d_grid != nullptr, "toss_population_fitness: no voxel grid uploaded");
    WsPtrs w = ws_ptrs(const_cast<void *>(d_ws), pop, n_man, knot_cap);
    GridArgs g = grid_args(ctx);
    const long long N = toss_num_samples(p);
    TOSS_REQUIRE(N >= 1, "toss_population_fitness: no samples (final_time <= start_time?)");
    size_t smem = 0;
    if (int rc = fitness_smem_bytes(ctx, g, n_man, &smem)) return rc;
    TOSS_CHECK_CUDA(cudaFuncSetAttribute(fitness_knots_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fitness_knots_kernel<<<pop, kFitThreads, smem, st>>>(*p, unit_axis_host(p->spin_axis), g, w.knots_t, w.knots_y,
                                                         w.knots_f, w.seg_start, w.n_seg, w.collision, w.counters,
                                                         n_man, knot_cap, N, d_fit, d_nvis);
    ctx->launches++;
    TOSS_CHECK_CUDA(cudaGetLastError());
    if (d_coll) TOSS_CHECK_CUDA(cudaMemcpyAsync(d_coll, w.collision, sizeof(int32_t) * pop, cudaMemcpyDeviceToDevice, st));
    if (d_cnt)
        TOSS_CHECK_CUDA(cudaMemcpyAsync(d_cnt, w.counters, sizeof(toss_counters) * pop, cudaMemcpyDeviceToDevice, st));
    return 0;
}
