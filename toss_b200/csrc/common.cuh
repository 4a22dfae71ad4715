// toss_b200/csrc/common.cuh -- shared declarations of libtoss_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/toss_b200.h"
#include "tmath.cuh"

// ------------------------------------------------------------------------------------------------
// face record: 25 doubles per face (SURVEY.md App. A.1 constants that do not depend on the field point)
//   0..8   v0 v1 v2          vertices (also what the ray caster reads, mesh_utility.py:132-166)
//   9..11  N                 unit outward plane normal
//   12..20 n0 n1 n2          unit in-plane outward edge normals
//   21..23 |G0| |G1| |G2|    edge lengths
//   24     2*area            |G0 x G1|   (r0.(r1 x r2) = (N.r0) * 2A)
// ------------------------------------------------------------------------------------------------
constexpr int kFaceFields = 25;
constexpr int kAosStride = 26;          // AoS record padded to 208 B = 13 x 16 B (LDS.128 broadcast reads)
constexpr int kAosTileFaces = 64;       // K1 (thread-per-point) TMA tile: 64 x 208 B = 13 312 B
#ifndef SOA_TILE_FACES
#define SOA_TILE_FACES 128
#endif
constexpr int kSoaTileFaces = SOA_TILE_FACES;   // K2 (lane-per-face) TMA tile: 25 x 128 x 8 B = 25 600 B
constexpr int kSoaTileDoubles = kFaceFields * kSoaTileFaces;
constexpr double kPadVertex = 1.0e9;    // padding faces sit here with N = n = |G| = 2A = 0 -> contribute exactly 0
constexpr double kGravConst = 6.67430e-11;   // CODATA-2018, as polyhedral_gravity >= 2

struct toss_ctx {
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    // mesh
    int V = 0, F = 0;
    int Fpad32 = 0;        // F rounded up to 32  (flat SoA, resident mode)
    int n_aos_tiles = 0;   // ceil(F / 64)
    int n_soa_tiles = 0;   // ceil(F / 128)
    double Grho = 0.0;
    double GM = 0.0, com[3] = {0.0, 0.0, 0.0};   // G rho V and the centre of mass
    double *d_mascons = nullptr;                 // [n_mascons][4] x y z G*m: proxy field of the work-queue cost predictor
    int n_mascons = 0;
    double mascon_soft2 = 0.0;
    double *d_face_aos = nullptr;    // [n_aos_tiles*64][26]
    double *d_face_soa = nullptr;    // [25][Fpad32]
    double *d_face_tiles = nullptr;  // [n_soa_tiles][25][128]
    // voxel grid
    int nr = 0, nth = 0, nph = 0;
    double *d_grid = nullptr;        // r | theta | phi
    uint32_t *d_bool_bits = nullptr; // packed bool_tensor, voxel (i,j,k) -> bit (i*nth + j)*nph + k
    uint8_t *h_bool = nullptr;       // host copy (nr*nth*nph bytes)
    // internal buffers of the *_host entry points
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0;
    void *h_pinned = nullptr;
    size_t pinned_bytes = 0;
    int64_t launches = 0;
};

void toss_set_error(const std::string &msg);

#define TOSS_CHECK_CUDA(expr)                                                                          \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            toss_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                        \
            return 1;                                                                                  \
        }                                                                                              \
    } while (0)

#define TOSS_REQUIRE(cond, msg)                                                                        \
    do {                                                                                               \
        if (!(cond)) {                                                                                 \
            toss_set_error(std::string(msg));                                                          \
            return 2;                                                                                  \
        }                                                                                              \
    } while (0)

// ------------------------------------------------------------------------------------------------
// un-contracted IEEE arithmetic: nvcc fuses a*b+c into FMA by default; wherever an INTEGER decision
// (ray parity, voxel index, step accept/reject, sample->interval) hangs on the value we spell the
// roundings out so they are the ones numpy (un-fused, left to right) performs.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double mul_(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dot3_(double ax, double ay, double az, double bx, double by, double bz)
{
    return add_(add_(mul_(ax, bx), mul_(ay, by)), mul_(az, bz));
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// equations_of_motion.py:98-117: Quaternion(axis, angle = -(w t)).rotate(x), axis normalised
// (Euler-Rodrigues form, tests/point_rotation_test.py:47-76).  k = unit axis.
__device__ __forceinline__ void rotate_point_dev(double t, double x, double y, double z, double kx, double ky,
                                                 double kz, double w, double &ox, double &oy, double &oz)
{
    double th = -mul_(w, t);
    double s, c;
    tm_sincos(th, s, c);
    double cx = sub_(mul_(ky, z), mul_(kz, y));
    double cy = sub_(mul_(kz, x), mul_(kx, z));
    double cz = sub_(mul_(kx, y), mul_(ky, x));
    double kd = mul_(dot3_(kx, ky, kz, x, y, z), sub_(1.0, c));
    ox = add_(add_(mul_(x, c), mul_(cx, s)), mul_(kx, kd));
    oy = add_(add_(mul_(y, c), mul_(cy, s)), mul_(ky, kd));
    oz = add_(add_(mul_(z, c), mul_(cz, s)), mul_(kz, kd));
}

struct UnitAxis {
    double x, y, z;
};
inline UnitAxis unit_axis_host(const double a[3])
{
    double n = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    return UnitAxis{a[0] / n, a[1] / n, a[2] / n};
}

// ---- kernels' host launchers (defined in the .cu files) -------------------------------------------
int launch_gravity(toss_ctx *ctx, const double *d_pts, int64_t n, double *d_acc, double *d_pot, cudaStream_t st);
int launch_compute_motion(toss_ctx *ctx, const toss_params *p, const double *d_t, const double *d_y, int64_t n,
                          double *d_out, cudaStream_t st);
int launch_propagate(toss_ctx *ctx, const toss_params *p, const double *d_x, int pop, int dim, const double *h_fixed,
                     int n_fixed, int n_man, void *d_ws, int knot_cap, cudaStream_t st);
int launch_collision(toss_ctx *ctx, const toss_params *p, void *d_ws, int pop, int n_man, int knot_cap,
                     cudaStream_t st);
int launch_resample(toss_ctx *ctx, const toss_params *p, const void *d_ws, int pop, int n_man, int knot_cap,
                    double *d_pos, double *d_vel, double *d_ts, cudaStream_t st);
int launch_dense_eval(toss_ctx *ctx, const double *kt, const double *ky, const double *kf, int n, const double *t,
                      int64_t nt, double *out, cudaStream_t st);
int launch_fitness_positions(toss_ctx *ctx, int fn, const toss_params *p, const double *d_pos, int batch, int64_t N,
                             const double *d_ts, double *d_fit, int64_t *d_nvis, uint32_t *d_bits_out,
                             cudaStream_t st);
int launch_fitness_knots(toss_ctx *ctx, const toss_params *p, const void *d_ws, int pop, int n_man, int knot_cap,
                         double *d_fit, int32_t *d_coll, toss_counters *d_cnt, int64_t *d_nvis, cudaStream_t st);
int launch_is_outside(toss_ctx *ctx, const double *d_pts, int64_t n, uint8_t *d_out, cudaStream_t st);
int launch_fp64_peak(toss_ctx *ctx, double *tflops);
int launch_math_probe(toss_ctx *ctx, int fn, const double *a, const double *b, int64_t n, double *out, double *out2,
                      cudaStream_t st);
int launch_div_sqrt_selftest(toss_ctx *ctx, int64_t samples, int64_t *bad_div, int64_t *bad_sqrt);
int launch_rotate_points(toss_ctx *ctx, const double *h_axis3, double w, const double *d_t, const double *d_x,
                         int64_t n, double *d_out, cudaStream_t st);
