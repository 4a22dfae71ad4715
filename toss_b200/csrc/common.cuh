// toss_b200/csrc/common.cuh -- shared declarations of libtoss_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/toss_b200.h"
#include "tmath.cuh"

// ------------------------------------------------------------------------------------------------
// pair record: two faces that share an edge, 43 doubles (SURVEY.md App. A.1 constants that do not depend on the
// field point).  A = (v0, v1, v2), B = (v1, v0, v3), each in the orientation the mesh gives it; edges
// e0 = v0v1 (shared), e1 = v1v2, e2 = v2v0, e3 = v0v3, e4 = v3v1.
//   0..11  v0 v1 v2 v3
//   12..14 N_A      15..17 N_B           unit outward plane normals
//   18..26 nA0 nA1 nA2                   A's unit in-plane outward edge normals (edges e0 e1 e2)
//   27..35 nB0 nB1 nB2                   B's (edges e0 e3 e4)
//   36..40 |G0| .. |G4|                  edge lengths
//   41 2*area(A)    42 2*area(B)         (r0.(r1 x r2) = (N.r0) * 2A)
// A face without a partner has a null B (v3 = v2, B's constants 0: S_B = 0 exactly); padding records sit at
// kPadVertex with every constant 0 and contribute exactly 0.
// The ray caster (mesh_utility.py:132-166) reads the faces in their ORIGINAL vertex order from a separate
// [9][Fpad32] table.
// ------------------------------------------------------------------------------------------------
constexpr int kPairFields = 43;
constexpr int kPairAosStride = 44;      // AoS pair record padded to 352 B = 22 x 16 B (LDS.128 broadcast reads)
constexpr int kPairAosTile = 32;        // K1 (thread-per-point) TMA tile: 32 pairs x 352 B = 11 264 B
#ifndef PAIR_TILE
#define PAIR_TILE 128
#endif
constexpr int kPairTile = PAIR_TILE;    // K2 (lane-per-pair) TMA tile: 22 x 128 x 16 B = 45 056 B, 4 rows of 32 pairs
// Lane-per-pair tables store the 43 fields as 22 double2 SLOTS (slot j = fields 2j, 2j+1; the 44th double is 0):
// [22][n] double2, so a lane fetches two fields with one 16-byte load (LDS.128 / LDG.128, conflict-free: consecutive
// lanes read consecutive 16-byte words) -- 22 load instructions per pair instead of 43.
constexpr int kPairSlots = 22;
constexpr int kPairTileDoubles = 2 * kPairSlots * kPairTile;

// field accessor of a slot table: col points at this lane's double2 of slot 0, `stride` = pairs per slot row
template <bool kLdg>
struct PairCols {
    const double2 *col;
    int stride;
    __device__ __forceinline__ double operator()(int k) const
    {
        const double2 *p = col + (size_t)(k >> 1) * stride;
        const double2 v = kLdg ? __ldg(p) : *p;
        return (k & 1) ? v.y : v.x;
    }
};
constexpr int kRayFields = 9;
constexpr int kProxyOcc = 16;           // the proxy's occupancy grid of the body: kProxyOcc^3 cells over the bounding box
constexpr int kProxyTail = 6 + kProxyOcc * kProxyOcc * kProxyOcc / 64;   // doubles after the mascons: lo | cells/m | bits
constexpr double kLaneFarEdge = 400.0;   // K1a decides the far tier per lane on meshes whose longest edge is shorter [m]
constexpr double kPadVertex = 1.0e9;    // padding faces sit here with N = n = |G| = 2A = 0 -> contribute exactly 0
constexpr double kGravConst = 6.67430e-11;   // CODATA-2018, as polyhedral_gravity >= 2

struct toss_ctx {
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    // mesh
    int V = 0, F = 0;
    int Fpad32 = 0;        // F rounded up to 32
    double Grho = 0.0;
    double GM = 0.0, com[3] = {0.0, 0.0, 0.0};   // G rho V and the centre of mass
    double *d_mascons = nullptr;                 // [n_mascons][4] x y z G*m: proxy field of the work-queue cost predictor
    int n_mascons = 0;
    double mascon_soft2 = 0.0;
    double max_edge = 0.0;                       // longest edge of the mesh [m] (chooses the thread-per-point kernel's tier rule)
    int n_pairs = 0;                 // pair records (matched pairs + faces left single)
    int Ppad32 = 0;                  // n_pairs rounded up to 32 (flat SoA)
    int n_pair_aos_tiles = 0;        // ceil(n_pairs / 32)
    int n_pair_tiles = 0;            // ceil(n_pairs / 128)
    double *d_pair_aos = nullptr;    // [n_pair_aos_tiles*32][44]      K1a
    double *d_pair_soa = nullptr;    // [22][Ppad32] double2 slots      K1b, resident stepper
    double *d_pair_tiles = nullptr;  // [n_pair_tiles][22][128] double2 streaming stepper
    double *d_ray_soa = nullptr;     // [9][Fpad32] v0 v1 v2 of every face in mesh order: ray casting
    int32_t *h_pairs = nullptr;      // [n_pairs][3] fA, qA, fB: the pairing (toss_mesh_pairs)
    // voxel grid
    int nr = 0, nth = 0, nph = 0;
    double *d_grid = nullptr;        // r | theta | phi
    uint32_t *d_bool_bits = nullptr; // packed bool_tensor, voxel (i,j,k) -> bit (i*nth + j)*nph + k
    uint8_t *h_bool = nullptr;       // host copy (nr*nth*nph bytes)
    // internal buffers of the *_host entry points
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0;
    void *d_tensor_scratch = nullptr;   // toss_update_tensor: [fitness | visited bits]
    size_t tensor_scratch_bytes = 0;
    void *h_pinned = nullptr;
    size_t pinned_bytes = 0;
    int64_t launches = 0;
    int queue_order_mode = 0;        // 0: the launcher orders the work queue itself (proxy run); 1: rows are pre-ordered
    int32_t *queue_scratch = nullptr;   // 256 B: work-queue cursor of toss_predict_cost's proxy run
    double *d_gaco_tab = nullptr;       // toss_gaco_sample: archive | cumulative weights | sigma | lb | ub
    size_t gaco_tab_words = 0;
};

void toss_set_error(const std::string &msg);

// Every ABI entry point runs on the context's device and leaves the caller's current device as it found it
// (a Context on cuda:1 must not silently change torch's current device).
struct DeviceScope {
    int prev = -1;
    cudaError_t err;
    explicit DeviceScope(int device)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
        else if (err == cudaSuccess) prev = -1;   // nothing to restore
    }
    ~DeviceScope()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
#define TOSS_ENTER(ctx)                                                                                \
    DeviceScope _toss_dev((ctx)->device);                                                              \
    TOSS_CHECK_CUDA(_toss_dev.err)

// NVTX ranges around the phases of a population evaluation (proxy run / queue sort / stepper / resample + fitness):
// visible to nsys / ncu --nvtx, free when no tool is attached (header-only NVTX v3).
#include <nvtx3/nvToolsExt.h>
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

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
// the same with sin / cos of the angle -(w t) already in hand (the stepper computes them early, next to the stage point)
__device__ __forceinline__ void rotate_point_sc(double s, double c, double x, double y, double z, double kx, double ky,
                                                double kz, double &ox, double &oy, double &oz);
__device__ __forceinline__ void rotate_point_dev(double t, double x, double y, double z, double kx, double ky,
                                                 double kz, double w, double &ox, double &oy, double &oz)
{
    double th = -mul_(w, t);
    double s, c;
    tm_sincos(th, s, c);
    rotate_point_sc(s, c, x, y, z, kx, ky, kz, ox, oy, oz);
}
__device__ __forceinline__ void rotate_point_sc(double s, double c, double x, double y, double z, double kx, double ky,
                                                double kz, double &ox, double &oy, double &oz)
{
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

// ---- face pairing (pairs.cu) -------------------------------------------------------------------------
#include <vector>
void toss_build_face_pairs(const int32_t *faces, int F, int V, std::vector<int32_t> &out, int *n_unmatched = nullptr,
                           int *first_unmatched = nullptr);
int launch_mesh_tables(toss_ctx *ctx, const double *h_verts, int V, const int32_t *h_faces, int F, const int32_t *h_pairs,
                       int n_pairs, int *bad_face);

// ---- kernels' host launchers (defined in the .cu files) -------------------------------------------
int launch_gravity(toss_ctx *ctx, const double *d_pts, int64_t n, double *d_acc, double *d_pot, cudaStream_t st);
int launch_compute_motion(toss_ctx *ctx, const toss_params *p, const double *d_t, const double *d_y, int64_t n,
                          double *d_out, cudaStream_t st);
int launch_propagate(toss_ctx *ctx, const toss_params *p, const double *d_x, int pop, int dim, const double *h_fixed,
                     int n_fixed, int n_man, void *d_ws, int knot_cap, cudaStream_t st);
int launch_predict_cost(toss_ctx *ctx, const toss_params *p, const double *d_x, int pop, int dim, const double *h_fixed,
                        int n_fixed, int n_man, int32_t *d_cost, cudaStream_t st);
int launch_cost_order(toss_ctx *ctx, const int32_t *d_cost, int n, int32_t *d_order, cudaStream_t st);
int launch_gaco_sample(toss_ctx *ctx, const double *h_archive_x, int ker, int dim, const double *h_cum,
                       const double *h_sigma, const double *h_lb, const double *h_ub, uint64_t seed, uint64_t generation,
                       int first_ant, int n, double *d_x_out, cudaStream_t st);
int launch_deal_rows(toss_ctx *ctx, const double *d_x, int n, int dim, const int32_t *d_order, int rank, int world,
                     double *d_x_local, cudaStream_t st);
int launch_pack_result(toss_ctx *ctx, const double *d_fit, const toss_counters *d_cnt, int n_local, int cap,
                       double *d_send, cudaStream_t st);
int launch_unpack_result(toss_ctx *ctx, const double *d_recv, const int32_t *d_order, int n, int world, int cap,
                         double *d_fitness, double *d_flags, cudaStream_t st);
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
