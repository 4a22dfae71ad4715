// toss_b200/csrc/api.cu -- the C ABI of include/toss_b200.h: context, mesh / grid upload, entry points.
#include <math.h>
#include <string.h>

#include <vector>

#include "common.cuh"

int toss_workspace_layout_impl(int pop, int n_man, int knot_cap, toss_ws_layout *out);

static thread_local std::string g_last_error;
void toss_set_error(const std::string &msg) { g_last_error = msg; }

extern "C" {

const char *toss_version(void) { return "toss_b200 0.1 (sm_100a)"; }
const char *toss_last_error(void) { return g_last_error.c_str(); }

int toss_ctx_create(int device, toss_ctx **out)
{
    TOSS_REQUIRE(out != nullptr, "toss_ctx_create: out is NULL");
    int n = 0;
    TOSS_CHECK_CUDA(cudaGetDeviceCount(&n));
    TOSS_REQUIRE(device >= 0 && device < n, "toss_ctx_create: no such CUDA device");
    cudaDeviceProp prop;
    TOSS_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    TOSS_REQUIRE(prop.major == 10, "toss_ctx_create: this library is built for sm_100a (B200) only");
    toss_ctx *c = new toss_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    *out = c;
    return 0;
}

static void free_mesh(toss_ctx *c)
{
    cudaFree(c->d_pair_aos);
    cudaFree(c->d_pair_soa);
    cudaFree(c->d_pair_tiles);
    cudaFree(c->d_ray_soa);
    cudaFree(c->d_mascons);
    free(c->h_pairs);
    c->d_pair_aos = c->d_pair_soa = c->d_pair_tiles = c->d_ray_soa = c->d_mascons = nullptr;
    c->h_pairs = nullptr;
    c->n_mascons = 0;
    c->n_pairs = 0;
}
static void free_grid(toss_ctx *c)
{
    cudaFree(c->d_grid);
    cudaFree(c->d_bool_bits);
    free(c->h_bool);
    c->d_grid = nullptr;
    c->d_bool_bits = nullptr;
    c->h_bool = nullptr;
}

int toss_ctx_destroy(toss_ctx *ctx)
{
    if (!ctx) return 0;
    DeviceScope dev(ctx->device);
    free_mesh(ctx);
    free_grid(ctx);
    cudaFree(ctx->d_scratch);
    cudaFree(ctx->d_tensor_scratch);
    cudaFree(ctx->queue_scratch);
    cudaFree(ctx->d_gaco_tab);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    delete ctx;
    return 0;
}

int toss_mesh_num_faces(const toss_ctx *ctx) { return ctx ? ctx->F : 0; }
int toss_mesh_num_pairs(const toss_ctx *ctx) { return ctx ? ctx->n_pairs : 0; }
int toss_mesh_pairs(const toss_ctx *ctx, int32_t *out)
{
    TOSS_REQUIRE(ctx && out && ctx->h_pairs, "toss_mesh_pairs: no mesh uploaded");
    memcpy(out, ctx->h_pairs, sizeof(int32_t) * 3 * (size_t)ctx->n_pairs);
    return 0;
}
int toss_mesh_pair_records(toss_ctx *ctx, double *h_out)
{
    TOSS_REQUIRE(ctx && h_out && ctx->d_pair_aos, "toss_mesh_pair_records: no mesh uploaded");
    TOSS_ENTER(ctx);
    TOSS_CHECK_CUDA(cudaMemcpy(h_out, ctx->d_pair_aos, sizeof(double) * kPairAosStride * (size_t)ctx->n_pairs,
                               cudaMemcpyDeviceToHost));
    return 0;
}
// the matching alone (host integer work, no device): out = int32[3 * F] capacity, returns the number of records
int toss_pair_faces(const int32_t *h_faces, int F, int V, int32_t *out)
{
    if (!h_faces || !out || F <= 0 || V <= 0) return -1;
    for (int i = 0; i < 3 * F; ++i)
        if (h_faces[i] < 0 || h_faces[i] >= V) return -1;
    std::vector<int32_t> pairs;
    toss_build_face_pairs(h_faces, F, V, pairs);
    memcpy(out, pairs.data(), sizeof(int32_t) * pairs.size());
    return (int)(pairs.size() / 3);
}
int64_t toss_launch_count(const toss_ctx *ctx) { return ctx ? ctx->launches : 0; }

// Per-face constants of SURVEY.md App. A.1 (what GravityEvaluable precomputes at construction,
// setup_parameters.py:63): N = G0 x G1 / |.|, n_q = G_q x N / |.|, |G_q|, 2A = |G0 x G1|.
int toss_mesh_upload(toss_ctx *ctx, const double *h_verts, int V, const int32_t *h_faces, int F, double density)
{
    TOSS_REQUIRE(ctx && h_verts && h_faces, "toss_mesh_upload: NULL argument");
    TOSS_REQUIRE(V >= 4 && F >= 4, "toss_mesh_upload: a closed mesh needs at least 4 vertices and 4 faces");
    TOSS_ENTER(ctx);
    for (int i = 0; i < 3 * F; ++i)
        TOSS_REQUIRE(h_faces[i] >= 0 && h_faces[i] < V, "toss_mesh_upload: face index out of range");
    // the pairing (host, integer work).  The sum runs over faces that share edges: the surface must be closed and
    // consistently oriented -- every edge used by exactly two faces, in opposite directions (polyhedral_gravity
    // requires the same of its input and silently returns garbage otherwise).
    std::vector<int32_t> pairs;
    int n_unmatched = 0, first_unmatched = 0;
    toss_build_face_pairs(h_faces, F, V, pairs, &n_unmatched, &first_unmatched);
    TOSS_REQUIRE(n_unmatched == 0,
                 "toss_mesh_upload: the mesh is not a closed, consistently oriented surface: " +
                     std::to_string(n_unmatched) + " face edge(s) without an opposite partner (first: face " +
                     std::to_string(first_unmatched / 3) + ", edge " + std::to_string(first_unmatched % 3) + ")");
    const int n_pairs = (int)(pairs.size() / 3);
    {   // outward orientation: positive volume (divergence theorem)
        double vol6 = 0.0;
        for (int f = 0; f < F; ++f) {
            const double *a = h_verts + 3 * (size_t)h_faces[3 * f], *b = h_verts + 3 * (size_t)h_faces[3 * f + 1],
                         *d = h_verts + 3 * (size_t)h_faces[3 * f + 2];
            vol6 += a[0] * (b[1] * d[2] - b[2] * d[1]) + a[1] * (b[2] * d[0] - b[0] * d[2]) +
                    a[2] * (b[0] * d[1] - b[1] * d[0]);
        }
        TOSS_REQUIRE(vol6 > 0.0, "toss_mesh_upload: the faces are wound clockwise seen from outside (negative volume): "
                                 "flip the vertex order of every face");
    }
    free_mesh(ctx);
    ctx->V = V;
    ctx->F = F;
    ctx->Grho = kGravConst * density;
    {   // volume and centre of mass by the divergence theorem (signed tetrahedra from the origin)
        double vol = 0.0, c[3] = {0.0, 0.0, 0.0};
        for (int f = 0; f < F; ++f) {
            const double *a = h_verts + 3 * (size_t)h_faces[3 * f], *b = h_verts + 3 * (size_t)h_faces[3 * f + 1],
                         *d = h_verts + 3 * (size_t)h_faces[3 * f + 2];
            const double v6 = a[0] * (b[1] * d[2] - b[2] * d[1]) + a[1] * (b[2] * d[0] - b[0] * d[2]) +
                              a[2] * (b[0] * d[1] - b[1] * d[0]);
            vol += v6 / 6.0;
            for (int k = 0; k < 3; ++k) c[k] += (a[k] + b[k] + d[k]) / 4.0 * v6 / 6.0;
        }
        ctx->GM = ctx->Grho * vol;
        for (int k = 0; k < 3; ++k) ctx->com[k] = vol != 0.0 ? c[k] / vol : 0.0;
        // mascon proxy of the body (only used to ORDER the work queue): centres of an 8^3 grid over the bounding box
        // that lie inside the mesh (+z ray parity), equal masses
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int v = 0; v < V; ++v)
            for (int k = 0; k < 3; ++k) {
                lo[k] = h_verts[3 * v + k] < lo[k] ? h_verts[3 * v + k] : lo[k];
                hi[k] = h_verts[3 * v + k] > hi[k] ? h_verts[3 * v + k] : hi[k];
            }
        const int G = 8;
        std::vector<double> mc;
        for (int i = 0; i < G; ++i)
            for (int j = 0; j < G; ++j)
                for (int k = 0; k < G; ++k) {
                    const double o[3] = {lo[0] + (i + 0.5) * (hi[0] - lo[0]) / G, lo[1] + (j + 0.5) * (hi[1] - lo[1]) / G,
                                         lo[2] + (k + 0.5) * (hi[2] - lo[2]) / G};
                    int hits = 0;
                    for (int f = 0; f < F; ++f) {
                        const double *a = h_verts + 3 * (size_t)h_faces[3 * f], *b = h_verts + 3 * (size_t)h_faces[3 * f + 1],
                                     *d = h_verts + 3 * (size_t)h_faces[3 * f + 2];
                        const double e1x = b[0] - a[0], e1y = b[1] - a[1], e1z = b[2] - a[2];
                        const double e2x = d[0] - a[0], e2y = d[1] - a[1], e2z = d[2] - a[2];
                        const double hx = -e2y, hy = e2x, det = e1x * hx + e1y * hy;
                        if (det < 1e-6 && det > -1e-6) continue;
                        const double inv = 1.0 / det, sx = o[0] - a[0], sy = o[1] - a[1], sz = o[2] - a[2];
                        const double u = (sx * hx + sy * hy) * inv;
                        if (u < 0.0 || u > 1.0) continue;
                        const double q0 = sy * e1z - sz * e1y, q1 = sz * e1x - sx * e1z, q2 = sx * e1y - sy * e1x;
                        const double w = q2 * inv;
                        if (w < 0.0 || u + w > 1.0) continue;
                        if (inv * (q0 * e2x + q1 * e2y + q2 * e2z) > 0.0) ++hits;
                    }
                    if (hits & 1) {
                        mc.push_back(o[0]);
                        mc.push_back(o[1]);
                        mc.push_back(o[2]);
                        mc.push_back(0.0);
                    }
                }
        cudaFree(ctx->d_mascons);
        ctx->d_mascons = nullptr;
        ctx->n_mascons = (int)(mc.size() / 4);
        if (ctx->n_mascons > 0) {
            for (int m = 0; m < ctx->n_mascons; ++m) mc[4 * m + 3] = ctx->GM / ctx->n_mascons;
            const double cell = ((hi[0] - lo[0]) + (hi[1] - lo[1]) + (hi[2] - lo[2])) / (3.0 * G);
            ctx->mascon_soft2 = 0.25 * cell * cell;
            // The proxy's idea of "inside the body" (it retires a trajectory there when the real run would retire it at
            // its first event point inside the mesh, so that the predicted cost of a collider is short like its real
            // one): occupancy of a kProxyOcc^3 grid over the bounding box -- a cell is inside when its centre is.  One
            // vertical line per (i, j) column: the z of its crossings with the faces, sorted; a centre is inside when an
            // odd number of crossings lies above it.  Appended to the mascon table: lo[3] | cells per metre[3] | bits.
            const int Q = kProxyOcc;
            std::vector<uint64_t> bits((size_t)Q * Q * Q / 64, 0ull);
            std::vector<double> zs;
            for (int i = 0; i < Q; ++i)
                for (int j = 0; j < Q; ++j) {
                    const double ox = lo[0] + (i + 0.5) * (hi[0] - lo[0]) / Q, oy = lo[1] + (j + 0.5) * (hi[1] - lo[1]) / Q;
                    zs.clear();
                    for (int f = 0; f < F; ++f) {
                        const double *a = h_verts + 3 * (size_t)h_faces[3 * f], *b = h_verts + 3 * (size_t)h_faces[3 * f + 1],
                                     *d = h_verts + 3 * (size_t)h_faces[3 * f + 2];
                        const double e1x = b[0] - a[0], e1y = b[1] - a[1], e2x = d[0] - a[0], e2y = d[1] - a[1];
                        const double det = e1x * e2y - e1y * e2x;
                        if (det == 0.0) continue;
                        const double sx = ox - a[0], sy = oy - a[1];
                        const double u = (sx * e2y - sy * e2x) / det, w = (e1x * sy - e1y * sx) / det;
                        if (u < 0.0 || w < 0.0 || u + w > 1.0) continue;
                        zs.push_back(a[2] + u * (b[2] - a[2]) + w * (d[2] - a[2]));
                    }
                    for (int k = 0; k < Q; ++k) {
                        const double oz = lo[2] + (k + 0.5) * (hi[2] - lo[2]) / Q;
                        int above = 0;
                        for (double z : zs) above += z > oz;
                        if (above & 1) {
                            const int c = (i * Q + j) * Q + k;
                            bits[c >> 6] |= 1ull << (c & 63);
                        }
                    }
                }
            {   // Erode by one cell: only cells whose 26 neighbours are inside too count.  A trajectory that merely passes
                // close must NOT be predicted short -- it would start last and run longest; a collider missed here only
                // keeps its long prediction and retires early, which costs nothing.
                std::vector<uint64_t> core(bits.size(), 0ull);
                auto in = [&](int i, int j, int k) {
                    if (i < 0 || j < 0 || k < 0 || i >= Q || j >= Q || k >= Q) return false;
                    const int c = (i * Q + j) * Q + k;
                    return ((bits[c >> 6] >> (c & 63)) & 1ull) != 0;
                };
                for (int i = 0; i < Q; ++i)
                    for (int j = 0; j < Q; ++j)
                        for (int k = 0; k < Q; ++k) {
                            bool all = true;
                            for (int d = 0; d < 27 && all; ++d) all = in(i + d / 9 - 1, j + (d / 3) % 3 - 1, k + d % 3 - 1);
                            if (all) {
                                const int c = (i * Q + j) * Q + k;
                                core[c >> 6] |= 1ull << (c & 63);
                            }
                        }
                bits.swap(core);
            }
            for (int k = 0; k < 3; ++k) mc.push_back(lo[k]);
            for (int k = 0; k < 3; ++k) mc.push_back(hi[k] > lo[k] ? Q / (hi[k] - lo[k]) : 0.0);
            for (uint64_t wbits : bits) {
                double dv;
                memcpy(&dv, &wbits, 8);
                mc.push_back(dv);
            }
            TOSS_CHECK_CUDA(cudaMalloc(&ctx->d_mascons, mc.size() * 8));
            TOSS_CHECK_CUDA(cudaMemcpy(ctx->d_mascons, mc.data(), mc.size() * 8, cudaMemcpyHostToDevice));
        }
    }
    ctx->max_edge = 0.0;
    for (int f = 0; f < F; ++f)
        for (int q = 0; q < 3; ++q) {
            const double *a = h_verts + 3 * (size_t)h_faces[3 * f + q], *b = h_verts + 3 * (size_t)h_faces[3 * f + (q + 1) % 3];
            const double e = sqrt((b[0] - a[0]) * (b[0] - a[0]) + (b[1] - a[1]) * (b[1] - a[1]) + (b[2] - a[2]) * (b[2] - a[2]));
            if (e > ctx->max_edge) ctx->max_edge = e;
        }
    ctx->Fpad32 = (F + 31) / 32 * 32;
    ctx->n_pairs = n_pairs;
    ctx->Ppad32 = (n_pairs + 31) / 32 * 32;
    ctx->n_pair_aos_tiles = (n_pairs + kPairAosTile - 1) / kPairAosTile;
    ctx->n_pair_tiles = (n_pairs + kPairTile - 1) / kPairTile;
    ctx->h_pairs = (int32_t *)malloc(sizeof(int32_t) * pairs.size());
    memcpy(ctx->h_pairs, pairs.data(), sizeof(int32_t) * pairs.size());

    // every device table (three encodings of the pair records + the ray caster's face table) is built ON the device
    // from the vertex / face / pair index arrays (pairs.cu: pair_tables_kernel, ray_table_kernel)
    int bad_face = -1;
    const int rc = launch_mesh_tables(ctx, h_verts, V, h_faces, F, pairs.data(), n_pairs, &bad_face);
    if (rc < 0) {
        free_mesh(ctx);
        return -1;
    }
    if (rc > 0) {
        free_mesh(ctx);
        TOSS_REQUIRE(false, "toss_mesh_upload: degenerate (zero-area) face " + std::to_string(bad_face));
    }
    return 0;
}

static int upload_bool_bits(toss_ctx *ctx)
{
    const long long nvox = (long long)ctx->nr * ctx->nth * ctx->nph;
    const size_t nw = (size_t)((nvox + 31) / 32);
    std::vector<uint32_t> bits(nw, 0u);
    for (long long b = 0; b < nvox; ++b)
        if (ctx->h_bool[b]) bits[b >> 5] |= 1u << (b & 31);
    TOSS_CHECK_CUDA(cudaMemcpy(ctx->d_bool_bits, bits.data(), nw * 4, cudaMemcpyHostToDevice));
    return 0;
}

int toss_grid_upload(toss_ctx *ctx, const double *h_r, int nr, const double *h_theta, int nth, const double *h_phi,
                     int nph, const uint8_t *h_bool_tensor)
{
    TOSS_REQUIRE(ctx && h_r && h_theta && h_phi && h_bool_tensor, "toss_grid_upload: NULL argument");
    TOSS_REQUIRE(nr >= 1 && nth >= 1 && nph >= 1, "toss_grid_upload: empty grid axis");
    TOSS_REQUIRE((long long)nr * nth * nph < (1LL << 31), "toss_grid_upload: grid too large");
    TOSS_ENTER(ctx);
    free_grid(ctx);
    ctx->nr = nr;
    ctx->nth = nth;
    ctx->nph = nph;
    const size_t nvox = (size_t)nr * nth * nph;
    std::vector<double> g;
    g.insert(g.end(), h_r, h_r + nr);
    g.insert(g.end(), h_theta, h_theta + nth);
    g.insert(g.end(), h_phi, h_phi + nph);
    TOSS_CHECK_CUDA(cudaMalloc(&ctx->d_grid, g.size() * 8));
    TOSS_CHECK_CUDA(cudaMemcpy(ctx->d_grid, g.data(), g.size() * 8, cudaMemcpyHostToDevice));
    TOSS_CHECK_CUDA(cudaMalloc(&ctx->d_bool_bits, ((nvox + 31) / 32) * 4));
    ctx->h_bool = (uint8_t *)malloc(nvox);
    memcpy(ctx->h_bool, h_bool_tensor, nvox);
    return upload_bool_bits(ctx);
}

int toss_gravity_eval(toss_ctx *ctx, const double *d_pts, int64_t n, double *d_acc, double *d_pot, uintptr_t stream)
{
    TOSS_REQUIRE(ctx != nullptr, "toss_gravity_eval: NULL context");
    TOSS_REQUIRE(n == 0 || (d_pts && d_acc), "toss_gravity_eval: NULL buffer");
    TOSS_ENTER(ctx);
    return launch_gravity(ctx, d_pts, n, d_acc, d_pot, (cudaStream_t)stream);
}

int toss_compute_motion(toss_ctx *ctx, const toss_params *p, const double *d_t, const double *d_y, int64_t n,
                        double *d_out, uintptr_t stream)
{
    TOSS_REQUIRE(ctx && p, "toss_compute_motion: NULL argument");
    TOSS_ENTER(ctx);
    return launch_compute_motion(ctx, p, d_t, d_y, n, d_out, (cudaStream_t)stream);
}

int toss_workspace_layout(int pop, int n_man, int knot_cap, toss_ws_layout *out)
{
    TOSS_REQUIRE(out && pop > 0 && n_man >= 0 && knot_cap > 0, "toss_workspace_layout: bad argument");
    return toss_workspace_layout_impl(pop, n_man, knot_cap, out);
}

int toss_propagate(toss_ctx *ctx, const toss_params *p, const double *d_x, int pop, int dim, const double *h_fixed,
                   int n_fixed, int n_man, void *d_workspace, int knot_cap, uintptr_t stream)
{
    TOSS_REQUIRE(ctx && p && d_workspace && (d_x || dim == 0), "toss_propagate: NULL argument");
    TOSS_ENTER(ctx);
    // the stepper ray-casts every event point while it sweeps the faces: the verdict is already in the workspace
    return launch_propagate(ctx, p, d_x, pop, dim, h_fixed, n_fixed, n_man, d_workspace, knot_cap,
                            (cudaStream_t)stream);
}

int toss_integrate(toss_ctx *ctx, const toss_params *p, const double *d_x, int pop, int dim, const double *h_fixed,
                   int n_fixed, int n_man, void *d_workspace, int knot_cap, uintptr_t stream)
{
    TOSS_REQUIRE(ctx && p && d_workspace && (d_x || dim == 0), "toss_integrate: NULL argument");
    TOSS_ENTER(ctx);
    return launch_propagate(ctx, p, d_x, pop, dim, h_fixed, n_fixed, n_man, d_workspace, knot_cap, (cudaStream_t)stream);
}

int toss_collision_verdict(toss_ctx *ctx, const toss_params *p, void *d_workspace, int pop, int n_man, int knot_cap,
                           uintptr_t stream)
{
    TOSS_REQUIRE(ctx && p && d_workspace, "toss_collision_verdict: NULL argument");
    TOSS_ENTER(ctx);
    return launch_collision(ctx, p, d_workspace, pop, n_man, knot_cap, (cudaStream_t)stream);
}

int toss_rotate_points(toss_ctx *ctx, const double *h_axis3, double spin_velocity, const double *d_t,
                       const double *d_x, int64_t n, double *d_out, uintptr_t stream)
{
    TOSS_REQUIRE(ctx && h_axis3 && (n == 0 || (d_t && d_x && d_out)), "toss_rotate_points: NULL argument");
    TOSS_ENTER(ctx);
    return launch_rotate_points(ctx, h_axis3, spin_velocity, d_t, d_x, n, d_out, (cudaStream_t)stream);
}

int64_t toss_num_samples(const toss_params *p)
{   // len(np.arange(start, final, period))  (trajectory_tools.py:46)
    const double n = ceil((p->final_time - p->start_time) / p->measurement_period);
    return n > 0 ? (int64_t)n : 0;
}

int toss_resample(toss_ctx *ctx, const toss_params *p, const void *d_workspace, int pop, int n_man, int knot_cap,
                  double *d_pos, double *d_vel, double *d_ts, uintptr_t stream)
{
    TOSS_REQUIRE(ctx && p && d_workspace && d_pos, "toss_resample: NULL argument");
    TOSS_ENTER(ctx);
    return launch_resample(ctx, p, d_workspace, pop, n_man, knot_cap, d_pos, d_vel, d_ts, (cudaStream_t)stream);
}

int toss_dense_eval(toss_ctx *ctx, const double *d_kt, const double *d_ky, const double *d_kf, int n_knots,
                    const double *d_t, int64_t nt, double *d_out, uintptr_t stream)
{
    TOSS_REQUIRE(ctx && d_kt && d_ky && d_kf && d_t && d_out, "toss_dense_eval: NULL argument");
    TOSS_ENTER(ctx);
    return launch_dense_eval(ctx, d_kt, d_ky, d_kf, n_knots, d_t, nt, d_out, (cudaStream_t)stream);
}

int toss_fitness_eval(toss_ctx *ctx, int fn, const toss_params *p, const double *d_pos, int batch, int64_t N,
                      const double *d_ts, double *d_fitness, int64_t *d_n_visited, uintptr_t stream)
{
    TOSS_REQUIRE(ctx && p && d_pos && d_ts && d_fitness, "toss_fitness_eval: NULL argument");
    TOSS_ENTER(ctx);
    return launch_fitness_positions(ctx, fn, p, d_pos, batch, N, d_ts, d_fitness, d_n_visited, nullptr,
                                    (cudaStream_t)stream);
}

int toss_update_tensor(toss_ctx *ctx, const toss_params *p, const double *d_pos, int64_t N, const double *d_ts,
                       uint8_t *h_bool_tensor_out, uintptr_t stream)
{
    TOSS_REQUIRE(ctx && p && d_pos && d_ts, "toss_update_tensor: NULL argument");
    TOSS_REQUIRE(ctx->d_grid != nullptr, "toss_update_tensor: no voxel grid uploaded");
    TOSS_ENTER(ctx);
    cudaStream_t st = (cudaStream_t)stream;
    const long long nvox = (long long)ctx->nr * ctx->nth * ctx->nph;
    const size_t nw = (size_t)((nvox + 31) / 32);
    // [fitness (8 B, unused) | visited bits] in a small buffer the context keeps (no allocation per call)
    if (ctx->tensor_scratch_bytes < 256 + nw * 4) {
        cudaFree(ctx->d_tensor_scratch);
        ctx->d_tensor_scratch = nullptr;
        ctx->tensor_scratch_bytes = 0;
        TOSS_CHECK_CUDA(cudaMalloc(&ctx->d_tensor_scratch, 256 + nw * 4));
        ctx->tensor_scratch_bytes = 256 + nw * 4;
    }
    double *d_fit = static_cast<double *>(ctx->d_tensor_scratch);
    uint32_t *d_bits = reinterpret_cast<uint32_t *>(static_cast<unsigned char *>(ctx->d_tensor_scratch) + 256);
    NvtxRange nv("toss:update_tensor");
    int rc = launch_fitness_positions(ctx, 8, p, d_pos, 1, N, d_ts, d_fit, nullptr, d_bits, st);
    std::vector<uint32_t> bits(nw);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(bits.data(), d_bits, nw * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            toss_set_error(std::string("toss_update_tensor: ") + cudaGetErrorString(e));
            rc = 1;
        }
    }
    if (rc) return rc;
    for (long long b = 0; b < nvox; ++b) ctx->h_bool[b] = (bits[b >> 5] >> (b & 31)) & 1u;
    if (h_bool_tensor_out) memcpy(h_bool_tensor_out, ctx->h_bool, (size_t)nvox);
    return upload_bool_bits(ctx);
}

int toss_is_outside(toss_ctx *ctx, const double *d_pts, int64_t n, uint8_t *d_out, uintptr_t stream)
{
    TOSS_REQUIRE(ctx && (n == 0 || (d_pts && d_out)), "toss_is_outside: NULL argument");
    TOSS_ENTER(ctx);
    return launch_is_outside(ctx, d_pts, n, d_out, (cudaStream_t)stream);
}

int toss_population_fitness(toss_ctx *ctx, const toss_params *p, const double *d_x, int pop, int dim,
                            const double *h_fixed, int n_fixed, int n_man, void *d_workspace, int knot_cap,
                            double *d_fitness, int32_t *d_collision, toss_counters *d_counters, int64_t *d_n_visited,
                            uintptr_t stream)
{
    TOSS_REQUIRE(ctx && p && d_workspace && d_fitness, "toss_population_fitness: NULL argument");
    TOSS_ENTER(ctx);
    cudaStream_t st = (cudaStream_t)stream;
    NvtxRange nv("toss:population_fitness");
    if (int rc = launch_propagate(ctx, p, d_x, pop, dim, h_fixed, n_fixed, n_man, d_workspace, knot_cap, st)) return rc;
    NvtxRange nv2("toss:K3+K4 resample+fitness");
    return launch_fitness_knots(ctx, p, d_workspace, pop, n_man, knot_cap, d_fitness, d_collision, d_counters,
                                d_n_visited, st);
}

int toss_population_fitness_host(toss_ctx *ctx, const toss_params *p, const double *h_x, int pop, int dim,
                                 const double *h_fixed, int n_fixed, int n_man, int knot_cap, double *h_fitness,
                                 int32_t *h_collision, toss_counters *h_counters, int64_t *h_n_visited)
{
    TOSS_REQUIRE(ctx && p && h_fitness && (h_x || dim == 0), "toss_population_fitness_host: NULL argument");
    TOSS_REQUIRE(pop > 0, "toss_population_fitness_host: empty population");
    TOSS_ENTER(ctx);
    toss_ws_layout L;
    toss_workspace_layout_impl(pop, n_man, knot_cap, &L);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_x = up(L.total_bytes);
    const size_t o_fit = up(o_x + sizeof(double) * (size_t)pop * (dim > 0 ? dim : 1));
    const size_t o_coll = up(o_fit + sizeof(double) * pop);
    const size_t o_cnt = up(o_coll + sizeof(int32_t) * pop);
    const size_t o_nv = up(o_cnt + sizeof(toss_counters) * pop);
    const size_t total = up(o_nv + sizeof(int64_t) * pop);
    if (ctx->scratch_bytes < total) {
        cudaFree(ctx->d_scratch);
        ctx->d_scratch = nullptr;
        ctx->scratch_bytes = 0;
        TOSS_CHECK_CUDA(cudaMalloc(&ctx->d_scratch, total));
        ctx->scratch_bytes = total;
    }
    // pinned staging: [x | fitness | collision | counters | n_visited]
    const size_t hp_x = 0, hp_fit = up(sizeof(double) * (size_t)pop * (dim > 0 ? dim : 1));
    const size_t hp_coll = up(hp_fit + sizeof(double) * pop), hp_cnt = up(hp_coll + sizeof(int32_t) * pop);
    const size_t hp_nv = up(hp_cnt + sizeof(toss_counters) * pop), hp_total = up(hp_nv + sizeof(int64_t) * pop);
    if (ctx->pinned_bytes < hp_total) {
        if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
        ctx->h_pinned = nullptr;
        ctx->pinned_bytes = 0;
        TOSS_CHECK_CUDA(cudaMallocHost(&ctx->h_pinned, hp_total));
        ctx->pinned_bytes = hp_total;
    }
    unsigned char *ws = static_cast<unsigned char *>(ctx->d_scratch);
    unsigned char *hp = static_cast<unsigned char *>(ctx->h_pinned);
    cudaStream_t st = 0;
    if (dim > 0) {
        memcpy(hp + hp_x, h_x, sizeof(double) * (size_t)pop * dim);
        TOSS_CHECK_CUDA(cudaMemcpyAsync(ws + o_x, hp + hp_x, sizeof(double) * (size_t)pop * dim, cudaMemcpyHostToDevice, st));
    }
    if (int rc = toss_population_fitness(ctx, p, reinterpret_cast<double *>(ws + o_x), pop, dim, h_fixed, n_fixed, n_man,
                                         ws, knot_cap, reinterpret_cast<double *>(ws + o_fit),
                                         reinterpret_cast<int32_t *>(ws + o_coll),
                                         reinterpret_cast<toss_counters *>(ws + o_cnt),
                                         reinterpret_cast<int64_t *>(ws + o_nv), (uintptr_t)st))
        return rc;
    TOSS_CHECK_CUDA(cudaMemcpyAsync(hp + hp_fit, ws + o_fit, sizeof(double) * pop, cudaMemcpyDeviceToHost, st));
    if (h_collision) TOSS_CHECK_CUDA(cudaMemcpyAsync(hp + hp_coll, ws + o_coll, sizeof(int32_t) * pop, cudaMemcpyDeviceToHost, st));
    if (h_counters) TOSS_CHECK_CUDA(cudaMemcpyAsync(hp + hp_cnt, ws + o_cnt, sizeof(toss_counters) * pop, cudaMemcpyDeviceToHost, st));
    if (h_n_visited) TOSS_CHECK_CUDA(cudaMemcpyAsync(hp + hp_nv, ws + o_nv, sizeof(int64_t) * pop, cudaMemcpyDeviceToHost, st));
    TOSS_CHECK_CUDA(cudaStreamSynchronize(st));
    memcpy(h_fitness, hp + hp_fit, sizeof(double) * pop);
    if (h_collision) memcpy(h_collision, hp + hp_coll, sizeof(int32_t) * pop);
    if (h_counters) memcpy(h_counters, hp + hp_cnt, sizeof(toss_counters) * pop);
    if (h_n_visited) memcpy(h_n_visited, hp + hp_nv, sizeof(int64_t) * pop);
    return 0;
}

// ---- multi-GPU plumbing (sharding.cu, stepper.cu) ------------------------------------------------------
int toss_set_queue_order(toss_ctx *ctx, int mode)
{
    TOSS_REQUIRE(ctx && (mode == 0 || mode == 1), "toss_set_queue_order: mode must be 0 (auto) or 1 (rows pre-ordered)");
    ctx->queue_order_mode = mode;
    return 0;
}

int toss_predict_cost(toss_ctx *ctx, const toss_params *p, const double *d_x, int pop, int dim, const double *h_fixed,
                      int n_fixed, int n_man, int32_t *d_cost, uintptr_t stream)
{
    TOSS_REQUIRE(ctx && p && d_cost && (d_x || dim == 0), "toss_predict_cost: NULL argument");
    TOSS_ENTER(ctx);
    return launch_predict_cost(ctx, p, d_x, pop, dim, h_fixed, n_fixed, n_man, d_cost, (cudaStream_t)stream);
}

int toss_cost_order(toss_ctx *ctx, const int32_t *d_cost, int n, int32_t *d_order, uintptr_t stream)
{
    TOSS_REQUIRE(ctx != nullptr, "toss_cost_order: NULL context");
    TOSS_ENTER(ctx);
    return launch_cost_order(ctx, d_cost, n, d_order, (cudaStream_t)stream);
}

int toss_deal_rows(toss_ctx *ctx, const double *d_x, int n, int dim, const int32_t *d_order, int rank, int world,
                   double *d_x_local, uintptr_t stream)
{
    TOSS_REQUIRE(ctx != nullptr, "toss_deal_rows: NULL context");
    TOSS_ENTER(ctx);
    return launch_deal_rows(ctx, d_x, n, dim, d_order, rank, world, d_x_local, (cudaStream_t)stream);
}

int toss_pack_result(toss_ctx *ctx, const double *d_fitness_local, const toss_counters *d_counters_local, int n_local,
                     int cap, double *d_send, uintptr_t stream)
{
    TOSS_REQUIRE(ctx != nullptr, "toss_pack_result: NULL context");
    TOSS_ENTER(ctx);
    return launch_pack_result(ctx, d_fitness_local, d_counters_local, n_local, cap, d_send, (cudaStream_t)stream);
}

int toss_unpack_result(toss_ctx *ctx, const double *d_recv, const int32_t *d_order, int n, int world, int cap,
                       double *d_fitness, double *d_flags, uintptr_t stream)
{
    TOSS_REQUIRE(ctx != nullptr, "toss_unpack_result: NULL context");
    TOSS_ENTER(ctx);
    return launch_unpack_result(ctx, d_recv, d_order, n, world, cap, d_fitness, d_flags, (cudaStream_t)stream);
}

int toss_gaco_sample(toss_ctx *ctx, const double *h_archive_x, int ker, int dim, const double *h_cum_weights,
                     const double *h_sigma, const double *h_lb, const double *h_ub, uint64_t seed, uint64_t generation,
                     int first_ant, int n, double *d_x_out, uintptr_t stream)
{
    TOSS_REQUIRE(ctx != nullptr, "toss_gaco_sample: NULL context");
    TOSS_ENTER(ctx);
    return launch_gaco_sample(ctx, h_archive_x, ker, dim, h_cum_weights, h_sigma, h_lb, h_ub, seed, generation, first_ant,
                              n, d_x_out, (cudaStream_t)stream);
}

int toss_math_probe(toss_ctx *ctx, int fn, const double *d_a, const double *d_b, int64_t n, double *d_out,
                    double *d_out2, uintptr_t stream)
{
    TOSS_REQUIRE(ctx && (n == 0 || (d_a && d_out)), "toss_math_probe: NULL argument");
    TOSS_REQUIRE(fn >= 0 && fn <= 14, "toss_math_probe: unknown function id");
    TOSS_REQUIRE(!(fn == 2 || fn == 7) || d_b, "toss_math_probe: this function needs a second operand");
    TOSS_ENTER(ctx);
    return launch_math_probe(ctx, fn, d_a, d_b, n, d_out, d_out2, (cudaStream_t)stream);
}

int toss_selftest_div_sqrt(toss_ctx *ctx, int64_t samples, int64_t *bad_div, int64_t *bad_sqrt)
{
    TOSS_REQUIRE(ctx && bad_div && bad_sqrt && samples > 0, "toss_selftest_div_sqrt: bad argument");
    TOSS_ENTER(ctx);
    return launch_div_sqrt_selftest(ctx, samples, bad_div, bad_sqrt);
}

int toss_measure_fp64_peak(toss_ctx *ctx, double *tflops_out)
{
    TOSS_REQUIRE(ctx && tflops_out, "toss_measure_fp64_peak: NULL argument");
    TOSS_ENTER(ctx);
    return launch_fp64_peak(ctx, tflops_out);
}

}   // extern "C"
