// toss_b200/csrc/pairs.cu -- host side of the face-PAIR formulation of the polyhedral sum.
//
// Two triangles that share an edge share two of their three vertices and one of their three edge terms:
// evaluating them together costs 4 vertex distances and 5 edge logarithms instead of 6 and 6 (the mesh is
// closed, so its dual graph is 3-regular and almost every face finds a partner).  This file (1) matches the
// faces into pairs -- integer work, deterministic, so a second implementation reproduces the same list -- and
// (2) builds the per-pair constants of SURVEY.md App. A.1 (what GravityEvaluable precomputes at construction,
// setup_parameters.py:63).
#include <math.h>

#include <deque>
#include <vector>

#include "common.cuh"

// Karp-Sipser matching on the dual graph.  nb[f][q] = the face across edge q (v_q -> v_{q+1}) of f when exactly
// two faces share that edge with opposite directions, else -1.  Faces with one free neighbour left are served
// first (FIFO), otherwise the lowest-numbered free face takes its free neighbour of smallest (degree, edge slot).
// Output, ascending in the first face: (fA, qA, fB) with fA < fB, qA = A's edge slot of the shared edge;
// fB = -1 for a face left without a partner.
void toss_build_face_pairs(const int32_t *faces, int F, int V, std::vector<int32_t> &out, int *n_unmatched, int *first_unmatched)
{
    // edge table: for every directed half-edge (u -> v) of face f slot q, key = min*V + max; bucket by counting sort
    std::vector<int64_t> key((size_t)3 * F);
    std::vector<int32_t> ord((size_t)3 * F);
    for (int h = 0; h < 3 * F; ++h) {
        const int f = h / 3, q = h % 3;
        const int64_t u = faces[3 * f + q], v = faces[3 * f + (q + 1) % 3];
        key[h] = (u < v ? u : v) * (int64_t)V + (u < v ? v : u);
        ord[h] = h;
    }
    // stable LSD radix sort of the half-edge ids by key (two passes of counting sort on max / min vertex)
    {
        std::vector<int32_t> tmp((size_t)3 * F), cnt((size_t)V + 1);
        for (int pass = 0; pass < 2; ++pass) {
            std::fill(cnt.begin(), cnt.end(), 0);
            for (int i = 0; i < 3 * F; ++i) {
                const int64_t k = key[ord[i]];
                const int d = (int)(pass == 0 ? k % V : k / V);
                cnt[d + 1]++;
            }
            for (int d = 0; d < V; ++d) cnt[d + 1] += cnt[d];
            for (int i = 0; i < 3 * F; ++i) {
                const int64_t k = key[ord[i]];
                const int d = (int)(pass == 0 ? k % V : k / V);
                tmp[cnt[d]++] = ord[i];
            }
            ord.swap(tmp);
        }
    }
    std::vector<int32_t> nb((size_t)3 * F, -1);
    for (int i = 0; i < 3 * F;) {
        int j = i;
        while (j < 3 * F && key[ord[j]] == key[ord[i]]) ++j;
        if (j - i == 2) {
            const int h0 = ord[i], h1 = ord[i + 1];
            const int f0 = h0 / 3, q0 = h0 % 3, f1 = h1 / 3, q1 = h1 % 3;
            const bool opposite = faces[3 * f0 + q0] == faces[3 * f1 + (q1 + 1) % 3] &&
                                  faces[3 * f0 + (q0 + 1) % 3] == faces[3 * f1 + q1];
            if (opposite && f0 != f1) {
                nb[h0] = f1;
                nb[h1] = f0;
            }
        }
        i = j;
    }
    if (n_unmatched) {   // a closed, consistently oriented surface has none (toss_mesh_upload refuses anything else)
        *n_unmatched = 0;
        for (int h = 0; h < 3 * F; ++h)
            if (nb[h] < 0 && (*n_unmatched)++ == 0 && first_unmatched) *first_unmatched = h;
    }
    std::vector<int32_t> mate((size_t)F, -1), deg((size_t)F, 0);
    for (int f = 0; f < F; ++f)
        for (int q = 0; q < 3; ++q) deg[f] += nb[3 * f + q] >= 0;
    std::deque<int32_t> ones;
    for (int f = 0; f < F; ++f)
        if (deg[f] == 1) ones.push_back(f);
    int cur = 0;
    for (;;) {
        int f;
        if (!ones.empty()) {
            f = ones.front();
            ones.pop_front();
            if (mate[f] >= 0 || deg[f] == 0) continue;
        } else {
            while (cur < F && (mate[cur] >= 0 || deg[cur] == 0)) ++cur;
            if (cur >= F) break;
            f = cur;
        }
        int g = -1, gdeg = 0;
        for (int q = 0; q < 3; ++q) {
            const int h = nb[3 * f + q];
            if (h >= 0 && mate[h] < 0 && (g < 0 || deg[h] < gdeg)) {
                g = h;
                gdeg = deg[h];
            }
        }
        mate[f] = g;
        mate[g] = f;
        for (int s = 0; s < 2; ++s) {
            const int x = s == 0 ? f : g;
            for (int q = 0; q < 3; ++q) {
                const int h = nb[3 * x + q];
                if (h >= 0 && mate[h] < 0) {
                    if (--deg[h] == 1) ones.push_back(h);
                }
            }
        }
    }
    out.clear();
    for (int f = 0; f < F; ++f) {
        if (mate[f] < 0) {
            out.push_back(f);
            out.push_back(0);
            out.push_back(-1);
        } else if (f < mate[f]) {
            int qa = 0;
            while (nb[3 * f + qa] != mate[f]) ++qa;
            out.push_back(f);
            out.push_back(qa);
            out.push_back(mate[f]);
        }
    }
}

namespace {
struct FaceConsts {
    double N[3], n[3][3], len[3], twoA;
};
// N = G0 x G1 / |.|, n_q = G_q x N / |.|, |G_q|, 2A = |G0 x G1|, G_q = w_{q+1} - w_q.  Plain IEEE + - * / sqrt in a
// fixed order (the file is compiled -fmad=false): a second implementation of the same sequence gives the same bits.
__device__ bool face_consts(const double *w0, const double *w1, const double *w2, FaceConsts &o)
{
    const double *w[3] = {w0, w1, w2};
    double G[3][3];
    for (int q = 0; q < 3; ++q)
        for (int c = 0; c < 3; ++c) G[q][c] = w[(q + 1) % 3][c] - w[q][c];
    double N[3] = {G[0][1] * G[1][2] - G[0][2] * G[1][1], G[0][2] * G[1][0] - G[0][0] * G[1][2],
                   G[0][0] * G[1][1] - G[0][1] * G[1][0]};
    const double nn = sqrt(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]);
    if (!(nn > 0.0)) return false;
    for (int c = 0; c < 3; ++c) o.N[c] = N[c] / nn;
    for (int q = 0; q < 3; ++q) {
        o.len[q] = sqrt(G[q][0] * G[q][0] + G[q][1] * G[q][1] + G[q][2] * G[q][2]);
        const double nq[3] = {G[q][1] * o.N[2] - G[q][2] * o.N[1], G[q][2] * o.N[0] - G[q][0] * o.N[2],
                              G[q][0] * o.N[1] - G[q][1] * o.N[0]};
        const double nl = sqrt(nq[0] * nq[0] + nq[1] * nq[1] + nq[2] * nq[2]);
        for (int c = 0; c < 3; ++c) o.n[q][c] = nq[c] / nl;
    }
    o.twoA = nn;
    return true;
}

// One pair record (common.cuh layout).  A = (v0, v1, v2) is face fA rotated so that the shared edge comes
// first; B = (v1, v0, v3) is face fB in its own orientation.  A face without a partner gets a null B
// (v3 = v2, all of B's constants zero: S_B = 0 exactly).
__device__ bool pair_record(const double *verts, const int32_t *faces, int fA, int qA, int fB, double *r)
{
    const int i0 = faces[3 * fA + qA], i1 = faces[3 * fA + (qA + 1) % 3], i2 = faces[3 * fA + (qA + 2) % 3];
    int i3 = i2;
    if (fB >= 0) {
        int qb = 0;
        while (qb < 2 && !(faces[3 * fB + qb] == i1 && faces[3 * fB + (qb + 1) % 3] == i0)) ++qb;
        i3 = faces[3 * fB + (qb + 2) % 3];
    }
    const double *v0 = verts + 3 * (size_t)i0, *v1 = verts + 3 * (size_t)i1, *v2 = verts + 3 * (size_t)i2,
                 *v3 = verts + 3 * (size_t)i3;
    for (int c = 0; c < 3; ++c) {
        r[c] = v0[c];
        r[3 + c] = v1[c];
        r[6 + c] = v2[c];
        r[9 + c] = v3[c];
    }
    FaceConsts A, B;
    if (!face_consts(v0, v1, v2, A)) return false;
    for (int c = 0; c < 3; ++c) r[12 + c] = A.N[c];
    for (int q = 0; q < 3; ++q)
        for (int c = 0; c < 3; ++c) r[18 + 3 * q + c] = A.n[q][c];
    r[36] = A.len[0];
    r[37] = A.len[1];
    r[38] = A.len[2];
    r[41] = A.twoA;
    if (fB >= 0) {
        if (!face_consts(v1, v0, v3, B)) return false;
        for (int c = 0; c < 3; ++c) r[15 + c] = B.N[c];
        for (int q = 0; q < 3; ++q)
            for (int c = 0; c < 3; ++c) r[27 + 3 * q + c] = B.n[q][c];
        r[39] = B.len[1];
        r[40] = B.len[2];
        r[42] = B.twoA;
    } else {
        for (int k = 15; k < 18; ++k) r[k] = 0.0;
        for (int k = 27; k < 36; ++k) r[k] = 0.0;
        r[39] = r[40] = r[42] = 0.0;
    }
    return true;
}

// Thread i builds the record of pair i (a padding record beyond n_pairs: every vertex at kPadVertex, every constant 0)
// and writes it into the three encodings of DESIGN.md section 4:
//   aos   [n_aos_slots][44]                  K1a, LDS.128 broadcast reads
//   soa   [22][Ppad32] double2               K1b, resident stepper
//   tiles [n_tiles][22][kPairTile] double2   streaming stepper (pair i -> tile i / kPairTile, position i % kPairTile)
// A zero-area face raises *degenerate (the upload then fails).
__global__ void __launch_bounds__(128) pair_tables_kernel(const double *__restrict__ verts, const int32_t *__restrict__ faces,
                                                          const int32_t *__restrict__ pairs, int n_pairs, int n_aos_slots,
                                                          int Ppad32, int n_tile_slots, double *__restrict__ aos,
                                                          double *__restrict__ soa, double *__restrict__ tiles,
                                                          int *__restrict__ degenerate)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_aos_slots && i >= Ppad32 && i >= n_tile_slots) return;
    double r[kPairFields];
    bool ok = true;
    if (i < n_pairs) ok = pair_record(verts, faces, pairs[3 * i], pairs[3 * i + 1], pairs[3 * i + 2], r);
    if (i >= n_pairs || !ok) {
        for (int k = 0; k < kPairFields; ++k) r[k] = k < 12 ? kPadVertex : 0.0;
        if (!ok) atomicExch(degenerate, 1 + pairs[3 * i]);
    }
    if (i < n_aos_slots) {
        for (int k = 0; k < kPairFields; ++k) aos[(size_t)i * kPairAosStride + k] = r[k];
        aos[(size_t)i * kPairAosStride + kPairFields] = 0.0;
    }
    if (i < Ppad32) {
        for (int k = 0; k < 2 * kPairSlots; ++k)
            soa[((size_t)(k >> 1) * Ppad32 + i) * 2 + (k & 1)] = k < kPairFields ? r[k] : 0.0;
    }
    if (i < n_tile_slots) {
        const int t = i / kPairTile, j = i % kPairTile;
        for (int k = 0; k < 2 * kPairSlots; ++k)
            tiles[(((size_t)t * kPairSlots + (k >> 1)) * kPairTile + j) * 2 + (k & 1)] = k < kPairFields ? r[k] : 0.0;
    }
}

// the faces themselves, original vertex order, [9][Fpad32] -> ray casting (K5, event points)
__global__ void __launch_bounds__(128) ray_table_kernel(const double *__restrict__ verts, const int32_t *__restrict__ faces,
                                                        int F, int Fpad32, double *__restrict__ ray)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= Fpad32) return;
    for (int k = 0; k < kRayFields; ++k)
        ray[(size_t)k * Fpad32 + f] = f < F ? verts[3 * (size_t)faces[3 * f + k / 3] + k % 3] : kPadVertex;
}
}   // namespace

// Builds every device table of the mesh from (verts, faces, pairs) ON the device: the host only matches the faces
// (integer work) and copies the three index / coordinate arrays up.  Returns 0, or 1 + face index of a zero-area face.
int launch_mesh_tables(toss_ctx *ctx, const double *h_verts, int V, const int32_t *h_faces, int F, const int32_t *h_pairs,
                       int n_pairs, int *bad_face)
{
    *bad_face = -1;
    const int n_aos_slots = ctx->n_pair_aos_tiles * kPairAosTile, n_tile_slots = ctx->n_pair_tiles * kPairTile;
    double *d_verts = nullptr;
    int32_t *d_faces = nullptr, *d_pairs = nullptr;
    int *d_flag = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_verts);
        cudaFree(d_faces);
        cudaFree(d_pairs);
        cudaFree(d_flag);
    };
#define MT_CHECK(call)                                        \
    do {                                                      \
        cudaError_t e_ = (call);                              \
        if (e_ != cudaSuccess) {                              \
            toss_set_error(std::string(#call) + ": " + cudaGetErrorString(e_)); \
            cleanup();                                        \
            return -1;                                        \
        }                                                     \
    } while (0)
    MT_CHECK(cudaMalloc(&d_verts, sizeof(double) * 3 * (size_t)V));
    MT_CHECK(cudaMalloc(&d_faces, sizeof(int32_t) * 3 * (size_t)F));
    MT_CHECK(cudaMalloc(&d_pairs, sizeof(int32_t) * 3 * (size_t)n_pairs));
    MT_CHECK(cudaMalloc(&d_flag, sizeof(int)));
    MT_CHECK(cudaMemcpy(d_verts, h_verts, sizeof(double) * 3 * (size_t)V, cudaMemcpyHostToDevice));
    MT_CHECK(cudaMemcpy(d_faces, h_faces, sizeof(int32_t) * 3 * (size_t)F, cudaMemcpyHostToDevice));
    MT_CHECK(cudaMemcpy(d_pairs, h_pairs, sizeof(int32_t) * 3 * (size_t)n_pairs, cudaMemcpyHostToDevice));
    MT_CHECK(cudaMemset(d_flag, 0, sizeof(int)));
    MT_CHECK(cudaMalloc(&ctx->d_pair_aos, sizeof(double) * (size_t)n_aos_slots * kPairAosStride));
    MT_CHECK(cudaMalloc(&ctx->d_pair_soa, sizeof(double) * 2 * kPairSlots * (size_t)ctx->Ppad32));
    MT_CHECK(cudaMalloc(&ctx->d_pair_tiles, sizeof(double) * (size_t)ctx->n_pair_tiles * kPairTileDoubles));
    MT_CHECK(cudaMalloc(&ctx->d_ray_soa, sizeof(double) * kRayFields * (size_t)ctx->Fpad32));
    int n = n_aos_slots > ctx->Ppad32 ? n_aos_slots : ctx->Ppad32;
    n = n > n_tile_slots ? n : n_tile_slots;
    pair_tables_kernel<<<(n + 127) / 128, 128>>>(d_verts, d_faces, d_pairs, n_pairs, n_aos_slots, ctx->Ppad32, n_tile_slots,
                                                 ctx->d_pair_aos, ctx->d_pair_soa, ctx->d_pair_tiles, d_flag);
    ray_table_kernel<<<(ctx->Fpad32 + 127) / 128, 128>>>(d_verts, d_faces, F, ctx->Fpad32, ctx->d_ray_soa);
    ctx->launches += 2;
    MT_CHECK(cudaGetLastError());
    int flag = 0;
    MT_CHECK(cudaMemcpy(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost));
#undef MT_CHECK
    cleanup();
    if (flag) *bad_face = flag - 1;
    return flag ? 1 : 0;
}
