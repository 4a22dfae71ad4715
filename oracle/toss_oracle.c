/* oracle/toss_oracle.c -- CPU restatement of the TOSS population-fitness path.
 * TEST INFRASTRUCTURE ONLY -- see toss_oracle.h for scope, citations and the parity pin.
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -pthread)
 */
#include "toss_oracle.h"
#include "rk8713m_tableau.h"
#include "orc_math.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>

#define ORC_G 6.67430e-11 /* CODATA-2018, the value polyhedral_gravity >= 2.x uses */
#define REC 34            /* doubles per face record */
#define PREC 43           /* doubles per pair record */

struct orc_mesh {
    int V, F;
    double *verts;
    int32_t *faces;
    double *rec; /* per face: v0 v1 v2 | N | g0 g1 g2 | n0 n1 n2 | len0 len1 len2 | 2*area */
    int n_pairs;
    int32_t *pairs; /* [n_pairs][3]: fA, qA (A's slot of the shared edge), fB (-1: no partner) */
    double *prec;   /* per pair: v0 v1 v2 v3 | N_A N_B | nA0 nA1 nA2 | nB0 nB1 nB2 | G0..G4 | 2A_A 2A_B */
    double Grho;
};

static inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(const double *a, const double *b, double *c)
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
static inline double norm3(const double *a) { return sqrt(dot3(a, a)); }

/* minimal pthread parallel-for with dynamic (atomic counter) scheduling; libgomp is not
 * usable in this image.  threads <= 1 runs inline. */
typedef void (*pf_body)(int64_t i, void *arg);
typedef struct {
    pf_body body;
    void *arg;
    int64_t n, chunk;
    atomic_llong next;
} pf_job;
static void *pf_worker(void *v)
{
    pf_job *j = (pf_job *)v;
    for (;;) {
        int64_t b = atomic_fetch_add(&j->next, j->chunk);
        if (b >= j->n) break;
        int64_t e = b + j->chunk < j->n ? b + j->chunk : j->n;
        for (int64_t i = b; i < e; ++i) j->body(i, j->arg);
    }
    return NULL;
}
static void parallel_for(int64_t n, int64_t chunk, int threads, pf_body body, void *arg)
{
    if (threads <= 1 || n <= 1) {
        for (int64_t i = 0; i < n; ++i) body(i, arg);
        return;
    }
    if (threads > 256) threads = 256;
    pf_job job = {body, arg, n, chunk > 0 ? chunk : 1, 0};
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < threads - 1; ++t)
        if (pthread_create(&th[started], NULL, pf_worker, &job) == 0) ++started;
    pf_worker(&job);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
}

/* ---------------------------------------------------------------------------------------
 * Face pairs.  The polyhedral sum is evaluated two faces at a time: faces that share an edge share two
 * vertices (their distances to the field point) and one edge term LN.  Which faces go together is fixed by a
 * deterministic matching on the dual graph (Karp-Sipser: faces with a single free neighbour first, FIFO;
 * otherwise the lowest-numbered free face takes its free neighbour of smallest (degree, edge slot)); records
 * ascend in the lower face index.  Integer work: any implementation of this rule produces the same list. */
static int cmp_i64_pair(const void *a, const void *b)
{
    const int64_t *x = (const int64_t *)a, *y = (const int64_t *)b;
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    return x[1] < y[1] ? -1 : (x[1] > y[1] ? 1 : 0);
}
static int build_pairs(const int32_t *faces, int F, int V, int32_t *out /* 3*F */)
{
    int64_t *he = (int64_t *)malloc(sizeof(int64_t) * 2 * 3 * (size_t)F); /* (key, half-edge id) */
    for (int h = 0; h < 3 * F; ++h) {
        const int f = h / 3, q = h % 3;
        const int64_t u = faces[3 * f + q], v = faces[3 * f + (q + 1) % 3];
        he[2 * h] = (u < v ? u : v) * (int64_t)V + (u < v ? v : u);
        he[2 * h + 1] = h;
    }
    qsort(he, 3 * (size_t)F, 2 * sizeof(int64_t), cmp_i64_pair);
    int32_t *nb = (int32_t *)malloc(sizeof(int32_t) * 3 * (size_t)F);
    for (int h = 0; h < 3 * F; ++h) nb[h] = -1;
    for (int i = 0; i < 3 * F;) {
        int j = i;
        while (j < 3 * F && he[2 * j] == he[2 * i]) ++j;
        if (j - i == 2) {
            const int h0 = (int)he[2 * i + 1], h1 = (int)he[2 * i + 3];
            const int f0 = h0 / 3, q0 = h0 % 3, f1 = h1 / 3, q1 = h1 % 3;
            if (f0 != f1 && faces[3 * f0 + q0] == faces[3 * f1 + (q1 + 1) % 3] &&
                faces[3 * f0 + (q0 + 1) % 3] == faces[3 * f1 + q1]) {
                nb[h0] = f1;
                nb[h1] = f0;
            }
        }
        i = j;
    }
    int32_t *mate = (int32_t *)malloc(sizeof(int32_t) * (size_t)F), *deg = (int32_t *)calloc((size_t)F, sizeof(int32_t));
    int32_t *fifo = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)F);
    int head = 0, tail = 0, cur = 0;
    for (int f = 0; f < F; ++f) {
        mate[f] = -1;
        for (int q = 0; q < 3; ++q) deg[f] += nb[3 * f + q] >= 0;
    }
    for (int f = 0; f < F; ++f)
        if (deg[f] == 1) fifo[tail++] = f;
    for (;;) {
        int f;
        if (head < tail) {
            f = fifo[head++];
            if (mate[f] >= 0 || deg[f] == 0) continue;
        } else {
            while (cur < F && (mate[cur] >= 0 || deg[cur] == 0)) ++cur;
            if (cur >= F) break;
            f = cur;
        }
        int g = -1;
        for (int q = 0; q < 3; ++q) {
            const int h = nb[3 * f + q];
            if (h >= 0 && mate[h] < 0 && (g < 0 || deg[h] < deg[g])) g = h;
        }
        mate[f] = g;
        mate[g] = f;
        const int both[2] = {f, g};
        for (int s = 0; s < 2; ++s)
            for (int q = 0; q < 3; ++q) {
                const int h = nb[3 * both[s] + q];
                if (h >= 0 && mate[h] < 0 && --deg[h] == 1) fifo[tail++] = h;
            }
    }
    int n = 0;
    for (int f = 0; f < F; ++f) {
        if (mate[f] < 0) {
            out[3 * n] = f;
            out[3 * n + 1] = 0;
            out[3 * n + 2] = -1;
            ++n;
        } else if (f < mate[f]) {
            int qa = 0;
            while (nb[3 * f + qa] != mate[f]) ++qa;
            out[3 * n] = f;
            out[3 * n + 1] = qa;
            out[3 * n + 2] = mate[f];
            ++n;
        }
    }
    free(he);
    free(nb);
    free(mate);
    free(deg);
    free(fifo);
    return n;
}

/* SURVEY.md App. A.1 constants of one face (w0, w1, w2): N = G0 x G1 / |.|, n_q = G_q x N / |.|, |G_q|, 2A */
static void face_constants(const double *w0, const double *w1, const double *w2, double N[3], double n[9],
                           double len[3], double *twoA)
{
    const double *w[3] = {w0, w1, w2};
    double Gq[3][3];
    for (int q = 0; q < 3; ++q)
        for (int c = 0; c < 3; ++c) Gq[q][c] = w[(q + 1) % 3][c] - w[q][c];
    cross3(Gq[0], Gq[1], N);
    const double nn = norm3(N);
    for (int c = 0; c < 3; ++c) N[c] /= nn;
    for (int q = 0; q < 3; ++q) {
        double nq[3];
        len[q] = norm3(Gq[q]);
        cross3(Gq[q], N, nq);
        const double nl = norm3(nq);
        for (int c = 0; c < 3; ++c) n[3 * q + c] = nq[c] / nl;
    }
    *twoA = nn;
}

/* A = (v0, v1, v2): face fA rotated so that the shared edge comes first; B = (v1, v0, v3): face fB in its own
 * orientation.  A face without a partner gets a null B (v3 = v2, constants 0). */
static void pair_record(const double *verts, const int32_t *faces, int fA, int qA, int fB, double *r)
{
    const int i0 = faces[3 * fA + qA], i1 = faces[3 * fA + (qA + 1) % 3], i2 = faces[3 * fA + (qA + 2) % 3];
    int i3 = i2;
    if (fB >= 0) {
        int qb = 0;
        while (!(faces[3 * fB + qb] == i1 && faces[3 * fB + (qb + 1) % 3] == i0)) ++qb;
        i3 = faces[3 * fB + (qb + 2) % 3];
    }
    const double *v0 = verts + 3 * (size_t)i0, *v1 = verts + 3 * (size_t)i1, *v2 = verts + 3 * (size_t)i2,
                 *v3 = verts + 3 * (size_t)i3;
    memcpy(r, v0, 24);
    memcpy(r + 3, v1, 24);
    memcpy(r + 6, v2, 24);
    memcpy(r + 9, v3, 24);
    double len[3], twoA;
    face_constants(v0, v1, v2, r + 12, r + 18, len, &twoA);
    r[36] = len[0];
    r[37] = len[1];
    r[38] = len[2];
    r[41] = twoA;
    if (fB >= 0) {
        face_constants(v1, v0, v3, r + 15, r + 27, len, &twoA);
        r[39] = len[1];
        r[40] = len[2];
        r[42] = twoA;
    } else {
        for (int k = 15; k < 18; ++k) r[k] = 0.0;
        for (int k = 27; k < 36; ++k) r[k] = 0.0;
        r[39] = r[40] = r[42] = 0.0;
    }
}

/* setup_parameters.py:63  GravityEvaluable((vertices, faces), density): per-face constants
 * of SURVEY.md App. A.1 (segment unit vectors g_q, plane normal N, segment normals n_q). */
orc_mesh *orc_mesh_create(const double *verts, int V, const int32_t *faces, int F, double density)
{
    orc_mesh *m = (orc_mesh *)calloc(1, sizeof(orc_mesh));
    m->V = V;
    m->F = F;
    m->verts = (double *)malloc(sizeof(double) * 3 * (size_t)V);
    m->faces = (int32_t *)malloc(sizeof(int32_t) * 3 * (size_t)F);
    m->rec = (double *)malloc(sizeof(double) * REC * (size_t)F);
    memcpy(m->verts, verts, sizeof(double) * 3 * (size_t)V);
    memcpy(m->faces, faces, sizeof(int32_t) * 3 * (size_t)F);
    m->Grho = ORC_G * density;
    for (int f = 0; f < F; ++f) {
        double *r = m->rec + (size_t)f * REC;
        const double *v[3];
        for (int i = 0; i < 3; ++i) {
            v[i] = verts + 3 * (size_t)faces[3 * f + i];
            memcpy(r + 3 * i, v[i], 3 * sizeof(double));
        }
        double Gq[3][3];
        for (int q = 0; q < 3; ++q)
            for (int c = 0; c < 3; ++c) Gq[q][c] = v[(q + 1) % 3][c] - v[q][c];
        double N[3];
        cross3(Gq[0], Gq[1], N);
        double nn = norm3(N);
        for (int c = 0; c < 3; ++c) N[c] /= nn;
        memcpy(r + 9, N, 3 * sizeof(double));
        r[33] = nn; /* |G0 x G1| = 2 * area: r0.(r1 x r2) = (N.r0) * 2A */
        for (int q = 0; q < 3; ++q) {
            double len = norm3(Gq[q]);
            double nq[3];
            cross3(Gq[q], N, nq);
            double nl = norm3(nq);
            for (int c = 0; c < 3; ++c) {
                r[12 + 3 * q + c] = Gq[q][c] / len;
                r[21 + 3 * q + c] = nq[c] / nl;
            }
            r[30 + q] = len;
        }
    }
    m->pairs = (int32_t *)malloc(sizeof(int32_t) * 3 * (size_t)F);
    m->n_pairs = build_pairs(faces, F, V, m->pairs);
    m->prec = (double *)malloc(sizeof(double) * PREC * (size_t)m->n_pairs);
    for (int i = 0; i < m->n_pairs; ++i)
        pair_record(verts, faces, m->pairs[3 * i], m->pairs[3 * i + 1], m->pairs[3 * i + 2], m->prec + (size_t)i * PREC);
    return m;
}

void orc_mesh_free(orc_mesh *m)
{
    if (!m) return;
    free(m->verts);
    free(m->faces);
    free(m->rec);
    free(m->pairs);
    free(m->prec);
    free(m);
}

int orc_mesh_faces(const orc_mesh *m) { return m->F; }
int orc_mesh_pairs(const orc_mesh *m, int32_t *out)
{
    if (out) memcpy(out, m->pairs, sizeof(int32_t) * 3 * (size_t)m->n_pairs);
    return m->n_pairs;
}

/* polyhedral_gravity (Tsoulis 2012) -- SURVEY.md App. A.1, transcribed term by term.  Package sign
 * convention.  Kept as the cross-check of gravity_point() below (orc_gravity_direct). */
static void gravity_point_direct(const orc_mesh *m, const double P[3], double acc[3], double *pot)
{
    double ax = 0.0, ay = 0.0, az = 0.0, vv = 0.0;
    for (int f = 0; f < m->F; ++f) {
        const double *rc = m->rec + (size_t)f * REC;
        const double *N = rc + 9;
        double r[3][3], l[3];
        for (int i = 0; i < 3; ++i) {
            for (int c = 0; c < 3; ++c) r[i][c] = rc[3 * i + c] - P[c];
            l[i] = norm3(r[i]);
        }
        double hp_s = dot3(N, r[0]);
        double hp = fabs(hp_s);
        double S = 0.0;
        int npos = 0, nzero = 0;
        for (int q = 0; q < 3; ++q) {
            const double *a = r[q], *b = r[(q + 1) % 3];
            const double *g = rc + 12 + 3 * q, *n = rc + 21 + 3 * q;
            double hq_s = dot3(n, a);
            double hq = fabs(hq_s);
            double sg = (hq_s > 0.0) ? 1.0 : ((hq_s < 0.0) ? -1.0 : 0.0);
            double s1 = dot3(g, a), s2 = dot3(g, b);
            double l1 = l[q], l2 = l[(q + 1) % 3];
            double d1 = s1 + l1, d2 = s2 + l2;
            double LN = (d1 > 0.0 && d2 > 0.0) ? log(d2 / d1) : 0.0;
            double AN = 0.0;
            if (hp != 0.0 && hq != 0.0) AN = atan(hp * s2 / (hq * l2)) - atan(hp * s1 / (hq * l1));
            S += hq_s * LN + hp * sg * AN;
            if (sg > 0.0) ++npos;
            if (sg == 0.0) ++nzero;
        }
        if (npos == 3)
            S -= 2.0 * M_PI * hp; /* projection strictly inside the face */
        else if (npos == 2 && nzero == 1)
            S -= M_PI * hp; /* on an edge (measure zero) */
        ax += N[0] * S;
        ay += N[1] * S;
        az += N[2] * S;
        vv += hp_s * S;
    }
    acc[0] = m->Grho * ax;
    acc[1] = m->Grho * ay;
    acc[2] = m->Grho * az;
    if (pot) *pot = 0.5 * m->Grho * vv;
}


/* Same Tsoulis sum S_p = sum_q hq_s LN_pq + hp (sum_q sigma_pq AN_pq + sing A_p), with the two
 * transcendental arguments written in their cancellation-free closed forms:
 *   LN_pq = ln((s2+l2)/(s1+l1)) = ln((l1+l2+|G|)/(l1+l2-|G|)) = 2 atanh(|G| / (l1+l2))
 *   hp (sum_q sigma_pq AN_pq + sing A_p) = -hp_s * omega_p,
 *   omega_p = 2 atan2(r0.(r1 x r2), l0 l1 l2 + l0 (r1.r2) + l1 (r2.r0) + l2 (r0.r1))   (signed solid angle)
 * (Werner & Scheeres 1997 eqs. 7-10 / van Oosterom & Strackee 1983; identical values, see
 * tests/test_oracle_gravity.py which checks both forms against a 50-digit mpmath evaluation.)
 * The direct form loses ~ (l1+l2)/|G| ulps in ln(ratio ~ 1) and in the difference of two O(1)
 * arctangents; this form keeps every term to a few ulp, so it is the one parity is anchored on. */
/* One PAIR of faces that share an edge.  Operation order = the arithmetic specification the CUDA kernels
 * implement too (toss_b200/csrc/gravity.cuh pair_S): un-fused except where ORC_FMA says so.  Shared by the two
 * faces: |v0 - P|, |v1 - P|, r0.r1, l0*l1 and the edge term LN_0.
 * The seven quotients share ONE division and atanh / atan are short odd minimax polynomials; a term whose
 * argument is out of the polynomial's range (an edge or a face close to the field point) falls back, on its own,
 * to a division of its own + ln-based atanh, or to the full-range atan2. */
#define ORC_HI_Z2MAX 0x3FC47AE1 /* high word of 0.16: z^2 < 0x3FC47AE1'00000000 = 0.16 - 3.3e-8 */
#define ORC_HI_Q2MAX 0x3F847AE1
#define ORC_HI_FAR 0x3F647AE1 /* high word of 0.0025: z^2, q^2 < 0x3F647AE1'00000000 = 0.0025 - 5.2e-10 */
static inline int32_t orc_hi(double x)   /* high word of the IEEE-754 pattern, as a signed integer */
{
    uint64_t b;
    memcpy(&b, &x, 8);
    return (int32_t)(b >> 32);
}

/* Everything of a pair term up to the arguments of the transcendental functions. */
typedef struct {
    const double *rc;
    double hpA, hpB, hA0, hA1, hA2, hB0, hB1, hB2;
    double a[5], z[5], zs[5], q[2], qs[2], num[2], den[2];
} pair_pre;

static inline void pair_prelim(const double *rc, const double P[3], pair_pre *o)
{
    const double r0x = rc[0] - P[0], r0y = rc[1] - P[1], r0z = rc[2] - P[2];
    const double r1x = rc[3] - P[0], r1y = rc[4] - P[1], r1z = rc[5] - P[2];
    const double r2x = rc[6] - P[0], r2y = rc[7] - P[1], r2z = rc[8] - P[2];
    const double r3x = rc[9] - P[0], r3y = rc[10] - P[1], r3z = rc[11] - P[2];
    const double l0 = sqrt(ORC_FMA(r0z, r0z, ORC_FMA(r0y, r0y, r0x * r0x)));
    const double l1 = sqrt(ORC_FMA(r1z, r1z, ORC_FMA(r1y, r1y, r1x * r1x)));
    const double l2 = sqrt(ORC_FMA(r2z, r2z, ORC_FMA(r2y, r2y, r2x * r2x)));
    const double l3 = sqrt(ORC_FMA(r3z, r3z, ORC_FMA(r3y, r3y, r3x * r3x)));
    const double *NA = rc + 12, *NB = rc + 15, *nA = rc + 18, *nB = rc + 27, *G = rc + 36;
    o->rc = rc;
    o->hpA = ORC_FMA(NA[2], r0z, ORC_FMA(NA[1], r0y, NA[0] * r0x));
    o->hpB = ORC_FMA(NB[2], r1z, ORC_FMA(NB[1], r1y, NB[0] * r1x));
    o->hA0 = ORC_FMA(nA[2], r0z, ORC_FMA(nA[1], r0y, nA[0] * r0x));
    o->hA1 = ORC_FMA(nA[5], r1z, ORC_FMA(nA[4], r1y, nA[3] * r1x));
    o->hA2 = ORC_FMA(nA[8], r2z, ORC_FMA(nA[7], r2y, nA[6] * r2x));
    o->hB0 = ORC_FMA(nB[2], r1z, ORC_FMA(nB[1], r1y, nB[0] * r1x));
    o->hB1 = ORC_FMA(nB[5], r0z, ORC_FMA(nB[4], r0y, nB[3] * r0x));
    o->hB2 = ORC_FMA(nB[8], r3z, ORC_FMA(nB[7], r3y, nB[6] * r3x));
    const double a0 = l0 + l1, a1 = l1 + l2, a2 = l2 + l0, a3 = l0 + l3, a4 = l3 + l1;
    const double d01 = ORC_FMA(r0z, r1z, ORC_FMA(r0y, r1y, r0x * r1x));
    const double d12 = ORC_FMA(r1z, r2z, ORC_FMA(r1y, r2y, r1x * r2x));
    const double d20 = ORC_FMA(r2z, r0z, ORC_FMA(r2y, r0y, r2x * r0x));
    const double d03 = ORC_FMA(r0z, r3z, ORC_FMA(r0y, r3y, r0x * r3x));
    const double d31 = ORC_FMA(r3z, r1z, ORC_FMA(r3y, r1y, r3x * r1x));
    const double l01 = l0 * l1;
    const double denA = ORC_FMA(l2, d01, ORC_FMA(l1, d20, ORC_FMA(l0, d12, l01 * l2)));
    const double denB = ORC_FMA(l3, d01, ORC_FMA(l0, d31, ORC_FMA(l1, d03, l01 * l3)));
    const double numA = o->hpA * rc[41], numB = o->hpB * rc[42];
    /* One division serves the five edge quotients and the two solid-angle quotients (X Y is a product of lengths
     * and of the two solid-angle denominators: it only has to be non-zero; were a denominator exactly 0 every
     * quotient would come out non-finite, fail its range test and take its own way). */
    const double P01 = a0 * a1, P2A = a2 * denA, P34 = a3 * a4;
    const double X = P01 * P2A, Y = P34 * denB;
    const double inv = 1.0 / (X * Y);
    const double iX = Y * inv, iY = X * inv;
    const double i01 = P2A * iX, i2A = P01 * iX, i34 = denB * iY, iB = P34 * iY;
    o->z[0] = (G[0] * a1) * i01;
    o->z[1] = (G[1] * a0) * i01;
    o->z[2] = (G[2] * denA) * i2A;
    o->q[0] = (numA * a2) * i2A;
    o->z[3] = (G[3] * a4) * i34;
    o->z[4] = (G[4] * a3) * i34;
    o->q[1] = numB * iB;
    for (int k = 0; k < 5; ++k) o->zs[k] = o->z[k] * o->z[k];
    o->qs[0] = o->q[0] * o->q[0];
    o->qs[1] = o->q[1] * o->q[1];
    o->a[0] = a0, o->a[1] = a1, o->a[2] = a2, o->a[3] = a3, o->a[4] = a4;
    o->num[0] = numA, o->num[1] = numB, o->den[0] = denA, o->den[1] = denB;
}

/* Tier of a pair: 2 = far (every z^2 and q^2 below 0x3F647AE1'00000000 = 0.0025 - 5e-10, both faces seen from the
 * front), 1 = middle (every z^2 and q^2 below 0x3F847AE1'00000000 = 0.01 - 2.1e-9, faces from the front), 0 = neither. */
static inline int pair_tier(const pair_pre *o)
{
    int32_t h = orc_hi(o->qs[0]) > orc_hi(o->qs[1]) ? orc_hi(o->qs[0]) : orc_hi(o->qs[1]);
    for (int k = 0; k < 5; ++k)
        if (orc_hi(o->zs[k]) > h) h = orc_hi(o->zs[k]);
    if (!(orc_hi(o->den[0]) > 0 && orc_hi(o->den[1]) > 0)) return 0;
    return h < ORC_HI_FAR ? 2 : (h < ORC_HI_Q2MAX ? 1 : 0);
}

/* The transcendental part.  `row_tier` = the lowest tier among the pairs of this pair's ROW (32 consecutive records --
 * what one warp of the CUDA kernels evaluates together): 2 -> 4-coefficient polynomials, 1 -> 5-coefficient atanh and
 * the 5-coefficient atan.  Otherwise each term chooses for itself: the
 * thresholds are doubles whose low word is 0 (0x3FC47AE1'00000000 = 0.16 - 3.3e-8 for z^2, 0x3F847AE1'00000000 =
 * 0.01 - 2.1e-9 for q^2), so for x >= 0 `x < threshold` is a comparison of the high words (the CUDA side does it on
 * the integer pipe; a NaN fails it; `den > 0` likewise reads hi(den) > 0): inside -> 11 / 5-coefficient minimax
 * polynomial; an edge seen under a larger angle takes a division of its own and the ln-based atanh, a face seen under
 * a larger solid angle (or from behind) the full-range atan2. */
static inline void pair_finish(const pair_pre *o, int row_tier, double *ax, double *ay, double *az, double *pv)
{
    const double *rc = o->rc, *NA = rc + 12, *NB = rc + 15, *G = rc + 36;
    double ln[5], om[2];
    if (row_tier == 2) {
        for (int k = 0; k < 5; ++k) ln[k] = orc_two_atanh_far(o->z[k], o->zs[k]);
        for (int f = 0; f < 2; ++f) om[f] = orc_two_atan_far(o->q[f], o->qs[f]);
    } else if (row_tier == 1) {
        for (int k = 0; k < 5; ++k) ln[k] = orc_two_atanh_mid(o->z[k], o->zs[k]);
        for (int f = 0; f < 2; ++f) om[f] = orc_two_atan_mm(o->q[f], o->qs[f]);
    } else {
        for (int k = 0; k < 5; ++k)
            ln[k] = (orc_hi(o->zs[k]) < ORC_HI_Z2MAX) ? orc_two_atanh_wide(o->z[k], o->zs[k]) : orc_two_atanh(G[k] / o->a[k]);
        for (int f = 0; f < 2; ++f) {
            const int ok = (orc_hi(o->den[f]) > 0) && (orc_hi(o->qs[f]) < ORC_HI_Q2MAX);
            om[f] = ok ? orc_two_atan_mm(o->q[f], o->qs[f]) : 2.0 * orc_atan2(o->num[f], o->den[f]);
        }
    }
    const double SA = ORC_FMA(-o->hpA, om[0], ORC_FMA(o->hA2, ln[2], ORC_FMA(o->hA1, ln[1], o->hA0 * ln[0])));
    const double SB = ORC_FMA(-o->hpB, om[1], ORC_FMA(o->hB2, ln[4], ORC_FMA(o->hB1, ln[3], o->hB0 * ln[0])));
    *ax = ORC_FMA(NB[0], SB, ORC_FMA(NA[0], SA, *ax));
    *ay = ORC_FMA(NB[1], SB, ORC_FMA(NA[1], SA, *ay));
    *az = ORC_FMA(NB[2], SB, ORC_FMA(NA[2], SA, *az));
    *pv = ORC_FMA(o->hpB, SB, ORC_FMA(o->hpA, SA, *pv));
}

/* Sum over pair records in the order of a 32-lane warp: lane l accumulates records l, l+32, ... in increasing
 * order, then an xor-butterfly adds the 32 partial sums (polyhedral_gravity itself reduces with an
 * unspecified parallel order, so any fixed order is as faithful as another).  The records are taken a ROW of 32 at a
 * time, as a warp does: the row's tier (far / middle / neither) is decided from all of its pairs (the padding records
 * a warp sees in the last row are far by construction). */
static void gravity_point(const orc_mesh *m, const double P[3], double acc[3], double *pot)
{
    double sx[32] = {0.0}, sy[32] = {0.0}, sz[32] = {0.0}, sv[32] = {0.0};
    pair_pre pre[32];
    for (int base = 0; base < m->n_pairs; base += 32) {
        const int n = m->n_pairs - base < 32 ? m->n_pairs - base : 32;
        int row_tier = 2;
        for (int i = 0; i < n; ++i) {
            pair_prelim(m->prec + (size_t)(base + i) * PREC, P, &pre[i]);
            const int t = pair_tier(&pre[i]);
            if (t < row_tier) row_tier = t;
        }
        for (int i = 0; i < n; ++i) pair_finish(&pre[i], row_tier, &sx[i], &sy[i], &sz[i], &sv[i]);
    }
    const double ax = orc_butterfly32(sx), ay = orc_butterfly32(sy), az = orc_butterfly32(sz);
    acc[0] = m->Grho * ax;
    acc[1] = m->Grho * ay;
    acc[2] = m->Grho * az;
    if (pot) *pot = 0.5 * m->Grho * orc_butterfly32(sv);
}

/* Census of the paths gravity_point takes for the given points (tests: every tier must have been exercised):
 * out[0] far rows, out[1] middle rows, out[2] rows where every term chooses for itself, out[3] edge terms that took
 * the ln-based atanh, out[4] solid-angle terms that took the full-range atan2. */
void orc_gravity_tiers(const orc_mesh *m, const double *pts, int64_t n, int64_t out[5])
{
    pair_pre pre[32];
    for (int k = 0; k < 5; ++k) out[k] = 0;
    for (int64_t ip = 0; ip < n; ++ip) {
        for (int base = 0; base < m->n_pairs; base += 32) {
            const int cnt = m->n_pairs - base < 32 ? m->n_pairs - base : 32;
            int row_tier = 2;
            for (int i = 0; i < cnt; ++i) {
                pair_prelim(m->prec + (size_t)(base + i) * PREC, pts + 3 * ip, &pre[i]);
                const int t = pair_tier(&pre[i]);
                if (t < row_tier) row_tier = t;
            }
            out[2 - row_tier] += 1;
            if (row_tier == 0)
                for (int i = 0; i < cnt; ++i) {
                    for (int k = 0; k < 5; ++k) out[3] += !(orc_hi(pre[i].zs[k]) < ORC_HI_Z2MAX);
                    for (int f = 0; f < 2; ++f)
                        out[4] += !((orc_hi(pre[i].den[f]) > 0) && (orc_hi(pre[i].qs[f]) < ORC_HI_Q2MAX));
                }
        }
    }
}

void orc_gravity_direct(const orc_mesh *m, const double *pts, int64_t n, double *acc, double *pot)
{
    for (int64_t i = 0; i < n; ++i) gravity_point_direct(m, pts + 3 * i, acc + 3 * i, pot ? pot + i : NULL);
}

typedef struct {
    const orc_mesh *m;
    const double *pts;
    double *acc, *pot;
} grav_job;
static void grav_body(int64_t i, void *v)
{
    grav_job *g = (grav_job *)v;
    gravity_point(g->m, g->pts + 3 * i, g->acc + 3 * i, g->pot ? g->pot + i : NULL);
}
void orc_gravity(const orc_mesh *m, const double *pts, int64_t n, double *acc, double *pot, int threads)
{
    grav_job g = {m, pts, acc, pot};
    parallel_for(n, 64, threads, grav_body, &g);
}

/* equations_of_motion.py:98-117: Quaternion(axis, angle=-(w t)).rotate(x); pyquaternion
 * normalises the axis.  Euler-Rodrigues form (tests/point_rotation_test.py:47-76). */
void orc_rotate_point(double t, const double x[3], const double axis[3], double w, double out[3])
{
    double na = norm3(axis);
    double k[3] = {axis[0] / na, axis[1] / na, axis[2] / na};
    double th = -(w * t);
    double c, s;
    orc_sincos(th, &s, &c);
    double kx[3];
    cross3(k, x, kx);
    double kd = dot3(k, x) * (1.0 - c);
    for (int i = 0; i < 3; ++i) out[i] = x[i] * c + kx[i] * s + k[i] * kd;
}

/* equations_of_motion.py:62-95: acceleration evaluated at the rotated position and NOT
 * rotated back (:84-95); sign flip of the package value (:59). */
void orc_compute_motion(const orc_mesh *m, const orc_params *p, double t, const double y[6], double out[6])
{
    double pos[3] = {y[0], y[1], y[2]};
    if (p->activate_rotation) orc_rotate_point(t, y, p->spin_axis, p->spin_velocity, pos);
    double a[3];
    gravity_point(m, pos, a, NULL);
    out[0] = y[3];
    out[1] = y[4];
    out[2] = y[5];
    out[3] = -a[0];
    out[4] = -a[1];
    out[5] = -a[2];
}

/* mesh_utility.py:132-166 Moeller-Trumbore with ray direction (0,0,1) (:83), one point. */
static int ray_hits(const double *o, const double *v0, const double *v1, const double *v2)
{
    double e1[3], e2[3];
    for (int c = 0; c < 3; ++c) {
        e1[c] = v1[c] - v0[c];
        e2[c] = v2[c] - v0[c];
    }
    const double d[3] = {0.0, 0.0, 1.0};
    double h[3];
    cross3(d, e2, h);
    double a = dot3(e1, h);
    if (a < 0.000001 && a > -0.000001) return 0;
    double f = 1.0 / a;
    double s[3] = {o[0] - v0[0], o[1] - v0[1], o[2] - v0[2]};
    double u = dot3(s, h) * f;
    if (u < 0.0 || u > 1.0) return 0;
    double q[3];
    cross3(s, e1, q);
    double v = dot3(q, d) * f;
    if (v < 0.0 || u + v > 1.0) return 0;
    double t = f * dot3(q, e2);
    return t > 0.0;
}

static int point_outside(const orc_mesh *m, const double *pt)
{
    int counter = 0;
    for (int f = 0; f < m->F; ++f) {
        const double *rc = m->rec + (size_t)f * REC;
        counter += ray_hits(pt, rc, rc + 3, rc + 6);
    }
    return (counter % 2) == 0; /* mesh_utility.py:89 */
}

void orc_is_outside(const orc_mesh *m, const double *pts, int64_t n, uint8_t *out)
{
    for (int64_t i = 0; i < n; ++i) out[i] = (uint8_t)point_outside(m, pts + 3 * i);
}

/* ---------------------------------------------------------------------------------------
 * desolver OdeSystem.integrate with method RK8713MSolver (call site compute_trajectory.py:243-259).
 * Step controller (confirmed against tests/fixed_step_trajectory_test.py:44-47
 * to 7e-5 m, where any other safety factor/exponent misses by > 1e-3 m):
 *    err  = max_i |h sum_j (b8_j - b7_j) k_j|
 *    tol  = max_i (atol + rtol |y_i| + rtol |dy_i / h|)
 *    h'   = 0.8 h (tol/err)^(1/8)            (applied after accepted and rejected attempts alike)
 *    redo while err > tol; last step clipped to the interval end.
 * Event (compute_trajectory.py:264-282, non-terminal :258): every accepted state with
 * |r|^2 <= R_in^2 is recorded (App. A.2 hypothesis); its position is tested by is_outside.
 * ------------------------------------------------------------------------------------- */
typedef struct {
    double *t, *y, *f;
    int n, cap;
} knotbuf;

static int push_knot(knotbuf *k, double t, const double *y, const double *f)
{
    if (k->n >= k->cap) return -1;
    k->t[k->n] = t;
    memcpy(k->y + 6 * (size_t)k->n, y, 6 * sizeof(double));
    memcpy(k->f + 6 * (size_t)k->n, f, 6 * sizeof(double));
    k->n++;
    return 0;
}

static int integrate_segment(const orc_mesh *m, const orc_params *p, double t0, double tf, const double y0[6],
                             knotbuf *kn, orc_counters *cnt, int *collided, double *event_pts, int event_cap)
{
    double t = t0, y[6], f0[6], k[13][6];
    memcpy(y, y0, sizeof(y));
    orc_compute_motion(m, p, t, y, f0);
    cnt->n_rhs++;
    if (push_knot(kn, t, y, f0)) return -1;
    double dt = p->initial_time_step;
    int steps = 0;
    const double R2 = p->radius_inner * p->radius_inner;
    while (t < tf) {
        int clipped = 0;
        if (t + dt > tf) {
            dt = tf - t;
            clipped = 1;
        }
        double dy[6], newdt;
        for (;;) {
            memcpy(k[0], f0, sizeof(f0));
            for (int i = 1; i < 13; ++i) {
                double ys[6];
                for (int c = 0; c < 6; ++c) {
                    double s = 0.0;
                    for (int j = 0; j < i; ++j) s += ORC_RK_A[13 * i + j] * k[j][c];
                    ys[c] = y[c] + dt * s;
                }
                orc_compute_motion(m, p, t + ORC_RK_C[i] * dt, ys, k[i]);
                cnt->n_rhs++;
            }
            double err = 0.0, tol = 0.0;
            int finite = 1;
            for (int c = 0; c < 6; ++c) {
                double s8 = 0.0, se = 0.0;
                for (int j = 0; j < 13; ++j) {
                    s8 += ORC_RK_B8[j] * k[j][c];
                    se += ORC_RK_E[j] * k[j][c];
                }
                dy[c] = dt * s8;
                double e = fabs(dt * se);
                double tl = p->atol + p->rtol * fabs(y[c]) + p->rtol * fabs(s8);
                if (e > err) err = e;
                if (tl > tol) tol = tl;
                if (!isfinite(e) || !isfinite(tl)) finite = 0; /* a NaN would lose every comparison above */
            }
            double corr = 1.0;
            if (err != 0.0) corr = 0.8 * orc_root8(tol / err);
            newdt = (corr != 0.0) ? corr * dt : dt;
            if (!finite) { /* NaN / Inf state */
                cnt->status = 2;
                return 0;
            }
            if (err > tol) {
                cnt->n_rejected++;
                dt = newdt;
                clipped = 0;
                if (++steps > p->max_steps || !(dt > 0.0)) {
                    cnt->status = 1;
                    return 0;
                }
                continue;
            }
            break;
        }
        t = clipped ? tf : t + dt;
        for (int c = 0; c < 6; ++c) y[c] += dy[c];
        dt = newdt;
        cnt->n_accepted++;
        orc_compute_motion(m, p, t, y, f0);
        cnt->n_rhs++;
        if (push_knot(kn, t, y, f0)) return -1;
        if (p->activate_event && (y[0] * y[0] + y[1] * y[1] + y[2] * y[2]) <= R2) {
            if (event_pts && cnt->n_event_points < event_cap)
                memcpy(event_pts + 3 * (size_t)cnt->n_event_points, y, 3 * sizeof(double));
            cnt->n_event_points++;
            if (!point_outside(m, y)) {
                *collided = 1;
                if (p->stop_on_collision) return 0;
            }
        }
        if (++steps > p->max_steps) {
            cnt->status = 1;
            return 0;
        }
    }
    return 0;
}

/* compute_trajectory.py:53-121: decode, int32-truncate, sort, merge simultaneous manoeuvres. */
static int decode_maneuvers(const orc_params *p, const double *x, int M, double *times /*M+2*/, double *dv /*3 x M*/)
{
    int32_t tm[64];
    double dvv[64][3];
    int idx[64];
    if (M > 64) M = 64;
    for (int k = 0; k < M; ++k) {
        const double *mm = x + 7 + 5 * k;
        tm[k] = (int32_t)mm[0]; /* astype(np.int32): truncation toward zero (:87) */
        for (int c = 0; c < 3; ++c) dvv[k][c] = mm[1] * mm[2 + c];
        idx[k] = k;
    }
    /* np.argsort on a short int array: insertion sort, stable (:93) */
    for (int i = 1; i < M; ++i) {
        int j = i, v = idx[i];
        while (j > 0 && tm[idx[j - 1]] > tm[v]) {
            idx[j] = idx[j - 1];
            --j;
        }
        idx[j] = v;
    }
    int n = 0;
    times[0] = (double)(int32_t)p->start_time;
    for (int i = 0; i < M; ++i) {
        int k = idx[i];
        if (n > 0 && (double)tm[k] == times[n]) { /* :101-114 equal times: dv summed */
            for (int c = 0; c < 3; ++c) dv[3 * (n - 1) + c] += dvv[k][c];
        } else {
            times[n + 1] = (double)tm[k];
            for (int c = 0; c < 3; ++c) dv[3 * n + c] = dvv[k][c];
            ++n;
        }
    }
    times[n + 1] = (double)(int32_t)p->final_time;
    return n; /* number of distinct manoeuvres; segments = n + 1 */
}

int orc_compute_trajectory(const orc_mesh *m, const orc_params *p, const double *x, int n_man,
                           double *knots_t, double *knots_y, double *knots_f, int cap,
                           int *seg_start, double *intervals, int *n_seg, int *n_knots,
                           int *collision_detected, double *event_pts, int event_cap, orc_counters *cnt)
{
    double times[66], dv[3 * 64];
    int nm = decode_maneuvers(p, x, n_man, times, dv);
    knotbuf kn = {knots_t, knots_y, knots_f, 0, cap};
    orc_counters local;
    if (!cnt) cnt = &local;
    memset(cnt, 0, sizeof(*cnt));
    double state[6] = {x[0], x[1], x[2], x[3] * x[4], x[3] * x[5], x[3] * x[6]}; /* :53-55 */
    int collided = 0, rc = 0, segs = 0;
    for (int s = 0; s <= nm; ++s) {
        if (s > 0) { /* :134-136 velocity REPLACED */
            state[3] = dv[3 * (s - 1)];
            state[4] = dv[3 * (s - 1) + 1];
            state[5] = dv[3 * (s - 1) + 2];
        }
        if (seg_start) seg_start[s] = kn.n;
        if (intervals) intervals[s] = times[s];
        rc = integrate_segment(m, p, times[s], times[s + 1], state, &kn, cnt, &collided, event_pts, event_cap);
        segs = s + 1;
        if (rc) break;
        memcpy(state, kn.y + 6 * (size_t)(kn.n - 1), 6 * sizeof(double)); /* :138 */
        if (cnt->status || (collided && p->stop_on_collision)) break;
    }
    if (seg_start) seg_start[segs] = kn.n;
    if (intervals) intervals[segs] = times[segs];
    if (n_seg) *n_seg = segs;
    if (n_knots) *n_knots = kn.n;
    if (collision_detected) *collision_detected = collided;
    return rc;
}

/* desolver CubicHermiteInterp between consecutive accepted steps. */
static void hermite(double t0, double t1, const double *p0, const double *p1, const double *m0, const double *m1,
                    double t, double *out)
{
    double h = t1 - t0;
    double s = (t - t0) / h, s2 = s * s, s3 = s2 * s;
    double h00 = 2.0 * s3 - 3.0 * s2 + 1.0, h10 = s3 - 2.0 * s2 + s, h01 = -2.0 * s3 + 3.0 * s2, h11 = s3 - s2;
    for (int c = 0; c < 6; ++c) out[c] = h00 * p0[c] + h10 * h * m0[c] + h01 * p1[c] + h11 * h * m1[c];
}

void orc_dense_eval(const double *kt, const double *ky, const double *kf, int n, const double *t, int64_t nt,
                    double *out)
{
    for (int64_t i = 0; i < nt; ++i) {
        if (n < 2) {
            memcpy(out + 6 * i, ky, 6 * sizeof(double));
            continue;
        }
        int lo = 0, hi = n - 1; /* interval j with kt[j] <= t <= kt[j+1] (first such from the left end) */
        while (hi - lo > 1) {
            int mid = (lo + hi) / 2;
            if (kt[mid] < t[i])
                lo = mid;
            else
                hi = mid;
        }
        hermite(kt[lo], kt[lo + 1], ky + 6 * (size_t)lo, ky + 6 * (size_t)(lo + 1), kf + 6 * (size_t)lo,
                kf + 6 * (size_t)(lo + 1), t[i], out + 6 * i);
    }
}

int64_t orc_num_samples(const orc_params *p)
{ /* len(np.arange(start, final, period)) */
    double n = ceil((p->final_time - p->start_time) / p->measurement_period);
    return n > 0 ? (int64_t)n : 0;
}

/* trajectory_tools.py:46-74 */
void orc_fixed_step(const orc_params *p, const double *knots_t, const double *knots_y, const double *knots_f,
                    const int *seg_start, int n_seg, double *pos, double *vel, double *ts)
{
    int64_t N = orc_num_samples(p);
    for (int64_t i = 0; i < N; ++i) ts[i] = p->start_time + (double)i * p->measurement_period;
    int64_t start_idx = 0;
    for (int s = 0; s < n_seg; ++s) {
        int k0 = seg_start[s], k1 = seg_start[s + 1];
        if (k1 <= k0) continue;
        double t_end = knots_t[k1 - 1];
        /* :55-58 nearest sample, minus one if later than the segment end */
        int64_t e = 0;
        double best = fabs(ts[0] - t_end);
        for (int64_t i = 1; i < N; ++i) {
            double d = fabs(ts[i] - t_end);
            if (d < best) {
                best = d;
                e = i;
            }
        }
        if (t_end < ts[e]) e -= 1;
        if (e >= start_idx) {
            double out[6];
            for (int64_t i = start_idx; i <= e; ++i) {
                orc_dense_eval(knots_t + k0, knots_y + 6 * (size_t)k0, knots_f + 6 * (size_t)k0, k1 - k0, ts + i, 1, out);
                for (int c = 0; c < 3; ++c) {
                    pos[c * N + i] = out[c];
                    vel[c * N + i] = out[3 + c];
                }
            }
        }
        start_idx = e + 1; /* :70 */
        if (N - 1 < start_idx) break;
    }
}

/* ------------------------------- fitness ------------------------------------------- */

/* fitness_function_utils.py:282-314 */
static void cart2sphere(const double *x, double *r, double *th, double *ph)
{
    *r = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    *th = orc_atan2(x[2], sqrt(x[0] * x[0] + x[1] * x[1]));
    *ph = orc_atan2(x[1], x[0]);
}

static int argmin_abs(const double *g, int n, double v)
{ /* np.argmin(|grid - v|): first minimum wins */
    int best = 0;
    double bd = fabs(g[0] - v);
    for (int i = 1; i < n; ++i) {
        double d = fabs(g[i] - v);
        if (d < bd) {
            bd = d;
            best = i;
        }
    }
    return best;
}

/* rotation clock of sample n: np.array_split(positions, k, axis=1) then timesteps[col] with
 * col counted from the start of the chunk (fitness_function_utils.py:96-102). */
static int64_t chunk_col(int64_t n, int64_t N, int k)
{
    if (k <= 1) return n;
    int64_t base = N / k, rem = N % k; /* first rem chunks have base+1 columns */
    int64_t big = rem * (base + 1);
    if (n < big) return n % (base + 1);
    if (base == 0) return 0;
    return (n - big) % base;
}

static double coverage_core(int nsc, const double axis[3], double w, const double *pos, int64_t N, const double *ts,
                            double rmin, double rmax, const double *gr, int nr, const double *gth, int nth,
                            const double *gph, int nph, const uint8_t *bool_tensor, int single,
                            int64_t *n_visited, uint8_t *out_tensor)
{
    size_t nvox = (size_t)nr * nth * nph;
    uint8_t *nt = out_tensor ? out_tensor : (uint8_t *)malloc(nvox);
    for (size_t i = 0; i < nvox; ++i) nt[i] = bool_tensor[i] ? 1 : 0;
    int64_t nsamp = single ? 1 : N;
    for (int64_t n = 0; n < nsamp; ++n) {
        double x[3], xr[3], r, th, ph;
        if (single) {
            x[0] = pos[0];
            x[1] = pos[1];
            x[2] = pos[2];
            orc_rotate_point(ts[0], x, axis, w, xr);
        } else {
            x[0] = pos[n];
            x[1] = pos[N + n];
            x[2] = pos[2 * N + n];
            orc_rotate_point(ts[chunk_col(n, N, nsc)], x, axis, w, xr);
        }
        cart2sphere(xr, &r, &th, &ph);
        if (!(r <= rmax) || !(r >= rmin)) continue; /* :113-122 inclusive */
        int i = argmin_abs(gr, nr, r), j = argmin_abs(gth, nth, th), k = argmin_abs(gph, nph, ph);
        nt[((size_t)i * nth + j) * nph + k] = 1;
    }
    /* :150-161 (and the empty branch :125-133, which is the same expression on bool_tensor) */
    int64_t total = 0;
    double wsum = 0.0;
    for (int i = 0; i < nr; ++i) {
        int64_t c = 0;
        for (size_t q = 0; q < (size_t)nth * nph; ++q) c += nt[(size_t)i * nth * nph + q];
        total += c;
        wsum += (double)c / gr[i];
    }
    if (n_visited) *n_visited = total;
    if (!out_tensor) free(nt);
    return (double)total / (double)nvox + wsum;
}

double orc_space_coverage(int n_spacecrafts, const double axis[3], double w, const double *pos, int64_t N,
                          const double *ts, double rmin, double rmax, const double *gr, int nr, const double *gth,
                          int nth, const double *gph, int nph, const uint8_t *bool_tensor, int single,
                          int64_t *n_visited, uint8_t *new_tensor)
{
    return coverage_core(n_spacecrafts, axis, w, pos, N, ts, rmin, rmax, gr, nr, gth, nth, gph, nph, bool_tensor,
                         single, n_visited, new_tensor);
}

void orc_update_tensor(int n_spacecrafts, const double axis[3], double w, const double *pos, int64_t N,
                       const double *ts, double rmin, double rmax, const double *gr, int nr, const double *gth,
                       int nth, const double *gph, int nph, uint8_t *bool_tensor)
{
    size_t nvox = (size_t)nr * nth * nph;
    uint8_t *tmp = (uint8_t *)malloc(nvox);
    coverage_core(n_spacecrafts, axis, w, pos, N, ts, rmin, rmax, gr, nr, gth, nth, gph, nph, bool_tensor, 0, NULL,
                  tmp);
    memcpy(bool_tensor, tmp, nvox);
    free(tmp);
}

/* fitness_function_utils.py:41-51 */
static inline double sqdist(const double *pos, int64_t N, int64_t n, double c)
{
    return pos[n] * pos[n] + pos[N + n] * pos[N + n] + pos[2 * N + n] * pos[2 * N + n] - c * c;
}

/* reductions over samples run in the order of a 256-thread block (orc_block_sum256): numpy itself sums
 * pairwise in blocks, so no particular order is "the reference's" -- this one is the GPU's. */
typedef struct {
    const double *pos;
    int64_t N;
    double c;
    int mode; /* 0: |d|, 1: d where d > 0 else 0 */
} sq_job;
static double sq_item(int64_t n, void *v)
{
    const sq_job *j = (const sq_job *)v;
    const double d = sqdist(j->pos, j->N, n, j->c);
    if (j->mode == 0) return fabs(d);
    return d > 0.0 ? d : 0.0;
}

/* fitness_functions.py:48-59 */
static double target_altitude_distance(double tsa, const double *pos, int64_t N)
{
    sq_job j = {pos, N, sqrt(tsa), 0};
    const double s = orc_block_sum256(N, sq_item, &j);
    return orc_root4((s / (double)N) / tsa);
}
/* fitness_functions.py:62-85 */
static double close_distance_penalty(double rin, const double *pos, int64_t N, double scale)
{
    double mn = 0.0;
    int any = 0;
    for (int64_t n = 0; n < N; ++n) {
        double d = sqdist(pos, N, n, rin);
        if (d < 0.0 && (!any || d < mn)) {
            mn = d;
            any = 1;
        }
    }
    if (!any) return 0.0;
    return orc_root4(fabs(mn) / (rin * rin)) * scale;
}
/* fitness_functions.py:88-113 */
static double far_distance_penalty(double rout, const double *pos, int64_t N, double scale)
{
    int64_t cnt = 0;
    for (int64_t n = 0; n < N; ++n) cnt += sqdist(pos, N, n, rout) > 0.0;
    if (!cnt) return 0.0;
    sq_job j = {pos, N, rout, 1};
    const double s = orc_block_sum256(N, sq_item, &j);
    return orc_root4((s / (double)cnt) / (rout * rout)) * scale;
}
/* fitness_function_utils.py:8-38 */
typedef struct {
    const double *pos;
    int64_t N;
} vol_job;
static double vol_item(int64_t n, void *v)
{
    const vol_job *q = (const vol_job *)v;
    const double *pos = q->pos;
    const int64_t N = q->N;
    /* radii[1:] = dist/2 ; radii[0] = radii[1] ; radii[-1] = radii[-2] */
    int64_t j = n;
    if (j == 0) j = 1;
    if (n == N - 1) j = (N - 2 >= 1) ? N - 2 : 1;
    double dx = pos[j] - pos[j - 1], dy = pos[N + j] - pos[N + j - 1], dz = pos[2 * N + j] - pos[2 * N + j - 1];
    double rad = sqrt(dx * dx + dy * dy + dz * dz) / 2.0;
    return (4.0 / 3.0) * M_PI * (rad * rad * rad);
}
static double estimate_covered_volume(const double *pos, int64_t N)
{
    if (N < 2) return 0.0;
    vol_job j = {pos, N};
    return orc_block_sum256(N, vol_item, &j);
}

double orc_get_fitness(int fn, const orc_params *p, const double *pos, int64_t N, const double *ts, const double *gr,
                       int nr, const double *gth, int nth, const double *gph, int nph, const uint8_t *bool_tensor,
                       int64_t *n_visited)
{
    double rin = p->radius_inner, rout = p->radius_outer, sc = p->penalty_scaling_factor;
    switch (fn) {
    case 1: return target_altitude_distance(p->target_squared_altitude, pos, N);
    case 2: return close_distance_penalty(rin, pos, N, sc);
    case 3: return far_distance_penalty(rout, pos, N, sc);
    case 4: return -(estimate_covered_volume(pos, N) / (p->maximal_measurement_sphere_volume * (double)N));
    case 5: return -(estimate_covered_volume(pos, N) / p->total_measurable_volume);
    case 6:
        return -(estimate_covered_volume(pos, N) / (p->maximal_measurement_sphere_volume * (double)N)) +
               far_distance_penalty(rout, pos, N, sc);
    case 7:
        return -(estimate_covered_volume(pos, N) / (p->maximal_measurement_sphere_volume * (double)N)) +
               close_distance_penalty(rin, pos, N, sc) + far_distance_penalty(rout, pos, N, sc);
    case 8:
        return coverage_core(p->number_of_spacecrafts, p->spin_axis, p->spin_velocity, pos, N, ts, rin, rout, gr, nr,
                             gth, nth, gph, nph, bool_tensor, 0, n_visited, NULL);
    case 9: { /* fitness_functions.py:198-227 */
        double cov = coverage_core(p->number_of_spacecrafts, p->spin_axis, p->spin_velocity, pos, N, ts, rin, rout, gr,
                                   nr, gth, nth, gph, nph, bool_tensor, 0, n_visited, NULL);
        return close_distance_penalty(rin, pos, N, sc) + far_distance_penalty(rout, pos, N, sc) - cov;
    }
    default: return NAN;
    }
}

/* udp_initial_condition.py:106-137 */
static double udp_fitness(const orc_mesh *m, const orc_params *p, const double *gr, int nr, const double *gth, int nth,
                          const double *gph, int nph, const uint8_t *bool_tensor, const double *x, int n_man,
                          int32_t *collision, orc_counters *cnt, int64_t *n_visited)
{
    int cap = p->max_steps + 8 * (n_man + 2);
    double *kt = (double *)malloc(sizeof(double) * (size_t)cap * 13);
    double *ky = kt + cap, *kf = ky + 6 * (size_t)cap;
    int seg_start[70], n_seg = 0, n_knots = 0, coll = 0;
    double intervals[70];
    orc_counters c;
    orc_compute_trajectory(m, p, x, n_man, kt, ky, kf, cap, seg_start, intervals, &n_seg, &n_knots, &coll, NULL, 0, &c);
    double fit;
    if (n_visited) *n_visited = -1;
    if (coll || c.status) {
        fit = 1e30; /* :126-128 (and non-finite / step-cap trajectories) */
    } else {
        int64_t N = orc_num_samples(p);
        double *pos = (double *)malloc(sizeof(double) * 7 * (size_t)N);
        double *vel = pos + 3 * N, *ts = vel + 3 * N;
        orc_fixed_step(p, kt, ky, kf, seg_start, n_seg, pos, vel, ts);
        fit = orc_get_fitness(9, p, pos, N, ts, gr, nr, gth, nth, gph, nph, bool_tensor, n_visited);
        free(pos);
    }
    free(kt);
    if (collision) *collision = coll;
    if (cnt) *cnt = c;
    return fit;
}

typedef struct {
    const orc_mesh *m;
    const orc_params *p;
    const double *gr, *gth, *gph;
    int nr, nth, nph;
    const uint8_t *bool_tensor;
    const double *xs, *fixed;
    int dim, n_fixed, n_man;
    double *fitness;
    int32_t *collision;
    orc_counters *cnt;
    int64_t *n_visited;
} pop_job;
static void pop_body(int64_t i, void *v)
{
    pop_job *j = (pop_job *)v;
    double x[7 + 5 * 64];
    for (int c = 0; c < j->n_fixed; ++c) x[c] = j->fixed[c]; /* np.hstack((initial_condition, chromosome)) :117-120 */
    for (int c = 0; c < j->dim; ++c) x[j->n_fixed + c] = j->xs[(size_t)i * j->dim + c];
    j->fitness[i] = udp_fitness(j->m, j->p, j->gr, j->nr, j->gth, j->nth, j->gph, j->nph, j->bool_tensor, x, j->n_man,
                                j->collision ? j->collision + i : NULL, j->cnt ? j->cnt + i : NULL,
                                j->n_visited ? j->n_visited + i : NULL);
}
void orc_population_fitness(const orc_mesh *m, const orc_params *p, const double *gr, int nr, const double *gth,
                            int nth, const double *gph, int nph, const uint8_t *bool_tensor, const double *xs, int pop,
                            int dim, const double *fixed, int n_fixed, int n_man, double *fitness, int32_t *collision,
                            orc_counters *cnt, int64_t *n_visited, int threads)
{
    pop_job j = {m, p, gr, gth, gph, nr, nth, nph, bool_tensor, xs, fixed, dim, n_fixed, n_man,
                 fitness, collision, cnt, n_visited};
    parallel_for(pop, 1, threads, pop_body, &j);
}

/* elementary functions of orc_math.h, exposed for tests/test_oracle_math.py and the GPU bit-parity tests.
 * fn: 0 two_atanh(a) 1 log(a) 2 atan2(a,b) 3 sincos(a)->(out,out2) 4 root8(a) 5 root4(a) 6 atan_series8(a)
 *     7 a/b 8 sqrt(a) 9 / 10 minimax 2 atanh / 2 atan of the common path 11 / 12 far tier 13 middle tier 14 wide 2 atanh */
void orc_math_probe(int fn, const double *a, const double *b, int64_t n, double *out, double *out2)
{
    for (int64_t i = 0; i < n; ++i) {
        switch (fn) {
        case 0: out[i] = orc_two_atanh(a[i]); break;
        case 1: out[i] = orc_log(a[i]); break;
        case 2: out[i] = orc_atan2(a[i], b[i]); break;
        case 3: orc_sincos(a[i], &out[i], &out2[i]); break;
        case 4: out[i] = orc_root8(a[i]); break;
        case 5: out[i] = orc_root4(a[i]); break;
        case 6: out[i] = orc_atan_series(a[i], 8); break;
        case 7: out[i] = a[i] / b[i]; break;
        case 8: out[i] = sqrt(a[i]); break;
        case 9: out[i] = orc_two_atanh_mm(a[i], a[i] * a[i]); break;
        case 10: out[i] = orc_two_atan_mm(a[i], a[i] * a[i]); break;
        case 11: out[i] = orc_two_atanh_far(a[i], a[i] * a[i]); break;
        case 12: out[i] = orc_two_atan_far(a[i], a[i] * a[i]); break;
        case 13: out[i] = orc_two_atanh_mid(a[i], a[i] * a[i]); break;
        case 14: out[i] = orc_two_atanh_wide(a[i], a[i] * a[i]); break;
        default: out[i] = NAN;
        }
    }
}
