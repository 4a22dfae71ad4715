// toss_b200/csrc/gravity.cuh -- the polyhedral-gravity face-pair term and the TMA/mbarrier tile plumbing.
//
// Replaces polyhedral_gravity.GravityEvaluable.__call__ (call site toss/trajectory/equations_of_motion.py:58).
// Tsoulis (2012) line-integral sum, per face p with edges q:
//     S_p   = sum_q hq_s * LN_pq  +  hp * (sum_q sigma_pq AN_pq + singA_p)
//     g_pkg = G rho sum_p N_p S_p            (package sign; TOSS uses -g_pkg)
// with the two transcendental groups in their cancellation-free closed forms (same values; the literal
// forms lose ~(l1+l2)/|G| ulps to cancellation):
//     LN_pq = ln((s2+l2)/(s1+l1)) = 2 atanh(|G_pq| / (l1 + l2))
//     hp (sum_q sigma_pq AN_pq + singA_p) = -hp_s * omega_p,
//     omega_p = 2 atan2(r0.(r1 x r2), l0 l1 l2 + l0 (r1.r2) + l1 (r2.r0) + l2 (r0.r1))
// The unit of work is a PAIR of faces that share an edge (pairs.cu): LN of an edge depends only on its two end
// points, |v - P| only on the vertex, so a pair needs 4 square roots and 5 edge terms where two single faces need
// 6 and 6.  Work per pair-point: 4 sqrt, ONE division shared by the seven quotients, five 7-coefficient atanh
// polynomials, two 5-coefficient atan polynomials, ~110 other FP64 instructions -- everything on the FP64 pipe,
// nothing to feed tensor cores.  A term whose argument is outside its polynomial's range (an edge or a face within a
// few hundred metres of the field point) falls back ON ITS OWN to a separate division + ln-based atanh, or to the
// full-range atan2.  Every operation is spelled out (tmath.cuh): the value is a fixed sequence of IEEE operations.
#pragma once
#include "common.cuh"
#include "tmath.cuh"

// Two faces that share an edge, one field point (pair record of common.cuh): S_A and S_B, the scalars that
// multiply N_A and N_B.  Shared between the two: the distances to v0 and v1, r0.r1, l0*l1 and the edge term
// LN_0 of e0; on the common path ONE division serves the five edge quotients and the two solid-angle quotients.
// kMode says how the caller's warp is organised:
//   kPairLane     every lane has its own field point (thread-per-point kernel): no warp votes, 7 / 5 coefficients;
//   kPairLaneFar  the same with the far tier decided PER LANE (fine meshes: lanes rarely disagree);
//   kPairRow      the warp is converged on the call and its 32 lanes hold one ROW of 32 consecutive pair records for
//                 ONE field point: the row's tier (far / not) is decided by a vote;
//   kPairRowSmem  the same, plus `scratch` = 96 doubles of shared memory owned by the warp: the out-of-range edge
//                 terms of the row are compacted across the lanes (below).
enum : int { kPairLane = 0, kPairRow = 1, kPairRowSmem = 2, kPairLaneFar = 3 };

// Everything of a pair term that does not depend on the tier: distances, plane / edge heights, the seven quotients
// (ONE division), their squares and the high words the range tests look at.  (A caller that only needs some of the
// fields leaves the rest to dead-code elimination.)
struct PairFront {
    double hpA, hpB, hA0, hA1, hA2, hB0, hB1, hB2;
    double z0, z1, z2, z3, z4, z0s, z1s, z2s, z3s, z4s, qA, qB, qAs, qBs;
    double a0, a1, a2, a3, a4, numA, numB, denA, denB, G0, G1, G2, G3, G4;
    int hz0, hz1, hz2, hz3, hz4, hqA, hqB, hdA, hdB, hzmax, hqmax, hdmin;
};
template <class Fld>
__device__ __forceinline__ PairFront pair_front(Fld fld, double Px, double Py, double Pz)
{
    PairFront F;
    const double r0x = fld(0) - Px, r0y = fld(1) - Py, r0z = fld(2) - Pz;
    const double r1x = fld(3) - Px, r1y = fld(4) - Py, r1z = fld(5) - Pz;
    const double r2x = fld(6) - Px, r2y = fld(7) - Py, r2z = fld(8) - Pz;
    const double r3x = fld(9) - Px, r3y = fld(10) - Py, r3z = fld(11) - Pz;
    const double l0 = tm_sqrt(fma(r0z, r0z, fma(r0y, r0y, r0x * r0x)));
    const double l1 = tm_sqrt(fma(r1z, r1z, fma(r1y, r1y, r1x * r1x)));
    const double l2 = tm_sqrt(fma(r2z, r2z, fma(r2y, r2y, r2x * r2x)));
    const double l3 = tm_sqrt(fma(r3z, r3z, fma(r3y, r3y, r3x * r3x)));
    F.hpA = fma(fld(14), r0z, fma(fld(13), r0y, fld(12) * r0x));     // N_A . r0
    F.hpB = fma(fld(17), r1z, fma(fld(16), r1y, fld(15) * r1x));     // N_B . r1 (B's first vertex)
    F.hA0 = fma(fld(20), r0z, fma(fld(19), r0y, fld(18) * r0x));     // A: e0 at v0, e1 at v1, e2 at v2
    F.hA1 = fma(fld(23), r1z, fma(fld(22), r1y, fld(21) * r1x));
    F.hA2 = fma(fld(26), r2z, fma(fld(25), r2y, fld(24) * r2x));
    F.hB0 = fma(fld(29), r1z, fma(fld(28), r1y, fld(27) * r1x));     // B: e0 at v1, e3 at v0, e4 at v3
    F.hB1 = fma(fld(32), r0z, fma(fld(31), r0y, fld(30) * r0x));
    F.hB2 = fma(fld(35), r3z, fma(fld(34), r3y, fld(33) * r3x));
    F.a0 = l0 + l1, F.a1 = l1 + l2, F.a2 = l2 + l0, F.a3 = l0 + l3, F.a4 = l3 + l1;
    const double d01 = fma(r0z, r1z, fma(r0y, r1y, r0x * r1x));
    const double d12 = fma(r1z, r2z, fma(r1y, r2y, r1x * r2x));
    const double d20 = fma(r2z, r0z, fma(r2y, r0y, r2x * r0x));
    const double d03 = fma(r0z, r3z, fma(r0y, r3y, r0x * r3x));
    const double d31 = fma(r3z, r1z, fma(r3y, r1y, r3x * r1x));
    const double l01 = l0 * l1;
    F.denA = fma(l2, d01, fma(l1, d20, fma(l0, d12, l01 * l2)));
    F.denB = fma(l3, d01, fma(l0, d31, fma(l1, d03, l01 * l3)));
    F.numA = F.hpA * fld(41), F.numB = F.hpB * fld(42);
    F.G0 = fld(36), F.G1 = fld(37), F.G2 = fld(38), F.G3 = fld(39), F.G4 = fld(40);
    // One division serves the five edge quotients and the two solid-angle quotients (X Y is a product of lengths and
    // of the two solid-angle denominators: it only has to be non-zero; were a denominator exactly 0 every quotient
    // would come out non-finite and every term would take its own way below).
    const double P01 = F.a0 * F.a1, P2A = F.a2 * F.denA, P34 = F.a3 * F.a4;
    const double X = P01 * P2A, Y = P34 * F.denB;
    const double inv = tm_div(1.0, X * Y);
    const double iX = Y * inv, iY = X * inv;
    const double i01 = P2A * iX, i2A = P01 * iX, i34 = F.denB * iY, iB = P34 * iY;
    F.z0 = (F.G0 * F.a1) * i01, F.z1 = (F.G1 * F.a0) * i01, F.z2 = (F.G2 * F.denA) * i2A;
    F.qA = (F.numA * F.a2) * i2A;
    F.z3 = (F.G3 * F.a4) * i34, F.z4 = (F.G4 * F.a3) * i34;
    F.qB = F.numB * iB;
    F.z0s = F.z0 * F.z0, F.z1s = F.z1 * F.z1, F.z2s = F.z2 * F.z2, F.z3s = F.z3 * F.z3, F.z4s = F.z4 * F.z4;
    F.qAs = F.qA * F.qA, F.qBs = F.qB * F.qB;
    // Range tests.  An edge seen under a large angle (z^2 >= kZ2Max) takes a division of its own and the ln-based
    // atanh; a face seen under a large solid angle or from behind (den <= 0 or q^2 >= kQ2Max) the full-range atan2.
    // Each term chooses for itself; the rule is part of the arithmetic specification.  The
    // thresholds are doubles whose low word is 0, so `x < threshold` for x >= 0 is a comparison of the HIGH words
    // alone (a NaN pattern is above every finite one): the whole test costs two 3-input integer maxima, one
    // max, one min and three compares on the integer pipe instead of nine FP64-pipe compares -- the FP64 pipe (two
    // issue slots per instruction) is what bounds this loop.
    F.hz0 = __double2hiint(F.z0s), F.hz1 = __double2hiint(F.z1s), F.hz2 = __double2hiint(F.z2s),
    F.hz3 = __double2hiint(F.z3s), F.hz4 = __double2hiint(F.z4s);
    F.hqA = __double2hiint(F.qAs), F.hqB = __double2hiint(F.qBs), F.hdA = __double2hiint(F.denA),
    F.hdB = __double2hiint(F.denB);
    F.hzmax = max(max(max(F.hz0, F.hz1), F.hz2), max(F.hz3, F.hz4)), F.hqmax = max(F.hqA, F.hqB),
    F.hdmin = min(F.hdA, F.hdB);
    return F;
}

// FAR ROW: every z^2 and q^2 of the 32 pairs below 0x3F647AE1'00000000 (= 0.0025 - 5e-10) and every face seen
// from the front -- 84 % (init-like) to 97 % (converged-like) of the rows a population evaluates.  The row
// takes 4-coefficient polynomials (17 FP64 instructions less per pair); the tier is a property of (row, field
// point), so it does not depend on which warp evaluates the row.  Half values, as in the general tail.
__device__ __forceinline__ bool pair_lane_far(const PairFront &F) { return (max(F.hzmax, F.hqmax) < kHiFar) & (F.hdmin > 0); }
__device__ __forceinline__ void pair_tail_far(const PairFront &F, double &SA, double &SB)
{
    const double f0 = tm_atanh_far(F.z0, F.z0s), f1 = tm_atanh_far(F.z1, F.z1s), f2 = tm_atanh_far(F.z2, F.z2s),
                 f3 = tm_atanh_far(F.z3, F.z3s), f4 = tm_atanh_far(F.z4, F.z4s);
    const double wA = tm_atan_far(F.qA, F.qAs), wB = tm_atan_far(F.qB, F.qBs);
    SA = fma(-F.hpA, wA, fma(F.hA2, f2, fma(F.hA1, f1, F.hA0 * f0)));
    SB = fma(-F.hpB, wB, fma(F.hB2, f4, fma(F.hB1, f3, F.hB0 * f0)));
}
// MIDDLE ROW: every z^2 and q^2 below 0x3F847AE1'00000000 (= 0.01 - 2e-9), faces from the front -- another
// 12 % of the rows of an init-like population (whose trajectories start 5 km above the surface: hardly any
// of them is far all the time, and a tile ring advances at the pace of its slowest warp): 5-coefficient
// atanh, the 5-coefficient atan, no range tests.
__device__ __forceinline__ bool pair_lane_mid(const PairFront &F) { return (max(F.hzmax, F.hqmax) < kHiQ2Max) & (F.hdmin > 0); }
__device__ __forceinline__ void pair_tail_mid(const PairFront &F, double &SA, double &SB)
{
    const double f0 = tm_atanh_mid(F.z0, F.z0s), f1 = tm_atanh_mid(F.z1, F.z1s), f2 = tm_atanh_mid(F.z2, F.z2s),
                 f3 = tm_atanh_mid(F.z3, F.z3s), f4 = tm_atanh_mid(F.z4, F.z4s);
    const double wA = tm_atan_mm(F.qA, F.qAs), wB = tm_atan_mm(F.qB, F.qBs);
    SA = fma(-F.hpA, wA, fma(F.hA2, f2, fma(F.hA1, f1, F.hA0 * f0)));
    SB = fma(-F.hpB, wB, fma(F.hB2, f4, fma(F.hB1, f3, F.hB0 * f0)));
}

// the rows that are left (and every pair of the thread-per-point kernel): each term chooses its own path
template <int kMode>
__device__ __forceinline__ void pair_tail_general(const PairFront &F, double &SA, double &SB, double *scratch)
{
    constexpr bool kCompact = kMode == kPairRowSmem;
    constexpr bool kLane = kMode == kPairLane || kMode == kPairLaneFar;
    constexpr unsigned kFull = 0xffffffffu;
    const double z0 = F.z0, z1 = F.z1, z2 = F.z2, z3 = F.z3, z4 = F.z4;
    const double z0s = F.z0s, z1s = F.z1s, z2s = F.z2s, z3s = F.z3s, z4s = F.z4s;
    const double qA = F.qA, qB = F.qB, qAs = F.qAs, qBs = F.qBs;
    const double G0 = F.G0, G1 = F.G1, G2 = F.G2, G3 = F.G3, G4 = F.G4;
    const double a0 = F.a0, a1 = F.a1, a2 = F.a2, a3 = F.a3, a4 = F.a4;
    const double numA = F.numA, numB = F.numB, denA = F.denA, denB = F.denB;
    const int hz0 = F.hz0, hz1 = F.hz1, hz2 = F.hz2, hz3 = F.hz3, hz4 = F.hz4, hqA = F.hqA, hqB = F.hqB, hdA = F.hdA,
              hdB = F.hdB, hzmax = F.hzmax, hqmax = F.hqmax, hdmin = F.hdmin;
    // the lane-per-pair kernels take an 11-coefficient atanh up to z^2 < 0.16 (a third as many
    // rows with out-of-range edge terms as with the 7-coefficient one up to 0.04; the extra 20 FP64 instructions fall
    // on the 5 % of the rows that get here); the thread-per-point kernel keeps 7 coefficients up to 0.04
    constexpr int kZ2Max = kLane ? kHiZ2Max : kHiZ2Wide;
    const bool anybad = (hzmax >= kZ2Max) | (hqmax >= kHiQ2Max) | (hdmin <= 0);
    // the warp vote is issued BEFORE the polynomials so that its latency (integer maxima -> compare -> vote) hides
    // behind them instead of sitting at the end of the row with nothing left to issue
    const bool warp_bad = kCompact ? (__any_sync(kFull, anybad) != 0) : false;
    // HALF values from here on: atanh(z) and atan(q) instead of 2 atanh(z) = LN and 2 atan(q) = omega.  Scaling by
    // 2 is exact, so S/2 and the accumulated sum/2 carry the same bits; the callers multiply by 2 G rho at the end.
    double ln0, ln1, ln2, ln3, ln4;
    if (kLane) {
        ln0 = tm_atanh_mm(z0, z0s), ln1 = tm_atanh_mm(z1, z1s), ln2 = tm_atanh_mm(z2, z2s), ln3 = tm_atanh_mm(z3, z3s),
        ln4 = tm_atanh_mm(z4, z4s);
    } else {
        ln0 = tm_atanh_wide(z0, z0s), ln1 = tm_atanh_wide(z1, z1s), ln2 = tm_atanh_wide(z2, z2s),
        ln3 = tm_atanh_wide(z3, z3s), ln4 = tm_atanh_wide(z4, z4s);
    }
    double omA = tm_atan_mm(qA, qAs);
    double omB = tm_atan_mm(qB, qBs);
    unsigned badw = 0u;
    if (kCompact) {
        // Out-of-range terms are rare (a pair within a few hundred metres of the field point; ~1.4 % of the 32-pair
        // rows of an init-like population have any, then ~13 edge terms spread over ~5 lanes).  In the stepper every
        // instruction spent here is paid by all the warps that share the tile ring, so the warp COMPACTS the bad
        // edge terms of a row: the owners post (|G|, l1 + l2) to the warp's scratch, lane s evaluates term s, the
        // owners read the values back -- ceil(#terms / 32) general evaluations per row instead of
        // max-over-lanes(#terms of a lane).  Same operations on the same operands: the values do not depend on who
        // computes them.
        if (warp_bad) {
            const unsigned bad = (hz0 < kZ2Max ? 0u : 1u) | (hz1 < kZ2Max ? 0u : 2u) | (hz2 < kZ2Max ? 0u : 4u) |
                                 (hz3 < kZ2Max ? 0u : 8u) | (hz4 < kZ2Max ? 0u : 16u);
            badw = ((hdA > 0 && hqA < kHiQ2Max) ? 0u : 1u) | ((hdB > 0 && hqB < kHiQ2Max) ? 0u : 2u);
            const unsigned lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
            const unsigned b0 = __ballot_sync(kFull, bad & 1u), b1 = __ballot_sync(kFull, bad & 2u),
                           b2 = __ballot_sync(kFull, bad & 4u), b3 = __ballot_sync(kFull, bad & 8u),
                           b4 = __ballot_sync(kFull, bad & 16u);
            const unsigned c1 = __popc(b0), c2 = c1 + __popc(b1), c3 = c2 + __popc(b2), c4 = c3 + __popc(b3),
                           total = c4 + __popc(b4);
            const unsigned p0 = __popc(b0 & lt), p1 = c1 + __popc(b1 & lt), p2 = c2 + __popc(b2 & lt),
                           p3 = c3 + __popc(b3 & lt), p4 = c4 + __popc(b4 & lt);
            double *sG = scratch, *sa = scratch + 32, *sv = scratch + 64;
            for (unsigned base = 0; base < total; base += 32u) {
                if ((bad & 1u) && p0 - base < 32u) { sG[p0 - base] = G0; sa[p0 - base] = a0; }
                if ((bad & 2u) && p1 - base < 32u) { sG[p1 - base] = G1; sa[p1 - base] = a1; }
                if ((bad & 4u) && p2 - base < 32u) { sG[p2 - base] = G2; sa[p2 - base] = a2; }
                if ((bad & 8u) && p3 - base < 32u) { sG[p3 - base] = G3; sa[p3 - base] = a3; }
                if ((bad & 16u) && p4 - base < 32u) { sG[p4 - base] = G4; sa[p4 - base] = a4; }
                __syncwarp();
                if (base + lane < total) sv[lane] = 0.5 * tm_two_atanh(tm_div(sG[lane], sa[lane]));
                __syncwarp();
                if ((bad & 1u) && p0 - base < 32u) ln0 = sv[p0 - base];
                if ((bad & 2u) && p1 - base < 32u) ln1 = sv[p1 - base];
                if ((bad & 4u) && p2 - base < 32u) ln2 = sv[p2 - base];
                if ((bad & 8u) && p3 - base < 32u) ln3 = sv[p3 - base];
                if ((bad & 16u) && p4 - base < 32u) ln4 = sv[p4 - base];
                __syncwarp();
            }
        }
    } else if (anybad) {
        // callers without a converged warp (thread-per-point kernel: 32 different field points): every lane fixes
        // its own terms
        if (hz0 >= kZ2Max) ln0 = 0.5 * tm_two_atanh(tm_div(G0, a0));
        if (hz1 >= kZ2Max) ln1 = 0.5 * tm_two_atanh(tm_div(G1, a1));
        if (hz2 >= kZ2Max) ln2 = 0.5 * tm_two_atanh(tm_div(G2, a2));
        if (hz3 >= kZ2Max) ln3 = 0.5 * tm_two_atanh(tm_div(G3, a3));
        if (hz4 >= kZ2Max) ln4 = 0.5 * tm_two_atanh(tm_div(G4, a4));
        badw = ((hdA > 0 && hqA < kHiQ2Max) ? 0u : 1u) | ((hdB > 0 && hqB < kHiQ2Max) ? 0u : 2u);
    }
    while (badw) {   // solid-angle terms out of range: 0.4 per row that has any bad term -- not worth compacting
        const bool isA = badw & 1u;
        badw &= badw - 1;
        const double v = tm_atan2(isA ? numA : numB, isA ? denA : denB);
        if (isA) omA = v;
        else omB = v;
    }
    SA = fma(-F.hpA, omA, fma(F.hA2, ln2, fma(F.hA1, ln1, F.hA0 * ln0)));
    SB = fma(-F.hpB, omB, fma(F.hB2, ln4, fma(F.hB1, ln3, F.hB0 * ln0)));
}

template <int kMode = kPairLane, class Fld>
__device__ __forceinline__ void pair_S(Fld fld, double Px, double Py, double Pz, double &SA, double &SB,
                                       double &hpA_out, double &hpB_out, double *scratch = nullptr)
{
    constexpr unsigned kFull = 0xffffffffu;
    const PairFront F = pair_front(fld, Px, Py, Pz);
    hpA_out = F.hpA;
    hpB_out = F.hpB;
    if (kMode == kPairLaneFar) {
        // thread-per-point kernel on a fine mesh: the far tier is the LANE's -- a function of (pair, point) alone, so a
        // point's value does not depend on its neighbours in the batch.  Lanes that disagree diverge and the warp runs
        // both tails: that pays when faces are small against the distances (full mesh, 1 M shell points: 136 -> 129 ms)
        // and costs when they are not (lp: 13.9 -> 16.5 ms), so the launcher picks by the mesh's longest edge.
        if (pair_lane_far(F)) {
            pair_tail_far(F, SA, SB);
            return;
        }
    } else if (kMode == kPairLane) {
    } else {
        if (__all_sync(kFull, pair_lane_far(F))) {
            pair_tail_far(F, SA, SB);
            return;
        }
        if (__all_sync(kFull, pair_lane_mid(F))) {
            pair_tail_mid(F, SA, SB);
            return;
        }
    }
    pair_tail_general<kMode>(F, SA, SB, scratch);
}

// acc += (N_A S_A + N_B S_B) / 2 (and pot += (hp_A S_A + hp_B S_B) / 2): pair_S works in exact halves
template <bool kPot, int kMode = kPairLane, class Fld>
__device__ __forceinline__ void pair_term(Fld fld, double Px, double Py, double Pz, double &ax, double &ay,
                                          double &az, double &pot, double *scratch = nullptr)
{
    double SA, SB, hpA, hpB;
    pair_S<kMode>(fld, Px, Py, Pz, SA, SB, hpA, hpB, scratch);
    ax = fma(fld(15), SB, fma(fld(12), SA, ax));
    ay = fma(fld(16), SB, fma(fld(13), SA, ay));
    az = fma(fld(17), SB, fma(fld(14), SA, az));
    if (kPot) pot = fma(hpB, SB, fma(hpA, SA, pot));
}

// mesh_utility.py:132-166: Moeller-Trumbore with ray direction (0,0,1) (:83).  Every rounding is the
// one numpy performs (un-fused), so the hit parity is bit-exact against the reference code.
__device__ __forceinline__ int ray_hits_dev(double ox, double oy, double oz, double v0x, double v0y, double v0z,
                                            double v1x, double v1y, double v1z, double v2x, double v2y, double v2z)
{
    const double e1x = sub_(v1x, v0x), e1y = sub_(v1y, v0y), e1z = sub_(v1z, v0z);
    const double e2x = sub_(v2x, v0x), e2y = sub_(v2y, v0y), e2z = sub_(v2z, v0z);
    const double hx = -e2y, hy = e2x;                       // cross((0,0,1), e2) = (-e2y, e2x, 0)
    const double a = add_(mul_(e1x, hx), mul_(e1y, hy));    // + e1z*0
    if (a < 0.000001 && a > -0.000001) return 0;
    const double f = __ddiv_rn(1.0, a);
    const double sx = sub_(ox, v0x), sy = sub_(oy, v0y), sz = sub_(oz, v0z);
    const double u = mul_(add_(mul_(sx, hx), mul_(sy, hy)), f);
    if (u < 0.0 || u > 1.0) return 0;
    const double q0 = sub_(mul_(sy, e1z), mul_(sz, e1y));
    const double q1 = sub_(mul_(sz, e1x), mul_(sx, e1z));
    const double q2 = sub_(mul_(sx, e1y), mul_(sy, e1x));
    const double v = mul_(q2, f);                           // dot(q, (0,0,1)) * f
    if (v < 0.0 || add_(u, v) > 1.0) return 0;
    const double t = mul_(f, add_(add_(mul_(q0, e2x), mul_(q1, e2y)), mul_(q2, e2z)));
    return t > 0.0 ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// TMA (cp.async.bulk, SASS UBLKCP) + mbarrier helpers.  Tiles are contiguous in global memory, so the
// 1-D bulk copy is all that is needed -- no tensor map.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_drop(uint64_t *bar)
{
    asm volatile("mbarrier.arrive_drop.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
#ifdef K2_WAIT_HINT
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
#else
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
#endif
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
#ifdef K2_WAIT_HINT
        , "r"((uint32_t)K2_WAIT_HINT)
#endif
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
