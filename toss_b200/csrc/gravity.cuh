// toss_b200/csrc/gravity.cuh -- the polyhedral-gravity face term and the TMA/mbarrier tile plumbing.
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
// Work per face-point pair on the common path (face seen under a small solid angle, edges under small
// angles): 3 sqrt, ONE division shared by the four quotients, three 11-term atanh series, one 8-term atan
// series, ~60 other FP64 instructions -- everything on the FP64 pipe, nothing to feed tensor cores.  Faces
// close to the field point take the general path (separate divisions, ln-based atanh, full-range atan2).
// Every operation is spelled out (tmath.cuh): the value is a fixed sequence of IEEE operations.
#pragma once
#include "common.cuh"
#include "tmath.cuh"

// One face, one field point P.  fld(k) returns field k of the face record (common.cuh layout); the
// accessor decides where it comes from (shared-memory AoS broadcast, shared-memory SoA, global SoA).
// S_p of one face (the scalar that multiplies N_p); hp_out = signed plane distance (potential term).
template <class Fld>
__device__ __forceinline__ double face_S(Fld fld, double Px, double Py, double Pz, double &hp_out)
{
    const double r0x = fld(0) - Px, r0y = fld(1) - Py, r0z = fld(2) - Pz;
    const double r1x = fld(3) - Px, r1y = fld(4) - Py, r1z = fld(5) - Pz;
    const double r2x = fld(6) - Px, r2y = fld(7) - Py, r2z = fld(8) - Pz;
    const double l0 = tm_sqrt(fma(r0z, r0z, fma(r0y, r0y, r0x * r0x)));
    const double l1 = tm_sqrt(fma(r1z, r1z, fma(r1y, r1y, r1x * r1x)));
    const double l2 = tm_sqrt(fma(r2z, r2z, fma(r2y, r2y, r2x * r2x)));
    const double Nx = fld(9), Ny = fld(10), Nz = fld(11);
    const double hp = fma(Nz, r0z, fma(Ny, r0y, Nx * r0x));                  // signed plane distance
    const double hq0 = fma(fld(14), r0z, fma(fld(13), r0y, fld(12) * r0x));  // signed edge-line distances
    const double hq1 = fma(fld(17), r1z, fma(fld(16), r1y, fld(15) * r1x));
    const double hq2 = fma(fld(20), r2z, fma(fld(19), r2y, fld(18) * r2x));
    const double A0 = l0 + l1, A1 = l1 + l2, A2 = l2 + l0;
    const double d12 = fma(r1z, r2z, fma(r1y, r2y, r1x * r2x));
    const double d20 = fma(r2z, r0z, fma(r2y, r0y, r2x * r0x));
    const double d01 = fma(r0z, r1z, fma(r0y, r1y, r0x * r1x));
    const double den = fma(l2, d01, fma(l1, d20, fma(l0, d12, (l0 * l1) * l2)));
    const double num = hp * fld(24);
    const double G0 = fld(21), G1 = fld(22), G2 = fld(23);
    // (measured alternatives, both slower: a shorter far-field series tier -- near- and far-side faces share a
    //  warp, so both tiers execute -- and computing the common path unconditionally -- register pressure)
    double ln0, ln1, ln2, om;
    bool fast = (den > 0.0) && (fabs(num) <= 0.1 * den);
    if (fast) {
        const double P01 = A0 * A1, P2D = A2 * den;
        const double inv = tm_div(1.0, P01 * P2D);
        const double i01 = P2D * inv, i2D = P01 * inv;
        const double z0 = (G0 * A1) * i01, z1 = (G1 * A0) * i01, z2 = (G2 * den) * i2D;
        const double q = (num * A2) * i2D;
        const double z0s = z0 * z0, z1s = z1 * z1, z2s = z2 * z2;
        fast = (z0s < 0.04) && (z1s < 0.04) && (z2s < 0.04);
        if (fast) {
            ln0 = tm_two_atanh_series_n<11>(z0, z0s);
            ln1 = tm_two_atanh_series_n<11>(z1, z1s);
            ln2 = tm_two_atanh_series_n<11>(z2, z2s);
            om = 2.0 * tm_atan_series_n<8>(q, q * q);
        }
    }
    if (!fast) {   // face close to the field point: separate divisions, ln-based atanh, full-range atan2
        ln0 = tm_two_atanh(tm_div(G0, A0));
        ln1 = tm_two_atanh(tm_div(G1, A1));
        ln2 = tm_two_atanh(tm_div(G2, A2));
        om = 2.0 * tm_atan2(num, den);
    }
    hp_out = hp;
    return fma(-hp, om, fma(hq2, ln2, fma(hq1, ln1, hq0 * ln0)));
}

// One face, one field point: acc += N_p S_p (and pot += hp S_p).
template <bool kPot, class Fld>
__device__ __forceinline__ void face_term(Fld fld, double Px, double Py, double Pz, double &ax, double &ay,
                                          double &az, double &pot)
{
    double hp;
    const double S = face_S(fld, Px, Py, Pz, hp);
    ax = fma(fld(9), S, ax);
    ay = fma(fld(10), S, ay);
    az = fma(fld(11), S, az);
    if (kPot) pot = fma(hp, S, pot);
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
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
