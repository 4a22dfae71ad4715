// toss_b200/csrc/stepper.cu -- K2: population-parallel RK8(7)-13M (the reference's "DOP853" slot:
// desolver RK8713MSolver, toss/trajectory/Integrator.py:14) through the polyhedral gravity field.
//
// Replaces compute_trajectory -> integrate_system -> desolver OdeSystem.integrate
// (toss/trajectory/compute_trajectory.py:13-162, 203-261) for a whole population.
//
// One warp per trajectory.  Lanes 0..5 own the six state components (y_c, k_j[c]); for every RHS
// evaluation the 32 lanes split the face pairs (lane -> pair records lane, lane+32, ...), accumulate in a fixed
// order and butterfly-reduce, so a trajectory's arithmetic does not depend on the population around
// it.  Warps pull trajectories from an atomic queue and retire them independently (finished, step
// cap, non-finite state, knot capacity).
//
// Two mesh-staging modes:
//   RESIDENT  the flat SoA pair table fits in shared memory (lllp, llp): TMA-copied once per CTA;
//             warps then run completely independently.
//   STREAM    (lp, full) the table streams through a 3-stage ring of 45 KB slot tiles (128 pairs, 22 double2 slots each) filled by TMA
//             bulk copies (mbarrier full/empty pairs); the 16 warps of a CTA consume the same tile
//             sequence, one RHS evaluation per warp per sweep, coupled only through the ring (a warp can
//             run at most two tiles ahead), so every byte fetched from L2 is used by 16 trajectories.
//             There is no CTA-wide barrier after start-up: warps retire from the ring with arrive_drop.
#include <stdlib.h>
#include <string.h>

#include "gravity.cuh"

#define TOSS_RK_QUAL __constant__
#include "rk8713m_tableau.h"

#ifndef K2_WARPS
#define K2_WARPS 16
#endif
#ifndef K2_STAGES
#define K2_STAGES 3
#endif
#ifndef K2_MINBLOCKS
#define K2_MINBLOCKS 1
#endif
#ifndef K2_UNROLL
#define K2_UNROLL 1
#endif
constexpr int kK2Warps = K2_WARPS;
constexpr int kK2WarpsFew = 12;            // CTA size for few trajectories per warp (divisible by the team sizes 2 and 4)
constexpr int kK2Threads = kK2Warps * 32;
constexpr int kK2Stages = K2_STAGES;
constexpr int kK2Unroll = K2_UNROLL;
constexpr int kK2BarSlots = 2 * K2_STAGES + 2 + (K2_STAGES + 2) / 2;   // full[] | empty[] | live + claim[] (ints) | pad
constexpr int kPairTileBytes = kPairTileDoubles * 8;

struct StepArgs {
    toss_params p;
    UnitAxis axis;
    double Grho;
    const double *pair_soa;    // [22][Ppad32] double2 slots (RESIDENT)
    const double *pair_tiles;  // [n_tiles][22][128] double2 (STREAM)
    const double *ray_soa;     // [9][Fpad32] faces in mesh order (global memory): the event-point ray test
    int Ppad32, n_tiles, last_tile_pairs;   // last_tile_pairs: real pairs of the last tile rounded up to 32
    int Fpad32;
    const double *x;           // pop x dim
    int pop, dim, n_fixed, n_man, knot_cap;
    double fixed[8];
    // workspace
    double *knots_t, *knots_y, *knots_f, *intervals;
    int32_t *seg_start, *n_seg, *collision, *queue;
    const int32_t *order;      // queue position -> trajectory index (longest predicted first), or NULL
    int first_round, first_extra;   // static first round: every CTA takes first_round queue positions, the first
                                    // first_extra CTAs one more (see the fetch)
    int interleave;            // static round: position slot * CTAs + b instead of a contiguous block per CTA
    int proxy_n;               // PROXY instantiation: number of mascons (pair_soa then points at [n][4] x y z G*m)
    double proxy_soft2;
    int32_t *proxy_cost;       // PROXY: predicted RHS count per trajectory
    toss_counters *counters;
};

// per-warp shared state (doubles): y[6] | k[13][6] | meta[2] | seg_t[n_man+2] | dv[3*n_man]
__host__ __device__ inline int warp_state_doubles(int n_man) { return 6 + 78 + 2 + (n_man + 2) + 3 * (n_man > 0 ? n_man : 1); }

// Called by lane 0 of a warp right after it arrived on empty[s] for tile `it`: if that arrival completed the phase
// (every warp still in the ring has released the tile), claim the refill and issue the TMA load of tile it + stages.
__device__ __forceinline__ void refill_if_released(uint64_t *bars, int *claim, double *mesh, const double *pair_tiles,
                                                   int n_tiles, int s, uint32_t it)
{
    uint32_t done;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(smem_u32(&bars[kK2Stages + s])), "r"((it / kK2Stages) & 1u)
        : "memory");
    if (done && atomicMax(&claim[s], (int)it) < (int)it) {
        mbar_arrive_expect_tx(&bars[s], kPairTileBytes);
        tma_load_1d(mesh + (size_t)s * kPairTileDoubles,
                    pair_tiles + (size_t)((it + kK2Stages) % (uint32_t)n_tiles) * kPairTileDoubles, kPairTileBytes, &bars[s]);
    }
}

// Teams (kTeam = 2 or 4 warps per trajectory) are for populations too small to fill the machine with one warp
// each (or to keep its tail short).  The owner warp (rank 0) runs the state machine; every warp of the team computes
// the per-face scalars S_A, S_B for its share of the 32-pair rows of a tile (rows rank, rank + kTeam, ...) into a
// shared buffer; after a team barrier the owner adds N_A S_A + N_B S_B in the canonical order (row 0, 1, 2, ... of every tile).  The
// arithmetic, and therefore every result bit, is the same for any team size.
enum : int { kModeGrav = 0, kModeSkip = 1, kModeRetire = 2 };
#ifdef K2_NO_COMPACT
constexpr int kCompactRows = kPairRow;       // developer A/B: per-lane loops over the out-of-range terms
#else
constexpr int kCompactRows = kPairRowSmem;
#endif
constexpr int kScratchDoubles = 96;             // pair_S scratch per warp: |G|[32] | l1+l2[32] | value[32]
#ifdef K2_NO_TEAMBUF
constexpr int kTeamBufs = 0;               // developer experiment (team size 1 only): no team mailboxes
#else
constexpr int kTeamBufs = kK2Warps / 2;    // one mailbox per team of two (the most teams a CTA can hold)
#endif
constexpr int kTeamBufDoubles = 4 + 2 * 4 * 64;   // P[3] + mode | S[2 buffers][4 rows][S_A x 32 lanes | S_B x 32 lanes]

__device__ __forceinline__ void team_bar(int team, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(nthreads) : "memory");
}

// kProxy: the same state machine integrating through a mascon proxy of the body (no mesh, no knots, no events); its
// RHS count per trajectory is the cost key of the longest-first work queue of the real run.
// kW: warps per CTA.  16 warps x 128 registers (ptxas spills ~330 bytes) or 12 warps x ~160 registers (no spills):
// the same throughput within 2.5 % on a population that queues deep, and a warp that runs 10 % faster when there are
// few trajectories per warp (small populations, shards of a strong-scaled one) -- the launcher picks.
template <bool kStream, int kTeam, bool kProxy = false, int kW = kK2Warps>
__global__ void __launch_bounds__(kW * 32, K2_MINBLOCKS) stepper_kernel(const StepArgs a)
{
    constexpr int kTeamBufsW = kTeamBufs ? kW / 2 : 0;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int team = wid / kTeam, rank = wid % kTeam;
    const int mesh_doubles = kProxy ? 4 * a.proxy_n + kProxyTail : (kStream ? kK2Stages * kPairTileDoubles : 2 * kPairSlots * a.Ppad32);
    double *mesh = reinterpret_cast<double *>(smem_raw);
    uint64_t *bars = reinterpret_cast<uint64_t *>(mesh + mesh_doubles);   // full[3] empty[3] (STREAM) / full[1]
    double *wstate = reinterpret_cast<double *>(bars + kK2BarSlots) + (size_t)wid * warp_state_doubles(a.n_man);
    double *sy = wstate, *sk = wstate + 6, *meta = wstate + 84, *seg_t = wstate + 86, *sdv = seg_t + (a.n_man + 2);
    double *tbuf = reinterpret_cast<double *>(bars + kK2BarSlots) + (size_t)kW * warp_state_doubles(a.n_man) +
                   (size_t)team * kTeamBufDoubles;                       // team mailbox: P, mode, S buffers
    // per-warp scratch of pair_S (compaction of the out-of-range edge terms of a row): 96 doubles
    double *scr = reinterpret_cast<double *>(bars + kK2BarSlots) + (size_t)kW * warp_state_doubles(a.n_man) +
                  (size_t)kTeamBufsW * kTeamBufDoubles + (size_t)wid * kScratchDoubles;
    volatile double *tP = tbuf;
    volatile int *tmode = reinterpret_cast<volatile int *>(tbuf + 3);
    double *sbuf = tbuf + 4;
    int spar = 0;                                                        // which S buffer the current tile uses

    int *live = reinterpret_cast<int *>(bars + 2 * kK2Stages);   // warps that have not retired yet (STREAM)
    if (tid == 0) {
        *live = kW;
        for (int s = 0; s < kK2Stages; ++s) live[1 + s] = -1;
        if (kStream) {
            for (int s = 0; s < kK2Stages; ++s) {
                mbar_init(&bars[s], 1);
                mbar_init(&bars[kK2Stages + s], kW);
            }
        } else {
            mbar_init(&bars[0], 1);
        }
        mbar_fence_init();
    }
    __syncthreads();
    int *claim = live + 1;    // claim[s] = newest tile of stage s whose refill has been issued (STREAM)
    if (tid == 0) {
        if (kStream) {
            for (int s = 0; s < kK2Stages; ++s) {   // the ring is cyclic over the mesh: always something to fetch
                mbar_arrive_expect_tx(&bars[s], kPairTileBytes);
                tma_load_1d(mesh + (size_t)s * kPairTileDoubles,
                            a.pair_tiles + (size_t)(s % a.n_tiles) * kPairTileDoubles, kPairTileBytes, &bars[s]);
            }
        } else {
            const uint32_t bytes = (uint32_t)mesh_doubles * 8u;
            mbar_arrive_expect_tx(&bars[0], bytes);
            uint32_t off = 0;   // bulk copies are capped well below 1 MB each; chunk to be safe
            while (off < bytes) {
                const uint32_t c = bytes - off < 65536u ? bytes - off : 65536u;
                tma_load_1d(reinterpret_cast<unsigned char *>(mesh) + off,
                            reinterpret_cast<const unsigned char *>(a.pair_soa) + off, c, &bars[0]);
                off += c;
            }
        }
    }
    if (!kStream) mbar_wait(&bars[0], 0);

    const toss_params &p = a.p;
    if (kTeam > 1 && rank != 0) {
        // ------------------------------------------------ helper warp: an S_p server for its team's owner
        uint32_t it = 0;
        for (;;) {
            team_bar(team, 32 * kTeam);   // the owner has published the evaluation point and the mode
            const int mode = *tmode;
            const double Px = tP[0], Py = tP[1], Pz = tP[2];
            if (mode == kModeRetire) {
                if (kStream) {
                    for (int s = 0; s < kK2Stages; ++s) {
                        if (it > (uint32_t)s) {
                            const uint32_t last = it - 1 - ((it - 1 - s) % kK2Stages);
                            mbar_wait(&bars[kK2Stages + s], (last / kK2Stages) & 1u);
                        }
                        __syncwarp();
                        if (lane == 0) {
                            const uint32_t nxt = it + (uint32_t)((s + kK2Stages - (int)(it % kK2Stages)) % kK2Stages);
                            mbar_arrive_drop(&bars[kK2Stages + s]);
                            refill_if_released(bars, claim, mesh, a.pair_tiles, a.n_tiles, s, nxt);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) atomicSub(live, 1);
                }
                return;
            }
            if (kStream) {
                for (int tl = 0; tl < a.n_tiles; ++tl, ++it) {
                    const int s = it % kK2Stages;
                    const int rows = ((tl == a.n_tiles - 1) ? a.last_tile_pairs : kPairTile) >> 5;
                    mbar_wait(&bars[s], (it / kK2Stages) & 1u);
                    if (mode == kModeGrav) {
                        const double2 *tile = reinterpret_cast<const double2 *>(mesh + (size_t)s * kPairTileDoubles) + lane;
                        for (int r = rank; r < rows; r += kTeam) {
                            const PairCols<false> cols{tile + 32 * r, kPairTile};
                            double SA, SB, hpA, hpB;
                            pair_S<kCompactRows>(cols, Px, Py, Pz, SA, SB, hpA, hpB, scr);
                            sbuf[(spar * 4 + r) * 64 + lane] = SA;
                            sbuf[(spar * 4 + r) * 64 + 32 + lane] = SB;
                        }
                        team_bar(team, 32 * kTeam);   // S of this tile complete; the owner accumulates
                        spar ^= 1;
                    }
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&bars[kK2Stages + s]);
                        refill_if_released(bars, claim, mesh, a.pair_tiles, a.n_tiles, s, it);
                    }
                }
            } else if (mode == kModeGrav) {
                const int Fp = a.Ppad32, rows_total = Fp >> 5;
                for (int v = 0; v * 4 < rows_total; ++v) {
                    const int rows = rows_total - 4 * v < 4 ? rows_total - 4 * v : 4;
                    for (int r = rank; r < rows; r += kTeam) {
                        const PairCols<false> cols{reinterpret_cast<const double2 *>(mesh) + 32 * (4 * v + r) + lane, Fp};
                        double SA, SB, hpA, hpB;
                        pair_S<kCompactRows>(cols, Px, Py, Pz, SA, SB, hpA, hpB, scr);
                        sbuf[(spar * 4 + r) * 64 + lane] = SA;
                        sbuf[(spar * 4 + r) * 64 + 32 + lane] = SB;
                    }
                    team_bar(team, 32 * kTeam);
                    spar ^= 1;
                }
            }
        }
    }
    bool team_alive = kTeam > 1;   // helpers still listening (owner side)
    const double R2 = mul_(p.radius_inner, p.radius_inner);
    uint32_t tile_it = 0;     // tiles consumed so far by this warp (== by every warp: lock step)

    // warp-uniform trajectory state
    bool active = false, queue_empty = false, counted_out = false, first_fetch = true;
    int traj = -1, stage = 0, seg = 0, nseg = 0, nk = 0, steps = 0, clipped = 0, after_accept = 0, status = 0;
    int n_acc = 0, n_rej = 0, n_ev = 0, collided = 0, stopped = 0, ray_sweep = 0;
    long long n_rhs = 0;
    double t = 0.0, dt = 0.0, tf = 0.0;

#ifdef K2_PROFILE
    long long prof_wait = 0, prof_comp = 0, prof_wait0 = 0, prof_idle = 0;
    const long long prof_t0 = clock64();
    // per-section clocks of the owner warp's loop: 0 advance / fetch, 1 stage point, 2 shuffles + rotation,
    // 3 team publish + ray sweep, 4 gravity sweep, 5 reduction + k store  (sections end at the marks below)
    long long prof_sec[6] = {0, 0, 0, 0, 0, 0}, prof_last = prof_t0;
#define K2_MARK(i) { const long long c_ = clock64(); prof_sec[i] += c_ - prof_last; prof_last = c_; }
#else
#define K2_MARK(i)
#endif
    for (;;) {
        // ------------------------------------------------ fetch a trajectory, decode the chromosome
        if (!active && !queue_empty) {
            // The first round is assigned statically: CTA b takes k (or k + 1) consecutive queue positions, k =
            // floor(pop / CTAs) capped by the teams of a CTA.  A population smaller than the machine spreads over ALL
            // SMs (a warp that shares its scheduler with fewer warps runs faster) instead of filling the first CTAs;
            // a large one hands every CTA a contiguous block of the longest-first order, so trajectories of similar
            // length -- the long ones are the ones that pass close to the body and take the slow rows -- share a tile
            // ring.  Later trajectories come from the atomic cursor, which starts behind the static round.
            int nxt = 0;
            const int b = (int)blockIdx.x;
            // Few trajectories per warp (a shard of a strong-scaled population, a.interleave): the static positions
            // are interleaved (slot * CTAs + b), so that every SM gets the same mix of lengths and none is left with
            // sixteen of the longest trajectories -- contiguous blocks cost 10 % at 1.7 trajectories per warp.
            const int mine = a.first_round + (b < a.first_extra ? 1 : 0);   // static positions of this CTA
            if (first_fetch && team < mine) {
                nxt = a.interleave ? team * (int)gridDim.x + b
                                   : b * a.first_round + (b < a.first_extra ? b : a.first_extra) + team;
            } else {
                if (lane == 0) nxt = atomicAdd(a.queue, 1) + (int)gridDim.x * a.first_round + a.first_extra;
                nxt = __shfl_sync(0xffffffffu, nxt, 0);
            }
            first_fetch = false;
            if (nxt >= a.pop) {
                queue_empty = true;
            } else {
                traj = a.order ? a.order[nxt] : nxt;
                active = true;
                // x = hstack(initial_condition, chromosome)  (udp_initial_condition.py:117-120)
                auto X = [&](int i) -> double {
                    return i < a.n_fixed ? a.fixed[i] : a.x[(size_t)traj * a.dim + (i - a.n_fixed)];
                };
                // compute_trajectory.py:53-55: v0 = v_mag * v_dir (direction NOT normalised)
                if (lane < 3) sy[lane] = X(lane);
                else if (lane < 6) sy[lane] = mul_(X(3), X(lane + 1));
                // compute_trajectory.py:75-121: int32 truncation, stable sort, equal times merged (dv summed)
                if (lane == 0) {
                    const int M = a.n_man;
                    // sort scratch in the warp's shared-memory scratch (free between sweeps): no local-memory arrays
                    int *idx = reinterpret_cast<int *>(scr);
                    int32_t *tm = reinterpret_cast<int32_t *>(scr) + 64;
                    for (int k = 0; k < M; ++k) {
                        tm[k] = (int32_t)X(7 + 5 * k);
                        idx[k] = k;
                    }
                    for (int i = 1; i < M; ++i) {
                        int j = i, v = idx[i];
                        while (j > 0 && tm[idx[j - 1]] > tm[v]) {
                            idx[j] = idx[j - 1];
                            --j;
                        }
                        idx[j] = v;
                    }
                    int n = 0;
                    seg_t[0] = (double)(int32_t)p.start_time;
                    for (int i = 0; i < M; ++i) {
                        const int k = idx[i];
                        const double mag = X(8 + 5 * k);
                        const double dx = mul_(mag, X(9 + 5 * k)), dy = mul_(mag, X(10 + 5 * k)),
                                     dz = mul_(mag, X(11 + 5 * k));
                        if (n > 0 && (double)tm[k] == seg_t[n]) {
                            sdv[3 * (n - 1)] = add_(sdv[3 * (n - 1)], dx);
                            sdv[3 * (n - 1) + 1] = add_(sdv[3 * (n - 1) + 1], dy);
                            sdv[3 * (n - 1) + 2] = add_(sdv[3 * (n - 1) + 2], dz);
                        } else {
                            seg_t[n + 1] = (double)tm[k];
                            sdv[3 * n] = dx;
                            sdv[3 * n + 1] = dy;
                            sdv[3 * n + 2] = dz;
                            ++n;
                        }
                    }
                    seg_t[n + 1] = (double)(int32_t)p.final_time;
                    if (!kProxy) {   // (the proxy run has no workspace)
                        a.n_seg[traj] = n + 1;
                        for (int s = 0; s <= n + 1; ++s) a.intervals[(size_t)traj * (a.n_man + 2) + s] = seg_t[s];
                    }
                    meta[0] = (double)(n + 1);   // segment count travels to the other lanes through shared memory
                }
                __syncwarp();
                nseg = (int)meta[0];
                __syncwarp();
                seg = 0;
                t = seg_t[0];
                tf = seg_t[1];
                dt = p.initial_time_step;
                stage = 0;
                nk = 0;
                steps = 0;
                clipped = 0;
                after_accept = 0;
                status = 0;
                n_acc = n_rej = n_ev = collided = stopped = ray_sweep = 0;
                n_rhs = 0;
                if (!kProxy && lane == 0) a.seg_start[(size_t)traj * (a.n_man + 2)] = 0;
            }
        }
        if (!active) {
            // queue drained, nothing in hand: this warp retires.  RESIDENT: just leave.  STREAM: leave the
            // tile ring cleanly -- wait until every tile this warp released has been released by all, then
            // arrive_drop on each `empty` barrier (counts as this warp's arrival for the tile it will not
            // read and removes it from all later phases).  Warp 0 owns the producer thread: it stays as an
            // idle consumer/producer until the other warps are gone.
            if (team_alive) {   // dissolve the team: the helpers leave (and drop out of the tile ring) on their own
                team_alive = false;
                if (lane == 0) *tmode = kModeRetire;
                __syncwarp();
                team_bar(team, 32 * kTeam);
            }
            if (!kStream) break;
            if (wid != 0) {
                for (int s = 0; s < kK2Stages; ++s) {
                    // last tile of stage s this warp has consumed: largest it < tile_it with it % kK2Stages == s
                    if (tile_it > (uint32_t)s) {
                        const uint32_t last = tile_it - 1 - ((tile_it - 1 - s) % kK2Stages);
                        mbar_wait(&bars[kK2Stages + s], (last / kK2Stages) & 1u);
                    }
                    __syncwarp();
                    if (lane == 0) {
                        // the phase this arrival belongs to: the first tile >= tile_it that lives in stage s
                        const uint32_t nxt = tile_it + (uint32_t)((s + kK2Stages - (int)(tile_it % kK2Stages)) % kK2Stages);
                        mbar_arrive_drop(&bars[kK2Stages + s]);
                        refill_if_released(bars, claim, mesh, a.pair_tiles, a.n_tiles, s, nxt);
                    }
                }
                __syncwarp();
                if (lane == 0) atomicSub(live, 1);
                break;
            }
            if (!counted_out) {
                counted_out = true;
                if (lane == 0) atomicSub(live, 1);
                __syncwarp();
            }
            if (*reinterpret_cast<volatile int *>(live) == 0) break;
        }

        // ------------------------------------------------ stage point: ys = y + dt sum_j a_ij k_j
        K2_MARK(0)
        double ysc = 0.0, te = t, rot_s = 0.0, rot_c = 1.0;
        if (active) {
            // One straight-line block: (1) the stage time and sin / cos of the body's rotation angle at it, which depend
            // on (t, dt, stage) only, and (2) s = sum_{j < stage} a_ij k_j in increasing j (the order is part of the
            // arithmetic specification).  The two dependency chains are independent, so the scheduler interleaves them;
            // all twelve coefficient / slope loads are issued up front and the terms beyond `stage` are predicated
            // off -- one load latency per stage point instead of one per remainder block of a loop with a run-time
            // trip count.  Lanes 6..31 mirror lane 5 instead of branching around the block (their value is not used).
            // This chain is on the critical path of a small population, where the latency of an RHS evaluation counts.
            const int l6 = lane < 6 ? lane : 5;
            te = stage > 0 ? add_(t, mul_(TOSS_RK_C[stage], dt)) : t;
            tm_sincos(-mul_(p.spin_velocity, te), rot_s, rot_c);
            double pr_[12];
#pragma unroll
            for (int j = 0; j < 12; ++j) pr_[j] = mul_(TOSS_RK_A[13 * stage + j], sk[6 * j + l6]);
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < 12; ++j)
                if (j < stage) s = add_(s, pr_[j]);
            const double y0 = sy[l6];
            ysc = stage > 0 ? add_(y0, mul_(dt, s)) : y0;
        }
        K2_MARK(1)
        const double qx = __shfl_sync(0xffffffffu, ysc, 0), qy = __shfl_sync(0xffffffffu, ysc, 1),
                     qz = __shfl_sync(0xffffffffu, ysc, 2);
        const double vsh = __shfl_sync(0xffffffffu, ysc, (lane + 3) & 31);   // lanes 0..2 receive ys[3..5]
        double Px = qx, Py = qy, Pz = qz;
        if (active && p.activate_rotation)
            rotate_point_sc(rot_s, rot_c, qx, qy, qz, a.axis.x, a.axis.y, a.axis.z, Px, Py, Pz);

        K2_MARK(2)
        if (team_alive) {   // tell the helpers what this sweep is
            if (lane == 0) {
                tP[0] = Px;
                tP[1] = Py;
                tP[2] = Pz;
                *tmode = ray_sweep ? kModeSkip : kModeGrav;
            }
            __syncwarp();
            team_bar(team, 32 * kTeam);
        }
        // A sweep is either a gravity evaluation or (ray_sweep) the point-in-polyhedron test of an event point:
        // compute_trajectory.py:264-282 records every ACCEPTED state inside the risk sphere, is_outside
        // (mesh_utility.py:70-89) ray-casts it.  Event points are rare (a few per close pass), so the test gets a
        // sweep of its own over the same tile stream instead of loading the hot loop's registers.
        int hits = 0;
        if (!kProxy && active && ray_sweep) {
            const int Fr = a.Fpad32;
#pragma unroll 2
            for (int f = lane; f < Fr; f += 32) {
                const double *col = a.ray_soa + f;
                hits += ray_hits_dev(qx, qy, qz, __ldg(col), __ldg(col + Fr), __ldg(col + 2 * Fr), __ldg(col + 3 * Fr),
                                     __ldg(col + 4 * Fr), __ldg(col + 5 * Fr), __ldg(col + 6 * Fr), __ldg(col + 7 * Fr),
                                     __ldg(col + 8 * Fr));
            }
        }

        // ------------------------------------------------ gravity sweep over all face pairs
        K2_MARK(3)
        double ax = 0.0, ay = 0.0, az = 0.0, pv = 0.0;
        if (kProxy) {
            for (int m = lane; m < a.proxy_n; m += 32) {
                const double dx = Px - mesh[4 * m], dy = Py - mesh[4 * m + 1], dz = Pz - mesh[4 * m + 2];
                const double r2 = fma(dz, dz, fma(dy, dy, dx * dx)) + a.proxy_soft2;
                const double g = tm_div(mesh[4 * m + 3], r2 * tm_sqrt(r2));
                ax = fma(-g, dx, ax);
                ay = fma(-g, dy, ay);
                az = fma(-g, dz, az);
            }
        } else if (kStream) {
            for (int tl = 0; tl < a.n_tiles; ++tl, ++tile_it) {
                const int s = tile_it % kK2Stages;
                // rows of 32 pairs in this tile that hold real pairs (the last tile is only partly filled)
                const int f_end = (tl == a.n_tiles - 1) ? a.last_tile_pairs : kPairTile;
#ifdef K2_PROFILE
                const long long c0 = clock64();
#endif
                mbar_wait(&bars[s], (tile_it / kK2Stages) & 1u);
#ifdef K2_PROFILE
                const long long c1 = clock64();
                prof_wait += c1 - c0;
                if (tl == 0) prof_wait0 += c1 - c0;
                if (!active) prof_idle += c1 - c0;
#endif
                if (active && ray_sweep) {
                    // the ray test ran on the global face table above; this sweep only keeps step with the ring
                } else if (active && kTeam > 1) {
                    // team sweep: S of my rows into the buffer, barrier, then N S of ALL rows in canonical order
                    const double2 *tile = reinterpret_cast<const double2 *>(mesh + (size_t)s * kPairTileDoubles) + lane;
                    const int rows = f_end >> 5;
                    for (int r = 0; r < rows; r += kTeam) {
                        const PairCols<false> cols{tile + 32 * r, kPairTile};
                        double SA, SB, hpA, hpB;
                        pair_S<kCompactRows>(cols, Px, Py, Pz, SA, SB, hpA, hpB, scr);
                        sbuf[(spar * 4 + r) * 64 + lane] = SA;
                        sbuf[(spar * 4 + r) * 64 + 32 + lane] = SB;
                    }
                    team_bar(team, 32 * kTeam);
                    for (int r = 0; r < rows; ++r) {
                        const PairCols<false> cols{tile + 32 * r, kPairTile};
                        const double SA = sbuf[(spar * 4 + r) * 64 + lane], SB = sbuf[(spar * 4 + r) * 64 + 32 + lane];
                        ax = fma(cols(15), SB, fma(cols(12), SA, ax));
                        ay = fma(cols(16), SB, fma(cols(13), SA, ay));
                        az = fma(cols(17), SB, fma(cols(14), SA, az));
                    }
                    spar ^= 1;
                } else if (active) {
                    // (a running row pointer and a down-counter: the loop carries no lane id, no tile length and no
                    //  kernel parameter -- ptxas otherwise re-reads all three for every row to save registers, ~80
                    //  cycles that a warp alone on its scheduler cannot hide)
                    const double2 *rp = reinterpret_cast<const double2 *>(mesh + (size_t)s * kPairTileDoubles) + lane;
#pragma unroll kK2Unroll
                    for (int r = f_end >> 5; r > 0; --r, rp += 32) {
                        const PairCols<false> cols{rp, kPairTile};
                        pair_term<false, kCompactRows>(cols, Px, Py, Pz, ax, ay, az, pv, scr);
                    }
                }
                __syncwarp();
#ifdef K2_PROFILE
                if (active) prof_comp += clock64() - c1;
#endif
                // release the stage; the warp whose arrival completes the phase (the last one out) refills it at
                // once -- no dedicated producer, no warp ever waits on another one's bookkeeping
                if (lane == 0) {
                    mbar_arrive(&bars[kK2Stages + s]);
                    refill_if_released(bars, claim, mesh, a.pair_tiles, a.n_tiles, s, tile_it);
                }
            }
        } else {
            const int Fp = a.Ppad32;
            if (ray_sweep) {
                // done above on the global face table
            } else if (kTeam > 1) {
                const int rows_total = Fp >> 5;
                for (int v = 0; v * 4 < rows_total; ++v) {
                    const int rows = rows_total - 4 * v < 4 ? rows_total - 4 * v : 4;
                    for (int r = 0; r < rows; r += kTeam) {
                        const PairCols<false> cols{reinterpret_cast<const double2 *>(mesh) + 32 * (4 * v + r) + lane, Fp};
                        double SA, SB, hpA, hpB;
                        pair_S<kCompactRows>(cols, Px, Py, Pz, SA, SB, hpA, hpB, scr);
                        sbuf[(spar * 4 + r) * 64 + lane] = SA;
                        sbuf[(spar * 4 + r) * 64 + 32 + lane] = SB;
                    }
                    team_bar(team, 32 * kTeam);
                    for (int r = 0; r < rows; ++r) {
                        const PairCols<false> cols{reinterpret_cast<const double2 *>(mesh) + 32 * (4 * v + r) + lane, Fp};
                        const double SA = sbuf[(spar * 4 + r) * 64 + lane], SB = sbuf[(spar * 4 + r) * 64 + 32 + lane];
                        ax = fma(cols(15), SB, fma(cols(12), SA, ax));
                        ay = fma(cols(16), SB, fma(cols(13), SA, ay));
                        az = fma(cols(17), SB, fma(cols(14), SA, az));
                    }
                    spar ^= 1;
                }
            } else {
                const double2 *rp = reinterpret_cast<const double2 *>(mesh) + lane;   // (as in the streaming loop)
#pragma unroll kK2Unroll
                for (int r = Fp >> 5; r > 0; --r, rp += 32) {
                    const PairCols<false> cols{rp, Fp};
                    pair_term<false, kCompactRows>(cols, Px, Py, Pz, ax, ay, az, pv, scr);
                }
            }
        }
        K2_MARK(4)
        if (!active) continue;
        bool post_event = false;   // run the tail of the stage-0 logic (after the event test, if there was one)
        if (ray_sweep) {
            ray_sweep = 0;
            ++n_ev;
            if (__reduce_add_sync(0xffffffffu, hits) & 1) collided = 1;   // odd number of +z hits: inside the mesh
            post_event = true;
        } else {
        ax = warp_sum(ax);
        ay = warp_sum(ay);
        az = warp_sum(az);
        ++n_rhs;
        // k_stage = [ys[3..5], -g_pkg]   (equations_of_motion.py:59,92-95)
        // (selects, not branches: lanes 3, 4 and 5 would otherwise serialise three copies of the same two instructions)
        const double asel = lane == 3 ? ax : (lane == 4 ? ay : az);
        const double gsel = kProxy ? asel : -((a.Grho + a.Grho) * asel);   // the pair sum runs in exact halves
        const double kc = lane < 3 ? vsh : gsel;
        if (lane < 6) sk[6 * stage + lane] = kc;
        __syncwarp();
        if (stage == 0) {
            // f0 at a segment start or at a freshly accepted state: this is a knot
            if (kProxy) {
                // no knots for the proxy run.  An accepted state inside the body (occupancy grid appended to the mascon
                // table) is where the real run finds its first event point inside the mesh and -- with
                // stop_on_collision -- retires: the proxy does the same, so that a collider's predicted cost is short
                if (after_accept && p.stop_on_collision && p.activate_event) {
                    const double *occ = mesh + 4 * a.proxy_n;
                    const double gx = (qx - occ[0]) * occ[3], gy = (qy - occ[1]) * occ[4], gz = (qz - occ[2]) * occ[5];
                    if (gx >= 0.0 && gy >= 0.0 && gz >= 0.0 && gx < (double)kProxyOcc && gy < (double)kProxyOcc &&
                        gz < (double)kProxyOcc) {
                        const int c = ((int)gx * kProxyOcc + (int)gy) * kProxyOcc + (int)gz;
                        const unsigned long long wbits = reinterpret_cast<const unsigned long long *>(occ + 6)[c >> 6];
                        if ((wbits >> (c & 63)) & 1ull) collided = 1;
                    }
                }
            } else if (nk >= a.knot_cap) {
                status = 3;
            } else {
                const size_t kb = (size_t)traj * a.knot_cap + nk;
                if (lane == 0) a.knots_t[kb] = t;
                if (lane < 6) {
                    a.knots_y[kb * 6 + lane] = sy[lane];
                    a.knots_f[kb * 6 + lane] = kc;
                }
                ++nk;
            }
            if (!kProxy && status == 0 && after_accept && p.activate_event &&
                add_(add_(mul_(qx, qx), mul_(qy, qy)), mul_(qz, qz)) <= R2) {
                ray_sweep = 1;   // event point: next sweep is its ray cast, then the logic below resumes
                continue;
            }
            post_event = true;
        }
        }

        // ------------------------------------------------ advance the state machine
        K2_MARK(5)
        if (post_event) {
            if (after_accept) {
                after_accept = 0;
                // a collided trajectory scores the constant 1e30 (udp_initial_condition.py:126-128): with
                // stop_on_collision it retires here; the reference behaviour (0) integrates on to final_time
                if (collided && p.stop_on_collision) {
                    stopped = 1;
                } else {
                    ++steps;
                    if (steps > p.max_steps) status = 1;
                }
            }
            if (stopped) {
                // fall through to retirement
            } else if (status == 0 && !(t < tf)) {   // segment finished
                ++seg;
                if (seg >= nseg) {
                    active = false;
                } else {
                    // compute_trajectory.py:134-136: velocity REPLACED by the manoeuvre's dv
                    if (lane >= 3 && lane < 6) sy[lane] = sdv[3 * (seg - 1) + (lane - 3)];
                    __syncwarp();
                    if (!kProxy && lane == 0) a.seg_start[(size_t)traj * (a.n_man + 2) + seg] = nk;
                    t = seg_t[seg];
                    tf = seg_t[seg + 1];
                    dt = p.initial_time_step;
                    steps = 0;
                    stage = 0;   // f0 of the new segment
                    continue;
                }
            } else if (status == 0) {
                clipped = 0;
                if (add_(t, dt) > tf) {
                    dt = sub_(tf, t);
                    clipped = 1;
                }
                stage = 1;
                continue;
            }
        } else if (stage < 12) {
            ++stage;
            continue;
        } else {
            // all 13 stages in: error estimate and desolver's controller:
            // err = max |h sum (b8-b7) k|, tol = max(atol + rtol |y| + rtol |dy/h|), h' = 0.8 h (tol/err)^(1/8)
            double e = 0.0, tl = 0.0, dy = 0.0;
            if (lane < 6) {
                double s8 = 0.0, se = 0.0;
                for (int j = 0; j < 13; ++j) {
                    const double kj = sk[6 * j + lane];
                    s8 = add_(s8, mul_(TOSS_RK_B8[j], kj));
                    se = add_(se, mul_(TOSS_RK_E[j], kj));
                }
                dy = mul_(dt, s8);
                e = fabs(mul_(dt, se));
                tl = add_(add_(p.atol, mul_(p.rtol, fabs(sy[lane]))), mul_(p.rtol, fabs(s8)));
                if (e != e) e = __longlong_as_double(0x7ff0000000000000LL);   // NaN must win the max
                if (tl != tl) tl = __longlong_as_double(0x7ff0000000000000LL);
            }
            const double err = warp_max(e), tol = warp_max(tl);
            if (isinf(err) || isinf(tol)) {
                status = 2;
            } else {
                double corr = 1.0;
                if (err != 0.0) corr = mul_(0.8, tm_root8(__ddiv_rn(tol, err)));
                const double newdt = (corr != 0.0) ? mul_(corr, dt) : dt;
                if (err > tol) {
                    ++n_rej;
                    dt = newdt;
                    clipped = 0;
                    ++steps;
                    if (steps > p.max_steps || !(dt > 0.0)) status = 1;
                    else {
                        stage = 1;
                        continue;
                    }
                } else {
                    t = clipped ? tf : add_(t, dt);
                    if (lane < 6) sy[lane] = add_(sy[lane], dy);
                    __syncwarp();
                    dt = newdt;
                    ++n_acc;
                    after_accept = 1;
                    stage = 0;
                    continue;
                }
            }
        }
        // ------------------------------------------------ retire (finished or failed)
        active = false;
        if (kProxy) {
            if (lane == 0) a.proxy_cost[traj] = (int32_t)(n_rhs > 2000000000LL ? 2000000000LL : n_rhs);
        } else if (lane == 0) {
            a.seg_start[(size_t)traj * (a.n_man + 2) + (seg < nseg ? seg + 1 : nseg)] = nk;
            if (status != 0 || stopped) a.n_seg[traj] = seg < nseg ? seg + 1 : nseg;
            toss_counters c;
            c.n_rhs = n_rhs;
            c.n_accepted = n_acc;
            c.n_rejected = n_rej;
            c.n_event_points = n_ev;
            c.status = status;
            a.counters[traj] = c;
            a.collision[traj] = collided;
        }
    }

#ifdef K2_PROFILE
    if (lane == 0) {
        unsigned long long *pr = reinterpret_cast<unsigned long long *>(a.queue) + 2;   // queue[4..] (LPT area, unused here)
        atomicAdd(pr + 0, (unsigned long long)prof_wait);
        atomicAdd(pr + 1, (unsigned long long)prof_comp);
        atomicAdd(pr + 2, (unsigned long long)(clock64() - prof_t0));
        atomicAdd(pr + 3, 1ULL);
        atomicAdd(pr + 4, (unsigned long long)prof_wait0);
        atomicAdd(pr + 5, (unsigned long long)prof_idle);
        atomicAdd(pr + 8 + (wid & 7), (unsigned long long)prof_wait);
        atomicAdd(pr + 16 + (wid & 7), (unsigned long long)prof_comp);
        if (kTeam == 1 || rank == 0)
            for (int i = 0; i < 6; ++i) atomicAdd(pr + 24 + i, (unsigned long long)prof_sec[i]);
    }
#endif
    // drain the TMA copies still in flight before the CTA's shared memory goes away
    if (kStream && tid == 0) {
        // (every tile below tile_it + kK2Stages has been issued: the stage of tile i is refilled with tile i + stages
        //  by whoever releases it last, and this warp has released all tiles below tile_it)
        for (uint32_t it = tile_it; it < tile_it + kK2Stages; ++it)
            mbar_wait(&bars[it % kK2Stages], (it / kK2Stages) & 1u);
    }
}


// ------------------------------------------------------------------------------------------------
// Work-queue ordering.  Trajectory lengths spread 4x (1.3 k .. 5 k RHS evaluations) and a small population has only a
// few trajectories per resident warp, so the order in which warps pull them decides the tail.  The PROXY instantiation
// of the stepper pre-integrates every chromosome through a mascon model of the body (softened point masses on the
// cells of an 8^3 grid that lie inside the mesh, rotated like the real field; same controller, same tolerances: its RHS
// count equals the real one for most trajectories, r = 0.96 overall) at ~5 % of the real cost; the counts are the sort
// key (longest first).  The order only schedules: results are written by trajectory index and do not depend on it.
// ------------------------------------------------------------------------------------------------
// One block: STABLE counting sort of the predicted costs, longest first, ties in index order -- a pure function of the
// cost array, so every rank of a multi-GPU job derives the same order from the same (all-gathered) costs and the
// dealt shards are consistent (parallel.py).  Chunks of 1024 elements in index order; inside a chunk the warps take
// their turns in warp order, inside a warp equal bins are ranked by lane (match_any): no atomics on the output.
__global__ void __launch_bounds__(1024) cost_order_kernel(const int32_t *__restrict__ cost, int n,
                                                          int32_t *__restrict__ order)
{
    __shared__ int hist[4096];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int b = threadIdx.x; b < 4096; b += blockDim.x) hist[b] = 0;
    __syncthreads();
    auto key = [](int c) { c >>= 2; return c < 0 ? 0 : (c > 4095 ? 4095 : c); };
    for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&hist[key(cost[i])], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int b = 4095; b >= 0; --b) {
            const int h = hist[b];
            hist[b] = run;
            run += h;
        }
    }
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + (int)threadIdx.x;
        const bool valid = i < n;
        const int b = valid ? key(cost[i]) : -1 - lane;            // invalid lanes match nobody
        const unsigned peers = __match_any_sync(0xffffffffu, b);
        const int rank_in_warp = __popc(peers & ((1u << lane) - 1u)), cnt = __popc(peers);
        const bool leader = rank_in_warp == 0;
        int off = 0;
        for (int w = 0; w < 32; ++w) {
            if (wid == w && valid && leader) {
                off = hist[b];
                hist[b] = off + cnt;
            }
            __syncthreads();
        }
        off = __shfl_sync(0xffffffffu, off, __ffs(peers) - 1);
        if (valid) order[off + rank_in_warp] = i;
    }
}

int toss_workspace_layout_impl(int pop, int n_man, int knot_cap, toss_ws_layout *out)
{
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t o = 0;
    out->knots_t = o;   o = up(o + sizeof(double) * (size_t)pop * knot_cap);
    out->knots_y = o;   o = up(o + sizeof(double) * (size_t)pop * knot_cap * 6);
    out->knots_f = o;   o = up(o + sizeof(double) * (size_t)pop * knot_cap * 6);
    out->seg_start = o; o = up(o + sizeof(int32_t) * (size_t)pop * (n_man + 2));
    out->intervals = o; o = up(o + sizeof(double) * (size_t)pop * (n_man + 2));
    out->n_seg = o;     o = up(o + sizeof(int32_t) * (size_t)pop);
    out->counters = o;  o = up(o + sizeof(toss_counters) * (size_t)pop);
    out->collision = o; o = up(o + sizeof(int32_t) * (size_t)pop);
    out->queue = o;     o = up(o + sizeof(int32_t) * (4 + 2 * (size_t)pop));   // cursor | order[pop] | cost[pop]
    out->total_bytes = o;
    return 0;
}

static size_t stepper_state_bytes(int n_man, int warps = kK2Warps)
{
    return (size_t)warps * warp_state_doubles(n_man) * 8 + kK2BarSlots * 8 +
           (size_t)(kTeamBufs ? warps / 2 : 0) * kTeamBufDoubles * 8 + (size_t)warps * kScratchDoubles * 8;
}

// the part of StepArgs that does not depend on a workspace
static int fill_step_args(StepArgs &a, toss_ctx *ctx, const toss_params *p, const double *d_x, int pop, int dim,
                          const double *h_fixed, int n_fixed, int n_man, const char *who)
{
    TOSS_REQUIRE(ctx && ctx->d_pair_soa, std::string(who) + ": no mesh uploaded");
    TOSS_REQUIRE(pop > 0 && dim >= 0 && n_fixed >= 0 && n_fixed <= 7, std::string(who) + ": bad population shape");
    TOSS_REQUIRE(n_man >= 0 && n_man <= 64, std::string(who) + ": number_of_maneuvers must be in [0, 64]");
    TOSS_REQUIRE(n_fixed + dim == 7 + 5 * n_man, std::string(who) + ": n_fixed + dim != 7 + 5*number_of_maneuvers");
    TOSS_REQUIRE((int32_t)p->final_time > (int32_t)p->start_time, std::string(who) + ": final_time <= start_time");
    memset(&a, 0, sizeof(a));
    a.p = *p;
    a.axis = unit_axis_host(p->spin_axis);
    a.Grho = ctx->Grho;
    a.pair_soa = ctx->d_pair_soa;
    a.pair_tiles = ctx->d_pair_tiles;
    a.ray_soa = ctx->d_ray_soa;
    a.Ppad32 = ctx->Ppad32;
    a.Fpad32 = ctx->Fpad32;
    a.n_tiles = ctx->n_pair_tiles;
    a.last_tile_pairs = ctx->Ppad32 - (ctx->n_pair_tiles - 1) * kPairTile;
    a.x = d_x;
    a.pop = pop;
    a.dim = dim;
    a.n_fixed = n_fixed;
    a.n_man = n_man;
    for (int i = 0; i < 8; ++i) a.fixed[i] = (i < n_fixed && h_fixed) ? h_fixed[i] : 0.0;
    return 0;
}

// The PROXY run: predicted RHS count of every trajectory of `a` (a.queue must point at a zeroed cursor).  It only has
// to RANK the trajectories by length: it runs at a loose tolerance (the step count of an 8th-order method scales
// like tol^(-1/8) for every trajectory alike: 7.5x fewer steps at 1e-5 than at 1e-12).
static int launch_proxy(toss_ctx *ctx, const StepArgs &a, int32_t *d_cost, cudaStream_t st)
{
    StepArgs px = a;
    px.order = nullptr;
    px.pair_soa = ctx->d_mascons;
    px.proxy_n = ctx->n_mascons;
    px.proxy_soft2 = ctx->mascon_soft2;
    px.proxy_cost = d_cost;
    double ptol = 1e-5;
    if (const char *e = getenv("TOSS_B200_PROXY_TOL")) ptol = atof(e) > 0.0 ? atof(e) : ptol;   // developer override
    if (px.p.rtol < ptol) px.p.rtol = ptol;
    if (px.p.atol < ptol) px.p.atol = ptol;
    const size_t psmem = ((size_t)4 * ctx->n_mascons + kProxyTail) * 8 + stepper_state_bytes(a.n_man);
    px.interleave = 0;
    const int pctas = a.pop < K2_MINBLOCKS * ctx->sm_count ? a.pop : K2_MINBLOCKS * ctx->sm_count;
    px.first_round = a.pop / pctas < kK2Warps ? a.pop / pctas : kK2Warps;
    px.first_extra = a.pop / pctas < kK2Warps ? a.pop % pctas : 0;
    TOSS_CHECK_CUDA(cudaFuncSetAttribute(stepper_kernel<false, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)psmem));
    stepper_kernel<false, 1, true><<<pctas, kK2Threads, psmem, st>>>(px);
    ctx->launches++;
    TOSS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// whether the proxy run pays: the mesh must be large enough that a mascon evaluation is cheap next to a real one
static bool proxy_pays(const toss_ctx *ctx) { return ctx->F >= 1000 && ctx->n_mascons > 0; }

// toss_predict_cost: the cost key of every chromosome (0 everywhere when the proxy would not pay: any order is as
// good as another then).  d_scratch_queue: >= 256 bytes of device memory for the proxy's work-queue cursor.
int launch_predict_cost(toss_ctx *ctx, const toss_params *p, const double *d_x, int pop, int dim, const double *h_fixed,
                        int n_fixed, int n_man, int32_t *d_cost, cudaStream_t st)
{
    StepArgs a;
    if (int rc = fill_step_args(a, ctx, p, d_x, pop, dim, h_fixed, n_fixed, n_man, "toss_predict_cost")) return rc;
    if (!proxy_pays(ctx)) {
        TOSS_CHECK_CUDA(cudaMemsetAsync(d_cost, 0, sizeof(int32_t) * (size_t)pop, st));
        return 0;
    }
    if (ctx->queue_scratch == nullptr) TOSS_CHECK_CUDA(cudaMalloc(&ctx->queue_scratch, 256));
    a.queue = ctx->queue_scratch;
    TOSS_CHECK_CUDA(cudaMemsetAsync(a.queue, 0, 256, st));
    NvtxRange nv("toss:proxy run (cost prediction)");
    return launch_proxy(ctx, a, d_cost, st);
}

int launch_cost_order(toss_ctx *ctx, const int32_t *d_cost, int n, int32_t *d_order, cudaStream_t st)
{
    TOSS_REQUIRE(d_cost && d_order && n > 0, "toss_cost_order: bad argument");
    cost_order_kernel<<<1, 1024, 0, st>>>(d_cost, n, d_order);
    ctx->launches++;
    TOSS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_propagate(toss_ctx *ctx, const toss_params *p, const double *d_x, int pop, int dim, const double *h_fixed,
                     int n_fixed, int n_man, void *d_ws, int knot_cap, cudaStream_t st)
{
    StepArgs a;
    if (int rc = fill_step_args(a, ctx, p, d_x, pop, dim, h_fixed, n_fixed, n_man, "toss_propagate")) return rc;
    TOSS_REQUIRE(knot_cap >= 4, "toss_propagate: knot capacity too small");
    toss_ws_layout L;
    toss_workspace_layout_impl(pop, n_man, knot_cap, &L);
    unsigned char *ws = static_cast<unsigned char *>(d_ws);
    a.knot_cap = knot_cap;
    a.knots_t = reinterpret_cast<double *>(ws + L.knots_t);
    a.knots_y = reinterpret_cast<double *>(ws + L.knots_y);
    a.knots_f = reinterpret_cast<double *>(ws + L.knots_f);
    a.intervals = reinterpret_cast<double *>(ws + L.intervals);
    a.seg_start = reinterpret_cast<int32_t *>(ws + L.seg_start);
    a.n_seg = reinterpret_cast<int32_t *>(ws + L.n_seg);
    a.collision = reinterpret_cast<int32_t *>(ws + L.collision);
    a.queue = reinterpret_cast<int32_t *>(ws + L.queue);
    a.counters = reinterpret_cast<toss_counters *>(ws + L.counters);
    TOSS_CHECK_CUDA(cudaMemsetAsync(a.queue, 0, 256, st));
    // warps per trajectory: one when every warp of the machine gets a trajectory; two or four below that, which fills
    // the SMs of small populations and shortens their longest trajectory.  Results do not depend on it.
    // warps per CTA (see the kernel): 16 when trajectories queue deep (>= 5 per warp), else 12
    int W = (long long)pop >= 5LL * K2_MINBLOCKS * ctx->sm_count * kK2Warps ? kK2Warps : kK2WarpsFew;
    if (const char *e = getenv("TOSS_B200_WARPS")) W = atoi(e) == kK2WarpsFew ? kK2WarpsFew : kK2Warps;   // developer override
    const int slots = K2_MINBLOCKS * ctx->sm_count * W;
    // Measured (lp / full mesh, pop 1024 .. 8192, longest-first queue): the best team size keeps the machine loaded
    // with one to two waves of teams -- 4 warps up to slots/2 trajectories, 2 up to slots, 1 beyond (a team of two
    // runs a trajectory 1.86x faster, not 2x, so once every warp has a trajectory of its own teams only cost).
    int team = 4 * (long long)pop <= 2LL * slots ? 4 : (2 * (long long)pop <= 2LL * slots ? 2 : 1);
    // a team splits the 32-pair ROWS of the table: more warps than rows only adds barrier hand-shakes (lllp: 1 row)
    const int rows = ctx->Ppad32 / 32;
    if (rows < 2) team = 1;
    else if (rows < 3 && team > 2) team = 2;
    if (const char *e = getenv("TOSS_B200_TEAM")) team = atoi(e) == 4 ? 4 : (atoi(e) == 2 ? 2 : 1);   // developer override
    // longest-first queue order when trajectories queue for warps (otherwise order is irrelevant) and the mesh is
    // large enough that the proxy run is cheap next to the real one (lp pop 32768: 1292 -> 1220 ms).  A caller that
    // has ordered the rows itself (the multi-GPU deal of parallel.py hands every rank a longest-first shard) says so
    // with toss_set_queue_order(ctx, 1): the queue then runs in row order.
    bool lpt = ctx->queue_order_mode == 0 && (long long)pop * team > slots && proxy_pays(ctx);
    if (const char *e = getenv("TOSS_B200_LPT")) lpt = atoi(e) != 0 && ctx->n_mascons > 0 && ctx->queue_order_mode == 0;
    a.order = nullptr;
    if (lpt) {
        NvtxRange nv("toss:proxy run + queue sort (longest first)");
        int32_t *order = a.queue + 4, *cost = order + pop;
        if (int rc = launch_proxy(ctx, a, cost, st)) return rc;
        if (int rc = launch_cost_order(ctx, cost, pop, order, st)) return rc;
        TOSS_CHECK_CUDA(cudaMemsetAsync(a.queue, 0, 16, st));
        a.order = order;
    }

    NvtxRange nv_k2("toss:K2 stepper");
    const size_t state_bytes = stepper_state_bytes(n_man, W);
    const size_t resident_bytes = (size_t)2 * kPairSlots * ctx->Ppad32 * 8 + state_bytes;
    const size_t stream_bytes = (size_t)kK2Stages * kPairTileBytes + state_bytes;
    const bool resident = resident_bytes <= (size_t)(200 / K2_MINBLOCKS) * 1024;   // every CTA on an SM keeps its own copy
    // one CTA per SM as soon as there is a trajectory for it (the static first round spreads them over the CTAs)
    const int max_ctas = K2_MINBLOCKS * ctx->sm_count;
    const int ctas = pop < max_ctas ? pop : max_ctas;
    a.first_round = pop / ctas < W / team ? pop / ctas : W / team;
    a.first_extra = pop / ctas < W / team ? pop % ctas : 0;
    if (const char *e = getenv("TOSS_B200_FIRST_ROUND"))   // developer A/B: 1 = one static position per CTA
        if (atoi(e) == 1) a.first_round = 1, a.first_extra = 0;
    // Few trajectories per warp and a longest-first queue (own proxy run, or rows pre-ordered by the multi-GPU deal):
    // interleaved static round (see the fetch).  lp, one 4096-row shard of the 32768 population: 175 -> 159 ms.
    // Measured and dropped: the head of the queue on CTAs that run only 4 / 8 / 12 warps (two per scheduler: 22
    // instead of 35 us per evaluation for the longest trajectories) -- 157 .. 160 ms for 4 .. 16 such CTAs, no gain:
    // at 1.7 trajectories per warp the tail is the bulk of two-trajectory warps, not the few longest ones.
    const bool ordered = lpt || ctx->queue_order_mode == 1;
    a.interleave = (ordered && (long long)pop * team < 4LL * slots) ? 1 : 0;
    if (const char *e = getenv("TOSS_B200_INTERLEAVE")) a.interleave = atoi(e) != 0;   // developer A/B
    const size_t smem = resident ? resident_bytes : stream_bytes;
#define TOSS_LAUNCH_STEPPER_W(STREAM, TEAM, WARPS)                                                                  \
    do {                                                                                                           \
        TOSS_CHECK_CUDA(cudaFuncSetAttribute(stepper_kernel<STREAM, TEAM, false, WARPS>,                           \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
        stepper_kernel<STREAM, TEAM, false, WARPS><<<ctas, WARPS * 32, smem, st>>>(a);                             \
    } while (0)
#define TOSS_LAUNCH_STEPPER(STREAM, TEAM)                                                                          \
    do {                                                                                                           \
        if (W == kK2Warps) TOSS_LAUNCH_STEPPER_W(STREAM, TEAM, kK2Warps);                                          \
        else TOSS_LAUNCH_STEPPER_W(STREAM, TEAM, kK2WarpsFew);                                                     \
    } while (0)
    if (resident) {
        if (team == 1) TOSS_LAUNCH_STEPPER(false, 1);
        else if (team == 2) TOSS_LAUNCH_STEPPER(false, 2);
        else TOSS_LAUNCH_STEPPER(false, 4);
    } else {
        if (team == 1) TOSS_LAUNCH_STEPPER(true, 1);
        else if (team == 2) TOSS_LAUNCH_STEPPER(true, 2);
        else TOSS_LAUNCH_STEPPER(true, 4);
    }
#undef TOSS_LAUNCH_STEPPER_W
#undef TOSS_LAUNCH_STEPPER
    ctx->launches++;
    TOSS_CHECK_CUDA(cudaGetLastError());
    return 0;
}
