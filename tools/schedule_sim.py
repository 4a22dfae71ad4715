"""developer tool: event-driven model of the stepper's schedule on one GPU (148 CTAs x 16 warps, the 16 warps of a CTA
advance through the tile ring in lock step, one RHS evaluation per warp per ring cycle; the cycle time depends on how
many warps the busiest scheduler of the SM holds -- calibrated by tools/calib_probe.py).  Used to compare queue orders,
team plans and helper schemes on the measured length distribution before spending GPU time.
usage: schedule_sim.py [gpurun_out/dist_lp.npz] [world]"""
import sys, heapq
import numpy as np

G1 = {1: 17.3, 2: 22.0, 3: 27.8, 4: 34.9}       # us per ring cycle, busiest scheduler holds 1..4 single-warp owners
TEAM4 = {1: 9.4, 2: 10.5, 3: 12.0, 4: 13.9}     # us per cycle, n teams of 4 on the SM
TEAM2 = {1: 12.1, 2: 12.6, 3: 14.3, 4: 14.8, 5: 16.4, 6: 17.9, 7: 19.7, 8: 21.6}
NSM, NW = 148, 16


def cycle_single(owners):
    """owners: set of warp ids that own a trajectory (one warp each)"""
    if not owners:
        return 0.0
    load = [0, 0, 0, 0]
    for w in owners:
        load[w & 3] += 1
    return G1[max(load)]


def cycle_helped(owners):
    """dynamic helpers: idle warps of the CTA take rows of the owners' tiles (at most 4 warps per owner per tile)"""
    m = len(owners)
    if m == 0:
        return 0.0
    plain = cycle_single(owners)
    if m > 4:
        return plain
    return min(plain, TEAM4[m] * 1.05)


def simulate(lengths, order, cycle=cycle_single, interleave=True, team=1, ctas=NSM):
    """lengths[i] = RHS evaluations of trajectory i; order = queue (trajectory indices).  Returns makespan in ms."""
    pop = len(order)
    slots = NW // team
    first_round = min(pop // ctas, slots)
    first_extra = pop % ctas if pop // ctas < slots else 0
    # static round
    cta_owned = [dict() for _ in range(ctas)]   # warp -> remaining evals
    taken = 0
    for b in range(ctas):
        mine = first_round + (1 if b < first_extra else 0)
        for tm in range(mine):
            pos = tm * ctas + b if interleave else b * first_round + min(b, first_extra) + tm
            if pos < pop:
                cta_owned[b][tm * team] = float(lengths[order[pos]])
    cursor = ctas * first_round + first_extra
    # event loop: each CTA advances in lock step; event = next finish in a CTA
    tnow = [0.0] * ctas

    def cyc(b):
        if team == 1:
            return cycle(set(cta_owned[b].keys()))
        n = len(cta_owned[b])
        return (TEAM4 if team == 4 else TEAM2)[n] if n else 0.0
    heap = []
    for b in range(ctas):
        # warps without a static trajectory fetch at t = 0
        for tm in range(slots):
            if tm * team not in cta_owned[b] and cursor < pop and first_round == slots:
                pass
        if cta_owned[b]:
            c = cyc(b)
            heapq.heappush(heap, (min(cta_owned[b].values()) * c, b))
    # idle warps of partially filled CTAs fetch from the cursor immediately (first_round < slots means queue is empty)
    end = 0.0
    while heap:
        tfin, b = heapq.heappop(heap)
        c = cyc(b)
        adv = (tfin - tnow[b]) / c if c > 0 else 0.0
        done = []
        for w in cta_owned[b]:
            cta_owned[b][w] -= adv
            if cta_owned[b][w] <= 1e-6:
                done.append(w)
        tnow[b] = tfin
        for w in done:
            if cursor < pop:
                cta_owned[b][w] = float(lengths[order[cursor]])
                cursor += 1
            else:
                del cta_owned[b][w]
        if cta_owned[b]:
            c = cyc(b)
            heapq.heappush(heap, (tnow[b] + min(cta_owned[b].values()) * c, b))
        else:
            end = max(end, tnow[b])
    return end * 1e-3


if __name__ == "__main__":
    d = np.load(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/dist_lp.npz")
    n, cost = d["n_rhs"], d["cost"]
    key = np.clip(cost.astype(np.int64) >> 2, 0, 4095)
    order_all = np.argsort(-key, kind="stable")
    ideal_rate = 16 / G1[4] * NSM   # evals per us, whole GPU
    print("full pop, LPT blocks:", simulate(n, order_all, interleave=False), "ms; ideal", n.sum() / ideal_rate * 1e-3)
    for world in (2, 4, 8):
        sh = order_all[0::world]
        ideal = n[sh].sum() / ideal_rate * 1e-3
        base = simulate(n, sh)
        helped = simulate(n, sh, cycle=cycle_helped)
        true_order = sh[np.argsort(-n[sh], kind="stable")]
        oracle_lpt = simulate(n, true_order)
        oracle_helped = simulate(n, true_order, cycle=cycle_helped)
        print(f"world {world}: ideal {ideal:.1f}  base {base:.1f} ({ideal/base:.3f})  helpers {helped:.1f} ({ideal/helped:.3f})  "
              f"true-LPT {oracle_lpt:.1f} ({ideal/oracle_lpt:.3f})  true-LPT+helpers {oracle_helped:.1f} ({ideal/oracle_helped:.3f})")


def simulate_trace(lengths, order, cycle=cycle_single, interleave=True, ctas=NSM, widths=None, cap=None):
    """like simulate (team 1) but returns per-trajectory (start, end, cta) and supports per-CTA owner caps
    (widths[b] = max concurrent owners of CTA b; the other warps help: cycle(owners) decides what that buys)"""
    pop = len(order)
    widths = widths or [NW] * ctas
    owned = [dict() for _ in range(ctas)]
    tnow = [0.0] * ctas
    start = np.zeros(pop); endt = np.zeros(pop); where = np.zeros(pop, dtype=int)
    cursor = 0
    # static interleaved round: position p -> CTA p % ctas while it has a free owner slot
    slot_pos = {}
    maxw = max(widths)
    for tm in range(maxw):
        for b in range(ctas):
            if tm < widths[b] and cursor < pop:
                # warps are spread over the schedulers: owner j sits on scheduler j % 4
                owned[b][tm] = [float(lengths[order[cursor]]), cursor]
                where[cursor] = b
                cursor += 1
    heap = []
    for b in range(ctas):
        if owned[b]:
            heapq.heappush(heap, (min(v[0] for v in owned[b].values()) * cycle(set(owned[b])), b))
    dry_time = None
    while heap:
        tfin, b = heapq.heappop(heap)
        c = cycle(set(owned[b]))
        adv = (tfin - tnow[b]) / c
        tnow[b] = tfin
        for w in list(owned[b]):
            owned[b][w][0] -= adv
            if owned[b][w][0] <= 1e-6:
                endt[owned[b][w][1]] = tfin
                if cursor < pop:
                    owned[b][w] = [float(lengths[order[cursor]]), cursor]
                    start[cursor] = tfin
                    where[cursor] = b
                    cursor += 1
                    if cursor == pop:
                        dry_time = tfin
                else:
                    del owned[b][w]
        if owned[b]:
            heapq.heappush(heap, (tnow[b] + min(v[0] for v in owned[b].values()) * cycle(set(owned[b])), b))
    return start * 1e-3, endt * 1e-3, where, (dry_time or 0) * 1e-3
