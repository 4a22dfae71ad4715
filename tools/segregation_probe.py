"""developer experiment: would CTAs that hold only close-pass trajectories (and CTAs that hold none) beat the mix?
Classify the trajectories of the init-like population by `n_event_points > 0` of a run whose event radius is R, then
time the two classes as populations of their own against the whole population: t(close) + t(far) vs t(mix)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from toss_b200 import _native as N
from oracle import oracle as O
from bench import init_like_population, FIXED_POS

pop = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
radii = [float(r) for r in sys.argv[2].split(",")] if len(sys.argv) > 2 else [3000.0, 4000.0, 5000.0]
c = N.Context(0); v, f = O.load_mesh("lp"); c.upload_mesh(v, f, 533.0); c.upload_grid(*O.default_grid())
p = N.make_params(stop_on_collision=1)
xs_h = init_like_population(pop, 2, 1005)


def timed(x_h, reps=2):
    xs = torch.from_numpy(np.ascontiguousarray(x_h)).cuda()
    c.integrate(p, xs, FIXED_POS, 2, 2048); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); c.integrate(p, xs, FIXED_POS, 2, 2048); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    h = c.integrate(p, xs, FIXED_POS, 2, 2048).to_host()
    return best, int(h["counters"]["n_rhs"].sum())


t_mix, rhs_mix = timed(xs_h)
print(f"mix    pop {pop:6d}  {t_mix:8.1f} ms  {rhs_mix / t_mix / 1e3:6.2f} M evals/s", flush=True)
for R in radii:
    pc = N.make_params(stop_on_collision=1, radius_inner=R)
    h = c.integrate(pc, torch.from_numpy(xs_h).cuda(), FIXED_POS, 2, 2048).to_host()
    close = h["counters"]["n_event_points"] > 0
    ta, ra = timed(xs_h[close]); tb, rb = timed(xs_h[~close])
    print(f"R {R:6.0f}: close {int(close.sum()):6d} {ta:8.1f} ms {ra / ta / 1e3:6.2f} M/s | far {int((~close).sum()):6d} "
          f"{tb:8.1f} ms {rb / tb / 1e3:6.2f} M/s | sum {ta + tb:8.1f} ms vs mix {t_mix:8.1f} ({100 * (1 - (ta + tb) / t_mix):+.1f} %)",
          flush=True)
