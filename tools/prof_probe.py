import os, sys; from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
os.environ.setdefault("TOSS_B200_LPT", "0")   # the instrumented counters share the queue area with the order arrays
os.environ["TOSS_B200_LIB"] = str(ROOT / "build" / "variants" / (sys.argv[2] if len(sys.argv) > 2 else "libtoss_prof.so"))
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from toss_b200 import _native as N
from oracle import oracle as O
from bench import init_like_population, converged_like_population, FIXED_POS
c = N.Context(0); v, f = O.load_mesh("lp"); c.upload_mesh(v, f, 533.0); c.upload_grid(*O.default_grid())
p = N.make_params(stop_on_collision=1); pop = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
kind = sys.argv[3] if len(sys.argv) > 3 else 'init'
xs = torch.from_numpy((init_like_population if kind == 'init' else converged_like_population)(pop, 2, 1005)).cuda()
c.integrate(p, xs, FIXED_POS, 2, 1024); torch.cuda.synchronize()
pr = c.integrate(p, xs, FIXED_POS, 2, 1024); torch.cuda.synchronize()
q = pr.ws[pr.L.queue + 16:pr.L.queue + 64].view(torch.int64).cpu().numpy()
wait, comp, total, n, wait0, idle = q
q2 = pr.ws[pr.L.queue + 16:pr.L.queue + 16 + 8 * 24].view(torch.int64).cpu().numpy()
print('per-wid wait/comp:', [f'{q2[8 + w] / max(q2[16 + w], 1):.2f}' for w in range(8)])
import time
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); c.integrate(p, xs, FIXED_POS, 2, 1024); e1.record(); torch.cuda.synchronize()
print(f"{sys.argv[2] if len(sys.argv) > 2 else 'base'}: {e0.elapsed_time(e1):.1f} ms  warps {n}: wait {wait / total:.3f}  compute {comp / total:.3f}  other {(total - wait - comp) / total:.3f} wait@tile0 {wait0 / total:.3f} wait-while-idle {idle / total:.3f}  mean warp lifetime {total / n / 1.965e6:.1f} ms")
