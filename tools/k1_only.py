import sys; from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from toss_b200 import _native as N
from oracle import oracle as O
from bench import gravity_points
mesh = sys.argv[1] if len(sys.argv) > 1 else "lp"
c = N.Context(0); v, f = O.load_mesh(mesh); c.upload_mesh(v, f, 533.0)
pts = torch.from_numpy(gravity_points(1_000_000, 0)).cuda()
for _ in range(4):
    c.gravity(pts)
torch.cuda.synchronize()
