#!/usr/bin/env python
"""bench.py -- the TOSS population-fitness path on B200 (contract: README of the driver).

A "step" is ONE evaluation of the whole hot path over one population:
    chromosomes -> [cost prediction + deal over the GPUs] -> RK8(7)-13M trajectories through the polyhedral gravity
    field (K2, event points ray-cast inline) -> fixed-step Hermite resampling + covered-space / penalty fitness
    (K3+K4 fused) -> fitness all-gather -> champion.

Workload = BASELINE.json configs[4], the scaling benchmark: ONE population of 32768 chromosomes on the lp 67P mesh
(1828 faces), 1 spacecraft, 2 manoeuvres, 7-day window, default_cfg.toml tolerances, 6048 samples, 6x8x17 voxel grid,
synthetic "init-like" chromosomes.  `--gpus N` splits THAT population over the N GPUs (strong scaling:
4096 per GPU at N = 8, as the reference's scaling script shrinks every worker's share,
scripts/scaling_performance_benchmark.py:126-137); the weak figure (one 32768 island per GPU) rides along in `weak`.
At N = 1 the line also carries the other BASELINE configs in `configs` (C1 lllp / llp pop 16, C2 the 1 M-point gravity
microbench, C3 Use Case 1 through the sequential-spacecraft driver, C4 full mesh pop 4096), each with the CPU port
timed on the same chromosomes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pop P] [--scaling strong|weak]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

W_FACE = 127            # algorithmic FP64 operations per face-point pair (each sqrt/div/log/atan = 1)
FIXED_POS = np.array([-135.13402075, -4089.53592604, 6050.17636635])   # default_cfg.toml initial_x/y/z
N_MAN = 2
MESH_PATHS = {"lllp": "3dmeshes/churyumov-gerasimenko_lllp.pk", "llp": "3dmeshes/churyumov-gerasimenko_llp.pk",
              "lp": "3dmeshes/churyumov-gerasimenko_lp.pk", "full": "3dmeshes/churyumov-gerasimenko.pk"}
COLLISION_SEMANTICS = ("desolver event records = every accepted state inside the risk sphere: hypothesis "
                       "(SURVEY.md App. A.2); collision parity vs real desolver unpinned")


def init_like_population(pop: int, n_man: int, seed: int, final_time: float = 604800.0) -> np.ndarray:
    """genes uniform in the cfg bounds (default_cfg.toml [chromosome], setup_state.py:32)."""
    rng = np.random.default_rng(seed)
    cols = [rng.uniform(0, 1, pop)] + [rng.uniform(-1, 1, pop) for _ in range(3)]
    for _ in range(n_man):
        cols += [rng.uniform(1, final_time - 1, pop), rng.uniform(0, 2.5, pop)]
        cols += [rng.uniform(-1, 1, pop) for _ in range(3)]
    return np.stack(cols, axis=1)


def converged_like_population(pop: int, n_man: int, seed: int, final_time: float = 604800.0) -> np.ndarray:
    """SURVEY.md 8(d) "converged-like": the golden state of the reference's tests (tests/integration_test.py:33)
    perturbed by 2 %, small manoeuvres -> long bound orbits that dip into the risk sphere."""
    x0 = (2.346971623591584122e-01, 6.959989121956766667e-01, -9.249848356174805719e-01, 7.262727928440093628e-01)
    rng = np.random.default_rng(seed)
    cols = [x0[0] * (1 + 0.02 * rng.normal(size=pop))] + [x0[1 + i] + 0.02 * rng.normal(size=pop) for i in range(3)]
    for _ in range(n_man):
        cols += [rng.uniform(1, final_time - 1, pop), rng.uniform(0, 0.05, pop)]
        cols += [rng.uniform(-1, 1, pop) for _ in range(3)]
    return np.stack(cols, axis=1)


def gravity_points(n: int, seed: int = 0) -> np.ndarray:
    """configs[1]: directions ~ normalised N(0,1)^3, radius ~ U[4000, 12500] m."""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    return d * rng.uniform(4000.0, 12500.0, size=(n, 1))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def workload_config(args, pop_total, world):
    return {"workload": f"BASELINE configs[4]: population fitness, ONE population of {pop_total} chromosomes dealt by "
                        f"predicted cost over {world} B200 ({-(-pop_total // world)}/GPU), {args.mesh} 67P mesh, 1 spacecraft, "
                        f"{N_MAN} manoeuvres, 7-day window, rtol=atol=1e-12, 6048 samples, covered-space + penalties + "
                        f"collision fitness (enum 9); the other BASELINE configs are in `configs` (N = 1 only)",
            "mesh": args.mesh, "population": pop_total, "pop_per_gpu": -(-pop_total // world), "maneuvers": N_MAN,
            "chromosomes": "init-like, seed 1005 (SURVEY.md 8d)",
            "parallelism": f"one population dealt over {world} GPU(s)" if args.scaling == "strong"
                           else f"weak: one {args.pop}-chromosome island per GPU x{world}",
            "collided_trajectories": "BOTH arms retire a collided trajectory at its first event point inside the mesh "
                                     "(its fitness is the constant 1e30 either way, udp_initial_condition.py:126-128): "
                                     "equal work; `value_reference_behaviour` integrates them to final_time like the "
                                     "reference does (compute_trajectory.py:154-159)",
            "collision_semantics": COLLISION_SEMANTICS,
            "l2": "256 MiB buffer overwritten between steps (inside the timed region; < 0.02 % of a step)"}


# ----------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's CPU path for this hot path, on the host cores.  The reference's own
    third-party binaries (polyhedral_gravity / desolver / pygmo) cannot be installed here (no network, not in
    the wheelhouse: DESIGN.md), so this times the CPU oracle port (scalar C, not vectorised) -- the one place besides
    cpu_baseline where bench.py may execute oracle/ -- with every host thread, each step a bounded sample of the same
    population, under the SAME collision policy as the GPU arm (early retirement: equal work)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    v, f = O.load_mesh(args.mesh)
    mesh = O.Mesh(v, f)
    p = O.default_params()
    p.stop_on_collision = 1
    grid = O.default_grid()
    pop_total = args.pop if args.scaling == "strong" else args.pop * args.gpus
    xs = init_like_population(pop_total, N_MAN, 1005)
    sample = min(pop_total, max(cores, args.ref_sample if args.ref_sample else 16 * cores))
    for _ in range(args.warmup):
        O.population_fitness(mesh, p, grid, xs[:min(sample, cores)], FIXED_POS, N_MAN, threads=cores)
    t0 = time.perf_counter()
    n_rhs = 0
    for k in range(args.steps):
        lo = (k * sample) % max(pop_total - sample + 1, 1)
        r = O.population_fitness(mesh, p, grid, xs[lo:lo + sample], FIXED_POS, N_MAN, threads=cores)
        n_rhs += int(r["counters"]["n_rhs"].sum())
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "candidate_trajectories_per_sec", "value": value, "unit": "trajectories/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, pop_total, args.gpus),
        "gravity_evals_per_sec": n_rhs / dt,
        "cpu_baseline": {"value": value, "unit": "trajectories/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} of the {pop_total} chromosomes per step x {args.steps} steps "
                                   f"(oracle/toss_oracle.c: scalar C port, not vectorised, {cores} pthreads, early retirement "
                                   f"of collided trajectories like the GPU arm; the reference's own binaries are not "
                                   f"installable here)"},
        "e2e": {"value": value, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def _time_steps(torch, fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps, out


def _traffic():
    """DRAM bytes per trajectory of the stepper, from the committed ncu capture of this tree (profiles/)."""
    f = ROOT / "profiles" / "r2_stepper_traffic.json"
    if not f.exists():
        return None, "no ncu capture of this tree committed under profiles/"
    d = json.loads(f.read_text())
    return d, f"profiles/{f.name} <- {d.get('source')}"


def run_ours(args):
    import torch
    import torch.distributed as dist
    from toss_b200 import _native as N
    from toss_b200 import parallel as P
    import toss_b200 as T

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = N.Context(local_rank)
    _, v, f, _ = T.create_mesh(MESH_PATHS[args.mesh])      # the package's own loader (mesh_utility.py:37-67)
    f = np.ascontiguousarray(f, dtype=np.int32)
    F = len(f)
    ctx.upload_mesh(v, f, 533.0)
    targs = T.setup_parameters()
    ctx.upload_grid(np.asarray(targs.problem.tensor_grid_r, dtype=np.float64),
                    np.asarray(targs.problem.tensor_grid_theta, dtype=np.float64),
                    np.asarray(targs.problem.tensor_grid_phi, dtype=np.float64),
                    np.asarray(targs.problem.bool_tensor))
    # collided trajectories score the constant 1e30: the path retires them at the first event point inside the mesh
    # (identical fitness vector); the CPU arms run under the same policy
    p = N.make_params(stop_on_collision=1)
    pop_total = args.pop if args.scaling == "strong" else args.pop * world
    xs_all = init_like_population(pop_total, N_MAN, 1005)
    xs_dev = torch.from_numpy(xs_all).cuda()               # the WHOLE population resident on every GPU (3.7 MB)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(params=p):
        """inputs resident in HBM: [cost prediction + deal] + kernels + fitness all-gather + champion"""
        flush.fill_(1)
        if world > 1:
            r = P.sharded_population_fitness(ctx, params, xs_dev, FIXED_POS, N_MAN, args.knot_cap)
            return r, torch.argmin(r["fitness"])
        res = ctx.population_fitness(params, xs_dev, FIXED_POS, N_MAN, args.knot_cap)
        return {"fitness": res["fitness"], "local": res}, torch.argmin(res["fitness"])

    # the user-facing call: the pygmo UDP's batch_fitness on HOST decision vectors (every rank holds them all, copies
    # them H2D from pinned memory, evaluates the rows dealt to it, all-gathers, copies the fitness vector D2H)
    targs.mesh.mesh_path = MESH_PATHS[args.mesh]
    (targs.mesh.body, targs.mesh.vertices, targs.mesh.faces,
     targs.mesh.largest_body_protuberant) = T.create_mesh(targs.mesh.mesh_path)
    lb, ub = T.setup_initial_state_domain(FIXED_POS, targs.problem.start_time, targs.problem.final_time, N_MAN,
                                          targs.problem.number_of_spacecrafts, targs.chromosome)
    udp = T.udp_initial_condition(targs, FIXED_POS, lb, ub)
    udp.knot_capacity = args.knot_cap
    dvs_host = xs_all.ravel().copy()

    def step_e2e():
        flush.fill_(1)
        return udp.batch_fitness(dvs_host)

    # ---- FP64 peak (roofline denominator), measured on this box
    fp64_peak = ctx.measure_fp64_peak()

    # ---- warm-up
    for _ in range(args.warmup):
        res, champ = step_device()
    torch.cuda.synchronize()
    loc = res["local"]
    if loc is not None:
        cnt = loc["counters"].cpu().numpy().view(N.COUNTERS_DTYPE)
        n_rhs_local = int(cnt["n_rhs"].sum())
        status_bad = int((cnt["status"] != 0).sum())
        collided = int(loc["collision"].sum().item())
        n_local = len(cnt)
    else:
        n_rhs_local = status_bad = collided = n_local = 0

    # ---- timed: K steps, device-resident inputs
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        res, champ = step_device()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    fit_device = res["fitness"].cpu().numpy()

    # ---- timed: K steps end to end (host buffers)
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_match = bool(np.array_equal(out, fit_device))

    # ---- the dominant kernel alone (stepper) on THIS rank's shard, CUDA events on the launching stream
    if world > 1:
        xs_shard = ctx.deal_rows(xs_dev, res["order"], rank, world)
        ctx.set_queue_order(True)
    else:
        xs_shard = xs_dev
    kernel_s, _ = _time_steps(torch, lambda: ctx.integrate(p, xs_shard, FIXED_POS, N_MAN, args.knot_cap), args.steps) \
        if n_local else (0.0, None)
    ctx.set_queue_order(False)
    kernel_ms = kernel_s * 1e3

    # ---- max over ranks / totals
    t = torch.tensor([ms, e2e_s * 1e3, kernel_ms, float(n_rhs_local), float(status_bad), float(collided)],
                     dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        tmin = t.clone()
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    else:
        tmax, tsum, tmin = t, t, t
    ms, e2e_ms, kernel_ms_max = float(tmax[0]), float(tmax[1]), float(tmax[2])
    n_rhs_total, status_bad, collided = float(tsum[3]), int(tsum[4]), int(tsum[5])

    # ---- the weak figure: one args.pop island per GPU (N > 1, strong runs only; 2 steps)
    weak = None
    if world > 1 and args.scaling == "strong" and not args.no_weak:
        xw = torch.from_numpy(init_like_population(args.pop, N_MAN, 1005 + 7919 * rank)).cuda()
        gathered = torch.empty(args.pop * world, dtype=torch.float64, device="cuda")

        def step_weak():
            flush.fill_(1)
            r = ctx.population_fitness(p, xw, FIXED_POS, N_MAN, args.knot_cap)
            dist.all_gather_into_tensor(gathered, r["fitness"])
            return torch.argmin(gathered)
        step_weak()
        barrier()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        for _ in range(2):
            step_weak()
        w1.record()
        barrier()
        tw = torch.tensor([w0.elapsed_time(w1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        weak = {"value": args.pop * world * 2 / (float(tw[0]) * 1e-3), "unit": "trajectories/s",
                "population": args.pop * world, "scaling": "weak",
                "note": f"one independent {args.pop}-chromosome island per GPU (different seeds), one fitness all-gather; 2 steps"}

    if rank == 0:
        value = pop_total * args.steps / (ms * 1e-3)
        alg_flops_per_launch = n_rhs_local * float(F) * W_FACE            # this rank's stepper launch
        achieved = alg_flops_per_launch / (kernel_ms * 1e-3) / 1e12
        traffic, traffic_src = _traffic()
        line = {
            "metric": "candidate_trajectories_per_sec", "value": value, "unit": "trajectories/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, pop_total, world),
            "gravity_evals_per_sec": n_rhs_total * args.steps / (ms * 1e-3),
            "gravity_evals_per_trajectory": n_rhs_total / pop_total,
            "trajectories_failed": status_bad, "trajectories_collided": collided,
            "e2e_equals_device_path": e2e_match,
            "clocks": clocks,
            "e2e": {"value": pop_total * args.steps / (e2e_ms * 1e-3), "unit": "trajectories/s",
                    "h2d_bytes_per_step": int(xs_all.size * 8 * world),
                    "d2h_bytes_per_step": int((pop_total + 2) * 8 * world),
                    "api": "toss_b200.udp_initial_condition.batch_fitness(host decision vectors) -> host fitness vector"},
            "gpu_launches": int(launches),
            "stepper_ms_rank_min_max": [float(tmin[2]), kernel_ms_max],
            "roofline": {"bound": "fp64", "kernel": "stepper_kernel (K2: RK8(7)-13M + polyhedral gravity)",
                         "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                         "peak_source": "measured live: register-resident DFMA chains on all SMs (toss_measure_fp64_peak); "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "algorithmic_flops_per_launch": alg_flops_per_launch, "kernel_ms": kernel_ms,
                         "traffic": (traffic["dram_bytes_per_trajectory"] * n_local) if traffic else None,
                         "traffic_note": f"{traffic_src}; DRAM bytes per launch scale with the trajectories (the knots "
                                         f"written), the mesh is L2 / shared-memory resident -- the bound is the FP64 "
                                         f"pipe's issue rate, not HBM",
                         "note": f"algorithmic = gravity evaluations x {F} faces x {W_FACE} ops (SURVEY.md 8d), rank 0's shard"},
        }
        if weak is not None:
            line["weak"] = weak
        if world == 1 and not args.no_microbench:
            # the reference's behaviour on the GPU: collided trajectories integrate to final_time
            p0 = N.make_params(stop_on_collision=0)
            step_device(p0)
            s0, r0 = _time_steps(torch, lambda: step_device(p0), 2)
            n0 = int(r0[0]["local"]["counters"].cpu().numpy().view(N.COUNTERS_DTYPE)["n_rhs"].sum())
            line["value_reference_behaviour"] = {
                "value": pop_total / s0, "unit": "trajectories/s", "gravity_evals_per_trajectory": n0 / pop_total,
                "fitness_identical": bool(np.array_equal(r0[0]["fitness"].cpu().numpy(), fit_device)),
                "note": "stop_on_collision = 0: collided trajectories integrate to final_time (compute_trajectory.py:154-159)"}
            line["configs"] = other_configs(ctx, N, T, torch, fp64_peak, args)
            line["converged_like"] = converged_like(ctx, N, p, args, torch, F, fp64_peak)
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as O      # the cpu_baseline leg: the only use of oracle/ in this arm
            line["cpu_baseline"] = cpu_baseline(O, args.mesh, xs_all, N_MAN, 160)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def converged_like(ctx, N, p, args, torch, F, fp64_peak):
    """The same step on a converged-like population (one GPU): bound orbits cost ~25 % more RHS evaluations per
    trajectory than the init-like ones and never collide."""
    _load_mesh(ctx, args.mesh)
    xs = torch.from_numpy(converged_like_population(args.pop, N_MAN, 2005)).cuda()
    ctx.population_fitness(p, xs, FIXED_POS, N_MAN, args.knot_cap)
    s, res = _time_steps(torch, lambda: ctx.population_fitness(p, xs, FIXED_POS, N_MAN, args.knot_cap), 2)
    n_rhs = int(res["counters"].cpu().numpy().view(N.COUNTERS_DTYPE)["n_rhs"].sum())
    ach = n_rhs * float(F) * W_FACE / s / 1e12
    return {"population": args.pop, "chromosomes": "converged-like, seed 2005 (SURVEY.md 8d)", "n_gpus": 1,
            "trajectories_per_sec": args.pop / s, "gravity_evals_per_sec": n_rhs / s,
            "gravity_evals_per_trajectory": n_rhs / args.pop,
            "roofline": {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak}}


def _load_mesh(ctx, mesh):
    import toss_b200 as T
    _, v, f, _ = T.create_mesh(MESH_PATHS[mesh])
    f = np.ascontiguousarray(f, dtype=np.int32)
    ctx.upload_mesh(v, f, 533.0)
    return len(f)


def gravity_microbench(ctx, torch, mesh, fp64_peak):
    """BASELINE configs[1]: 1 M random field points on the mesh, FP64, K1 kernel alone."""
    F = _load_mesh(ctx, mesh)
    pts = torch.from_numpy(gravity_points(1_000_000, 0)).cuda()
    for _ in range(3):
        ctx.gravity(pts)
    s, _ = _time_steps(torch, lambda: ctx.gravity(pts), 5)
    ach = 1e6 * F * W_FACE / s / 1e12
    return {"mesh": mesh, "faces": F, "gravity_evals_per_sec": 1e6 / s, "ms": s * 1e3,
            "roofline": {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak}}


def other_configs(ctx, N, T, torch, fp64_peak, args):
    """BASELINE configs[0], [1], [2], [3] on ONE GPU, each next to the CPU port on the same chromosomes.  The population
    runs use the same collision policy on both sides (early retirement); C3 goes through the sequential-spacecraft
    driver (toss_b200/driver.py = toss.py:71-154) with the bool_tensor feedback and the number_of_spacecrafts = 4
    chunking of the coverage clock."""
    from oracle import oracle as O      # CPU port beside each config (the cpu_baseline leg)
    out = {}
    p1 = N.make_params(stop_on_collision=1)

    def population(mesh, pop, n_man, seed, evals, cpu_per_core):
        F = _load_mesh(ctx, mesh)
        xs_h = init_like_population(pop, n_man, seed)
        xs = torch.from_numpy(xs_h).cuda()
        ctx.population_fitness(p1, xs, FIXED_POS, n_man, args.knot_cap)
        s, res = _time_steps(torch, lambda: ctx.population_fitness(p1, xs, FIXED_POS, n_man, args.knot_cap), evals)
        n_rhs = int(res["counters"].cpu().numpy().view(N.COUNTERS_DTYPE)["n_rhs"].sum())
        ach = n_rhs * float(F) * W_FACE / s / 1e12
        return {"mesh": mesh, "faces": F, "population": pop, "maneuvers": n_man, "evaluations_timed": evals,
                "chromosomes": f"init-like, seed {seed}", "ms_per_evaluation": s * 1e3, "trajectories_per_sec": pop / s,
                "gravity_evals_per_sec": n_rhs / s, "gravity_evals_per_trajectory": n_rhs / pop,
                "us_per_rhs_of_the_longest_trajectory": s * 1e6 / float(
                    res["counters"].cpu().numpy().view(N.COUNTERS_DTYPE)["n_rhs"].max()),
                "roofline": {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak},
                "cpu_baseline": cpu_baseline(O, mesh, xs_h, n_man, cpu_per_core)}

    # C1: 1 spacecraft, 1 manoeuvre, pop 16, init + 2 generations = 3 evaluations; lllp (the config) and llp (the mesh
    # the reference's tests actually load, tests/integration_test.py:26)
    out["C1_lllp_pop16"] = population("lllp", 16, 1, 1001, 3, 16)
    out["C1_llp_pop16"] = population("llp", 16, 1, 1001, 3, 16)
    # C2: the gravity microbench, 1 M points, K1 alone: lp (the config), llp and full beside it
    out["C2_gravity_1M_points"] = {m: gravity_microbench(ctx, torch, m, fp64_peak) for m in ("lp", "llp", "full")}
    # C4: full mesh, pop 4096, M = 2 (number_of_spacecrafts = 1 in the fitness clock)
    p_c4 = N.make_params(stop_on_collision=1, number_of_spacecrafts=1)
    F = _load_mesh(ctx, "full")
    xs_h = init_like_population(4096, N_MAN, 1004)
    xs = torch.from_numpy(xs_h).cuda()
    ctx.population_fitness(p_c4, xs, FIXED_POS, N_MAN, args.knot_cap)
    s, res = _time_steps(torch, lambda: ctx.population_fitness(p_c4, xs, FIXED_POS, N_MAN, args.knot_cap), 2)
    n_rhs = int(res["counters"].cpu().numpy().view(N.COUNTERS_DTYPE)["n_rhs"].sum())
    ach = n_rhs * float(F) * W_FACE / s / 1e12
    out["C4_full_pop4096"] = {"mesh": "full", "faces": F, "population": 4096, "maneuvers": N_MAN,
                              "chromosomes": "init-like, seed 1004", "ms_per_evaluation": s * 1e3,
                              "trajectories_per_sec": 4096 / s, "gravity_evals_per_sec": n_rhs / s,
                              "gravity_evals_per_trajectory": n_rhs / 4096,
                              "roofline": {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                                           "frac": ach / fp64_peak},
                              "cpu_baseline": cpu_baseline(O, "full", xs_h, N_MAN, 4, number_of_spacecrafts=1)}
    # C3: Use Case 1 through the driver: llp, pop 256, 4 sequential spacecraft, 2 manoeuvres, bool_tensor feedback
    from toss_b200 import driver as D
    a = T.setup_parameters()
    a.optimization.population_size, a.optimization.number_of_generations = 256, 2
    pr = a.problem
    init_state = np.array_split([pr.initial_x, pr.initial_y, pr.initial_z] * pr.number_of_spacecrafts,
                                pr.number_of_spacecrafts)
    lb, ub = T.setup_initial_state_domain(init_state[0], pr.start_time, pr.final_time, pr.number_of_maneuvers,
                                          pr.number_of_spacecrafts, a.chromosome)
    D.run_optimization(_fresh_args(T, 256, 2), init_state, lb, ub, seed=3)   # warm-up (contexts, kernels)
    torch.cuda.synchronize()
    a = _fresh_args(T, 256, 2)
    t0 = time.perf_counter()
    run_time, champ_f, champ_x, fit_arr = D.run_optimization(a, init_state, lb, ub, seed=3)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    n_eval = 4 * (1 + 2)
    xs_h = lb + (ub - lb) * np.random.default_rng(3).random((256, len(lb)))   # spacecraft 0's initial population
    out["C3_use_case_1_llp_pop256_x4_spacecraft"] = {
        "mesh": "llp", "faces": 182, "population": 256, "spacecraft": 4, "maneuvers": 2, "generations": 2,
        "population_evaluations": n_eval, "wall_ms": wall * 1e3, "ms_per_evaluation": wall * 1e3 / n_eval,
        "trajectories_per_sec": 256 * n_eval / wall,
        "covered_voxels_after_4_spacecraft": int(np.asarray(a.problem.bool_tensor).sum()),
        "champion_f": [float(np.ravel(c)[0]) for c in champ_f],
        "note": "toss_b200.driver.run_optimization (toss.py:71-154): per spacecraft 3 batch_fitness calls (host decision "
                "vectors), champion trajectory, update_spherical_tensor_grid feedback; wall clock, optimiser included",
        "cpu_baseline": cpu_baseline(O, "llp", xs_h, 2, 64)}
    _load_mesh(ctx, args.mesh)
    return out


def _fresh_args(T, pop, gens):
    a = T.setup_parameters()
    a.optimization.population_size, a.optimization.number_of_generations = pop, gens
    return a


def cpu_baseline(O, mesh_name, xs_all, n_man, per_core, number_of_spacecrafts=4):
    """The CPU oracle port (scalar C, not vectorised) on the box's host cores, bounded sample of the same population,
    same collision policy as the GPU arm (early retirement: equal work) + a smaller sample under the reference's
    behaviour (collided trajectories integrate to final_time).  Reported, not the optimisation target."""
    cores = os.cpu_count() or 1
    v, f = O.load_mesh(mesh_name)
    mesh = O.Mesh(v, f)
    sample = min(len(xs_all), per_core * cores)
    p = O.default_params()
    p.stop_on_collision = 1
    p.number_of_spacecrafts = number_of_spacecrafts
    t0 = time.perf_counter()
    r = O.population_fitness(mesh, p, O.default_grid(), xs_all[:sample], FIXED_POS, n_man, threads=cores)
    dt = time.perf_counter() - t0
    s2 = max(min(sample, cores), sample // 4)
    p.stop_on_collision = 0
    t0 = time.perf_counter()
    r2 = O.population_fitness(mesh, p, O.default_grid(), xs_all[:s2], FIXED_POS, n_man, threads=cores)
    dt2 = time.perf_counter() - t0
    return {"value": sample / dt, "unit": "trajectories/s", "cores": cores, "kind": "port",
            "gravity_evals_per_sec": float(r["counters"]["n_rhs"].sum()) / dt,
            "value_reference_behaviour": s2 / dt2,
            "gravity_evals_per_trajectory_reference_behaviour": float(r2["counters"]["n_rhs"].sum()) / s2,
            "sample": f"first {sample} chromosomes of the same population, oracle/toss_oracle.c (scalar C port, not "
                      f"vectorised) on {cores} pthreads, {dt:.1f} s, early retirement like the GPU arm; "
                      f"value_reference_behaviour: first {s2}, integrated to final_time, {dt2:.1f} s"}


_REAL_STDOUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner with
    printf), so fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pop", type=int, default=32768,
                    help="strong: the ONE population split over the GPUs; weak: the island size per GPU")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"])
    ap.add_argument("--mesh", default="lp", choices=["lllp", "llp", "lp", "full"])
    ap.add_argument("--knot-cap", type=int, default=1024)
    ap.add_argument("--ref-sample", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-microbench", action="store_true")
    ap.add_argument("--no-weak", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
