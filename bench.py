#!/usr/bin/env python
"""bench.py -- the mesh-and-pressure step of LagrangianVoronoi.jl on B200.

One "step" = the hot path of one reference time step on one batch of synthetic input:
2 x remesh! (move.jl:20 and relaxation.jl:72 each call it) + 1 x find_pressure!
(pressure.jl:215-225: operator assembly + niter = 10 x (RHS assembly + Krylov solve to
atol = rtol = 1e-6, itmax = 1000)), on the periodic random-jittered generator box of
SURVEY.md section 8(d) (splitmix64 lattice jitter, Taylor-Green fields, rho = 1, dt = 0.1 dr).

  value     device-resident throughput (inputs already in HBM), Mcell-steps/s over all ranks; the headline leg is the
            16.8M-cell box on one GPU and, for N > 1, 16.8M cells per GPU (weak scaling, box [0,1] x [0,N])
  e2e       the same step through the host-buffer C ABI (pinned host arrays, H2D/D2H inside the timed region), in the
            library's pipelined mode: uploads overlap the queued clipping kernel, the mesh crosses PCIe as 20 B/edge and host
            threads of the library expand it into the 40-byte Edge records (--e2e-mode all: plain lazy downloads)
  submetrics.plain_cg / shuffled_labels   N = 1: the same leg with unpreconditioned CG / with randomly permuted labels
  roofline  the dominant kernel (CSR Voronoi-Laplacian matvec) against the measured HBM peak
  submetrics.strong_64M   second timed leg in EVERY run: the fixed 8192^2 = 67.1M-cell box split into N y-strips
            (N = 1 included), the north star's strong-scaling configuration, with its own e2e and mesh witness
  cpu_baseline / --impl reference   the CPU restatement of the reference (oracle/, OpenMP, all host cores) -- the
            real `julia -t N` cannot run here (no Julia in the image); --impl reference runs it on the SAME
            16.8M-cell configuration as the headline leg

Usage: python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
Under torchrun one rank drives one GPU.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "mesh+pressure step throughput (2 x Voronoi remesh + find_pressure, 10 x Krylov solve)"
UNIT = "Mcell-steps/s"
E2E_MODE = "pipeline"  # --e2e-mode
MATVEC_BYTES_PER_CELL = 100.0   # rowptr 4 + 6*(col 4 + w 8) + diag 8 + x 8 + y 8
CG_BYTES_PER_CELL_ITER = 172.0  # matvec 100 + 3 vector updates x 24
REMESH_BYTES_PER_CELL = 250.0
STRONG_SIDE = 8192              # north star: 64M-cell box for the strong-scaling leg


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--side", type=int, default=4096, help="lattice side M per GPU; n = M*M cells (4096 -> 16M)")
    ap.add_argument("--c0", type=float, default=10.0, help="sound speed of the synthetic fields (c2 = c0^2)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--niter", type=int, default=10)
    ap.add_argument("--cpu-side", type=int, default=1024, help="lattice side of the bounded CPU sample (cpu_baseline leg)")
    ap.add_argument("--ref-budget", type=float, default=300.0, help="--impl reference: wall-clock budget of the timed steps, s")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-mode", default="pipeline", choices=["pipeline", "all"],
                    help="host-buffer leg: pipeline = lv_set_async_edges(h, 3) (deferred remesh, 20 B/edge wire format), all = mode 2")
    ap.add_argument("--no-strong", action="store_true", help="skip the 64M strong-scaling leg")
    ap.add_argument("--no-shuffle", action="store_true", help="N = 1: skip the shuffled-label leg (worst case of Lagrangian drift)")
    ap.add_argument("--strong-side", type=int, default=STRONG_SIDE)
    ap.add_argument("--strong-steps", type=int, default=5, help="timed steps of the strong leg (min with --steps)")
    ap.add_argument("--sweep", action="store_true", help="N = 1: also time the 1M / 4M sizes and c0 = 1000 (submetrics.sweep)")
    ap.add_argument("--krylov", default="pcg", choices=["cg", "pcg", "minres"],
                    help="Krylov method of the GPU arm: cg, pcg (CG + Jacobi preconditioner 1/A_ii, default: 20 %% fewer iterations "
                         "on this workload at the same stopping rule) or minres (the reference's)")
    ap.add_argument("--nccl-halo", action="store_true", help="N > 1: CG halo by ncclSend/Recv instead of peer-memory loads")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="headline leg for N > 1: weak = M^2 cells per GPU (box [0,1]x[0,N]); strong = one M x M box split N ways")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def workload_text(M, My, world, c0, niter):
    n_total = M * My
    return (f"synthetic periodic random-jittered box, {n_total} cells over {world} GPU(s) (lattice {M} x {My}), dr=1/{M}, h=2dr, "
            f"r_max=10dr, Taylor-Green v/P, rho=1, c0={c0}, dt=0.1dr; step = 2 x remesh + find_pressure(niter={niter}, "
            f"rtol=atol=1e-6, itmax=1000)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.stop_flag, self.index = [], False, index
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop_flag = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = max(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in self.rows if len(r) > 3 + k)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (restatement of the reference) on the box's host cores
# ---------------------------------------------------------------------------------------------------------------
def host_threads() -> int:
    """Cores this process may run on.  torchrun exports OMP_NUM_THREADS=1: the CPU arm must not inherit that."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_step_rate(side, c0, niter, seed, steps=1, warmup=0, budget_s=None, solver="minres"):
    """`steps` steps of the workload at side^2 cells on ALL host cores.  With a budget, stops after the step that
    exceeds it (at least one timed step); the number of steps actually timed is reported."""
    from oracle import oracle as orc
    from lvb200 import synthetic
    orc.build()
    orc.set_threads(host_threads())
    dr = 1.0 / side
    xy = synthetic.jittered_lattice(side, seed)
    n = side * side
    g = orc.OracleGrid((0.0, 0.0), (1.0, 1.0), dr, xperiodic=True, yperiodic=True)
    g.set_points(xy)
    assert g.remesh() == 0
    area = g.area()
    v, P = synthetic.taylor_green_fields(xy)
    g.set("rho", 1.0); g.set("mass", area); g.set("c2", c0 * c0); g.set("v", v)

    def one():
        g.set("P", P)
        t0 = time.perf_counter()
        assert g.remesh() == 0
        assert g.remesh() == 0
        t1 = time.perf_counter()
        iters, _ = g.find_pressure(0.1 * dr, niter, rtol=1e-6, atol=1e-6, itmax=1000, solver=solver)
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1, int(iters.sum())

    for _ in range(warmup):
        one()
    t_rem = t_pr = 0.0
    iters_total = done = 0
    t_start = time.perf_counter()
    for _ in range(max(1, steps)):
        a, b, it = one()
        t_rem += a; t_pr += b; iters_total += it; done += 1
        if budget_s is not None and time.perf_counter() - t_start > budget_s:
            break
    total = t_rem + t_pr
    return {"value": n * done / total / 1e6, "unit": UNIT, "cores": orc.get_threads(), "kind": "port", "steps_timed": done,
            "sample": f"{done} step(s) of the same workload at {n} cells (M={side}), {solver.upper()} restatement of the reference on "
                      f"{orc.get_threads()} host threads, {iters_total // done} Krylov iterations/step",
            "remesh_mcells_s": 2 * n * done / t_rem / 1e6, "krylov_mcell_iters_s": n * iters_total / t_pr / 1e6,
            "krylov_iters_per_step": iters_total // done, "s_per_step_sample": total / done}


def run_reference(args):
    """The reference arm: same configuration as our headline leg at N = 1 (the whole args.side^2 box, same c0, niter,
    tolerances), the reference's own Krylov method (MINRES), all host cores.  A full 16.8M-cell step takes ~15-20 s on the
    host, so the number of timed steps is bounded by --ref-budget (reported as steps_timed)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    res = cpu_step_rate(args.side, args.c0, args.niter, args.seed, steps=max(1, args.steps), warmup=1 if args.warmup else 0,
                        budget_s=args.ref_budget)
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * res["s_per_step_sample"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(args.side, args.side, 1, args.c0, args.niter), "cells_total": args.side * args.side,
                       "krylov_iters_per_step": res["krylov_iters_per_step"], "krylov_method": "MINRES (Krylov.jl restatement)",
                       "note": "CPU restatement of the reference (not Julia: no Julia in the image); the host arm always runs the "
                               "single-box configuration of the N = 1 headline leg, whatever --gpus says",
                       "steps_timed": res["steps_timed"], "warmup_steps_run": 1 if args.warmup else 0},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "submetrics": {"remesh_mcells_s": res["remesh_mcells_s"], "krylov_mcell_iters_s": res["krylov_mcell_iters_s"]},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    emit(line)


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
class Env:
    pass


def hash_combine(parts) -> str:
    """16-bit-chunk sums (lv_mesh_hash) -> one 64-bit word, hex."""
    a0, a1, a2, a3 = (int(x) for x in parts[:4])
    return "%016x" % ((a0 + (a1 << 16) + (a2 << 32) + (a3 << 48)) & 0xFFFFFFFFFFFFFFFF)


def expected_hash(M, My, seed):
    try:
        with open(os.path.join(ROOT, "tests", "golden", "mesh_witness.json")) as f:
            return json.load(f).get(f"{M}x{My}_seed{seed}")
    except Exception:
        return None


def leg_single(env, args, M, steps, warmup, c0, with_e2e, profile=True, shuffle_labels=False):
    """One GPU, the whole periodic unit box through the plain VoronoiGrid / PressureSolver handles.
    shuffle_labels: the same generators under a random permutation of their labels -- the limit of Lagrangian drift, where
    a label says nothing about the position any more (label-ordered inputs scatter over the whole cell list)."""
    import torch
    lv, dev, stream = env.lv, env.dev, env.stream
    dr = 1.0 / M
    dt = 0.1 * dr
    n = M * M
    xy = lv.synthetic.jittered_lattice(M, args.seed)
    if shuffle_labels:
        xy = np.ascontiguousarray(xy[np.random.default_rng(args.seed).permutation(n)])
    g = lv.VoronoiGrid(lv.Rectangle((0.0, 0.0), (1.0, 1.0)), dr, xperiodic=True, yperiodic=True, device=env.local)
    g.set_stream(stream.cuda_stream)
    g.set_points(xy)
    xy_dev = torch.from_numpy(xy).to(dev)
    g.remesh_dev(xy_dev)
    solver = lv.PressureSolver(g, solver=args.krylov)
    _, _, area, _ = g.mesh_download(n, edges=False)
    v, P = lv.synthetic.taylor_green_fields(xy)
    f_host = {"mass": area.copy(), "rho": np.ones(n), "c2": np.full(n, c0 ** 2), "P": P, "v": v}
    f_dev = {k: torch.from_numpy(np.ascontiguousarray(a)).to(dev) for k, a in f_host.items()}
    for k, a in f_host.items():
        getattr(g, k)[...] = a

    def step_dev():
        g.remesh_dev(xy_dev)
        g.remesh_dev(xy_dev)
        solver.upload_fields(f_dev["mass"], f_dev["rho"], f_dev["c2"], f_dev["P"], f_dev["v"], device=True)
        iters, _ = solver.find_pressure_dev(dt, args.niter)
        return int(iters.sum())

    out = time_steps(env, g, step_dev, steps, warmup, profile)
    out.update(n=n, n_total=n, parallelism="single GPU", M=M, My=M)
    hv = (C_uint64 * 6)()
    lvcheck(g._L.lv_mesh_hash(g._h, None, hv), g)
    out["checks"] = mesh_checks(env, hv, float(area.sum()), 1.0, n, M, M, None if shuffle_labels else args.seed, periodic=True)

    if with_e2e:
        # every step must solve the same problem (cold P), so the initial P is staged once per step in pinned host memory
        # outside the timed region -- a user's P is simply wherever their arrays are; no reset copy belongs to the step
        from lvb200.host import _host_empty
        p_in = []
        for _ in range(steps + 1):
            a = _host_empty((n,), np.float64)
            a[...] = P
            p_in.append(a)

        def step_e2e():
            # host buffers in, host buffers out: positions / fields go up, rowptr + edges + areas + centroids + P come back.
            # Every mesh output (rowptr, areas, centroids, edge records) is downloaded lazily on a second stream so that
            # it overlaps the next remesh and the pressure solve; the step ends only when every byte is in host memory.
            g.P = p_in.pop() if p_in else g.P
            t = [time.perf_counter()]
            lv.remesh(g, lazy=E2E_MODE); t.append(time.perf_counter())
            lv.remesh(g, lazy=E2E_MODE); t.append(time.perf_counter())
            lv.find_pressure(solver, dt, args.niter); t.append(time.perf_counter())
            lv.wait_edges(g); t.append(time.perf_counter())
            for k in range(4):
                call_ms[k] += 1e3 * (t[k + 1] - t[k])
            return int(solver.iters.sum())

        call_ms = [0.0] * 4
        global E2E_MODE
        try:
            step_e2e()
        except Exception as ex:  # the pipelined mode needs pinned ring buffers and host threads: never lose the line over it
            if E2E_MODE != "pipeline":
                raise
            sys.stderr.write(f"[bench] pipelined host-buffer mode unavailable ({ex!r}); falling back to --e2e-mode all\n")
            E2E_MODE = "all"
            try:
                lv.wait_edges(g)
            except Exception:
                pass
            p_in.append(_host_empty((n,), np.float64)); p_in[-1][...] = P
            step_e2e()
        env.barrier()
        call_ms = [0.0] * 4
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(steps):
            step_e2e()
        ev1.record(stream)
        env.barrier()
        ms_e = max(ev0.elapsed_time(ev1), 1e3 * (time.perf_counter() - t0))
        nnz = int(g.rowptr[-1])
        h2d = 2 * n * 16 + n * 8 * 6                     # 2 x positions + mass, rho, c2, P, v(2)
        # bytes that cross PCIe: the pipelined mode ships 20 B per edge (start vertex + label word, + 16 B per 2^18-edge
        # chunk) and host threads of the library expand them into the 40-byte records the caller reads
        per_edge = 20 if E2E_MODE == "pipeline" else 40
        d2h = 2 * ((n + 1) * 8 + nnz * per_edge + n * 8 + n * 16) + n * 8  # 2 x (rowptr, edges, area, centroid) + P
        e2e = {"value": n * steps / (ms_e / 1e3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": ms_e / steps, "mode": E2E_MODE,
               "host_wall_ms_per_call": dict(zip(("remesh_1", "remesh_2", "find_pressure", "wait_edges"), (c / steps for c in call_ms))),
               "contract": "positions + fields up; rowptr, 40-B edge records, areas, centroids (x2) and P delivered into the caller's "
                           "host buffers" + ("; pipelined: uploads overlap the queued clip kernel, edges cross PCIe as 20 B and are "
                                             "expanded by host threads of the library" if E2E_MODE == "pipeline" else "")}
        # size-independent properties of the result at the full bench size, checked on the host copies after the timed
        # region: Euler count of a periodic triangulation (sum of degrees = 6n), the cells tile the unit box, every
        # pass converged, pressures are finite
        try:
            e2e["checks"] = {"euler_sum_deg_eq_6n": bool(nnz == 6 * n), "area_sum_minus_1": float(lv.area(g).sum() - 1.0),
                             "all_passes_converged": bool((solver.iters < 1000).all()), "P_finite": bool(np.isfinite(g.P).all()),
                             "labels_in_range": bool(g.edges["label"].min() >= 1 and g.edges["label"].max() <= n)}
        except Exception as ex:  # never lose the measurement over a check
            e2e["checks"] = {"error": repr(ex)}
        out["e2e"] = e2e
        lv.wait_edges(g)
    del solver, g, xy_dev, f_dev
    torch.cuda.empty_cache()
    return out


def leg_strips(env, args, M, scaling, steps, warmup, c0, with_e2e, e2e_edges, profile=True):
    """N y-strips of one periodic box (N = 1 allowed), ghost-generator exchange per remesh, halo + all-reduce in CG.
    weak: the box grows to [0,1] x [0,N] (one unit square of M^2 cells per GPU); strong: a fixed M x M box."""
    import torch
    import torch.distributed as dist
    from lvb200.distributed import StripGrid, StripSolver
    lv, dev, stream, world, rank = env.lv, env.dev, env.stream, env.world, env.rank
    dr = 1.0 / M
    dt = 0.1 * dr
    My = M * world if scaling == "weak" else M
    n_total = M * My
    j0, j1 = (My * rank) // world, (My * (rank + 1)) // world
    xy, k = lv.synthetic.jittered_lattice(M, args.seed, rows=(j0, j1), My=My, return_index=True)
    sg = StripGrid(lv.Rectangle((0.0, 0.0), (1.0, My / M)), dr, xperiodic=True, yperiodic=True, device=env.local,
                   use_peer_memory=not args.nccl_halo)
    g = sg.grid
    sg.set_owned(xy, k + 1)
    del xy, k
    sg.migrate()  # lattice strips and bucket-row strips agree up to a row: settle ownership once
    sg.remesh()
    solver = StripSolver(sg, solver=args.krylov)
    _, _, area, _ = sg.mesh_download(edges=False)
    xy_loc = sg.xy_loc.cpu().numpy()
    v, P = lv.synthetic.taylor_green_fields(xy_loc)
    nl = xy_loc.shape[0]
    f_dev = {k2: torch.from_numpy(np.ascontiguousarray(a)).to(dev) for k2, a in
             {"mass": np.where(area > 0, area, 1.0), "rho": np.ones(nl), "c2": np.full(nl, c0 ** 2), "P": P, "v": v}.items()}
    n = int(sg.mask_loc.sum().item())
    halo = getattr(sg, "halo_counts", {})
    parallelism = (f"{world} y-strip(s) ({scaling} scaling), ghost generators exchanged per remesh, "
                   f"CG halo = {'ncclSend/Recv + 2-scalar ncclAllReduce per dot product' if not sg.use_peer_memory else 'NVLink peer-memory loads (CUDA IPC) + 2-scalar all-reduce through peer mailboxes fused into the producer kernels'}; "
                   f"halo (send, recv) per peer = {halo}")
    phase_ev = []

    def step_dev():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(stream)
        sg.remesh()
        sg.remesh()
        ev[1].record(stream)
        solver.upload_fields(f_dev["mass"], f_dev["rho"], f_dev["c2"], f_dev["P"], f_dev["v"], device=True)
        ev[2].record(stream)
        iters, _ = solver.find_pressure_dev(dt, args.niter)
        ev[3].record(stream)
        phase_ev.append(ev)
        return int(iters.sum())

    out = time_steps(env, g, step_dev, steps, warmup, profile)
    last = phase_ev[-steps:]
    out["wall_phase_ms_per_step_rank0"] = {nm: sum(e[i].elapsed_time(e[i + 1]) for e in last) / steps
                                           for i, nm in enumerate(("remesh_x2_incl_ghost_exchange", "field_upload", "find_pressure"))}
    out.update(n=n, n_total=n_total, parallelism=parallelism, M=M, My=My)
    # multi-GPU parity witness: the order-independent hash of all (label, neighbour label) pairs, summed over the ranks,
    # must equal the value a single-GPU run of the same box prints (tests/golden/mesh_witness.json holds the N = 1 values)
    hv = (C_uint64 * 6)()
    lvcheck(g._L.lv_mesh_hash(g._h, lv._capi.ptr(sg.key_loc), hv), g)
    t = torch.tensor([int(x) for x in hv] + [0], dtype=torch.int64, device=dev)
    ta = torch.tensor([float(area[: sg.local.n_own].sum()) if hasattr(sg, "local") else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t)
        dist.all_reduce(ta)
    out["checks"] = mesh_checks(env, t.cpu().tolist(), float(ta.item()), My / M, n_total, M, My, args.seed, periodic=True)
    if with_e2e:
        e2e = e2e_strips(env, args, sg, solver, g, dt, n_total, steps, c0, e2e_edges)
        out["e2e"] = e2e
    sg.close()
    del solver, sg, g, f_dev
    torch.cuda.empty_cache()
    return out


def mesh_checks(env, hv, area_sum, box_area, n_total, M, My, seed, periodic):
    pairs, rows = int(hv[4]), int(hv[5])
    h = hash_combine(hv)
    exp = expected_hash(M, My, seed)
    # sum of degrees = 6n - 2 x (vertices where four cells meet): the reference decides near-degenerate cuts with an ABSOLUTE
    # epsilon (SIGNUM_EPS, polygon.jl:2) while the cut function scales with dr^2, so at dr = 1/8192 a handful of such vertices
    # is expected (about 3e-8 per vertex and candidate) and none at dr = 1/4096
    return {"rows": rows, "rows_eq_cells": rows == n_total, "euler_sum_deg_eq_6n": pairs == 6 * n_total if periodic else None,
            "sum_deg_minus_6n": pairs - 6 * n_total,
            "area_sum_minus_box": area_sum - box_area, "mesh_witness": h, "mesh_witness_expected_from_1gpu": exp,
            "mesh_witness_matches_1gpu": (h == exp) if exp else None}


def time_steps(env, g, step_dev, steps, warmup, profile):
    import torch
    import torch.distributed as dist
    stream, dev, world = env.stream, env.dev, env.world
    for _ in range(warmup):
        step_dev()
    env.barrier()
    g.prof_reset()
    g.prof_enable(bool(profile))
    l0 = g.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(env.local) as clocks:
        env.barrier()
        ev0.record(stream)
        iters_total = 0
        for _ in range(steps):
            iters_total += step_dev()
        ev1.record(stream)
        env.barrier()
    ms = ev0.elapsed_time(ev1)
    launches = g.launch_count() - l0
    prof = {k: g.prof_get(k) for k in ("cells", "clip", "assemble", "matvec", "vecops")}
    g.prof_enable(False)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"ms_max": float(t.item()), "steps": steps, "warmup": warmup, "iters_total": iters_total, "launches": int(launches),
            "prof": prof, "clocks": clocks.summary()}


def e2e_strips(env, args, sg, solver, g, dt, n_total, steps, c0, edges):
    """End-to-end leg of the strip (multi-GPU) API: every step each rank uploads its owned positions and the fields of
    its local generator list from pinned host memory, remeshes twice (rowptr / areas / centroids and -- with `edges` --
    the 40-byte edge records of its local list come back to the host each time, the edge records lazily on a second
    stream), solves, and reads P back.  A watchdog ends the run with ``e2e: null`` rather than hanging the job."""
    import torch
    import torch.distributed as dist
    lv, dev, rank, world = env.lv, env.dev, env.rank, env.world
    from lvb200._capi import EDGE_DTYPE, check, ptr
    from lvb200.host import _host_empty

    n_own, n_loc = int(sg.xy_own.shape[0]), int(sg.n_loc)
    xy_h = _host_empty((n_own, 2), np.float64)
    xy_h[...] = sg.xy_own.cpu().numpy()
    lab_own = sg.lab_own
    xy_loc = sg.xy_loc.cpu().numpy()
    v, P = lv.synthetic.taylor_green_fields(xy_loc)
    _, _, area, _ = sg.mesh_download(edges=False)
    src = {"mass": np.where(area > 0, area, 1.0), "rho": np.ones(n_loc), "c2": np.full(n_loc, c0 ** 2), "P": P, "v": v}
    f_h = {}
    for k, a in src.items():
        f_h[k] = _host_empty(a.shape, np.float64)
        f_h[k][...] = a
    rowptr_h, area_h, cen_h = _host_empty((n_loc + 1,), np.int64), _host_empty((n_loc,), np.float64), _host_empty((n_loc, 2), np.float64)
    P_h = _host_empty((n_loc,), np.float64)
    nnz = g.mesh_nnz()
    cap = nnz + nnz // 16 + 1024
    edge_h = _host_empty((cap,), EDGE_DTYPE) if edges else None
    # pipeline: 20 B/edge wire format expanded by host threads of the library (the ranks of a node share its cores);
    # all: 40-byte records copied lazily on a second stream
    # The expansion needs host cores AND host memory bandwidth, which the ranks of a node share: measured per step, wire format
    # against plain 40-byte DMA: 2 ranks / 16 cores 261 vs 284 ms; 4 ranks / 32 cores 511 vs 410 ms; 8 ranks / 16 cores 945 vs
    # 805 ms -- so it is used with at most 2 ranks per node and at least 6 cores per rank.
    local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
    cores_per_rank = host_threads() / local_world
    wire = edges and E2E_MODE == "pipeline" and local_world <= 2 and cores_per_rank >= 6
    if edges:
        check(g._L.lv_set_async_edges(g._h, 3 if wire else 1), g._h)

    def step():
        sg.set_owned_from_host(xy_h, lab_own)
        for _ in range(2):
            sg.remesh()
            check(g._L.lv_mesh_download(g._h, ptr(rowptr_h), ptr(edge_h), cap if edges else 0, ptr(area_h), ptr(cen_h)), g._h)
        solver.upload_fields(f_h["mass"], f_h["rho"], f_h["c2"], f_h["P"], f_h["v"], device=False)
        solver.find_pressure_dev(dt, args.niter)
        solver.download_P(out=P_h)
        if edges:
            check(g._L.lv_mesh_wait(g._h), g._h)

    step()
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    env.barrier()
    te = torch.tensor([1e3 * (time.perf_counter() - t0)], device=dev, dtype=torch.float64)
    byt = torch.tensor([n_own * 16 + n_loc * 48, 2 * ((n_loc + 1) * 8 + n_loc * 24 + (nnz * (20 if wire else 40) if edges else 0)) + n_loc * 8],
                       device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(byt)
    if edges:
        check(g._L.lv_set_async_edges(g._h, 0), g._h)
    ms = float(te.item())
    return {"value": n_total * steps / (ms / 1e3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(byt[0]),
            "d2h_bytes_per_step": int(byt[1]), "ms_per_step": ms / steps,
            "mode": ("pipeline" if wire else "all") if edges else None, "host_cores_per_rank": cores_per_rank,
            "contract": ("strip API, per rank: positions + fields up; rowptr, "
                         + ("40-B edge records" + (" (20 B on the wire, expanded by host threads), " if wire else ", ") if edges else "")
                         + "areas, centroids (x2) and P down"
                         + ("" if edges else "; edge records stay in HBM (not comparable with the headline e2e)"))}


def run_ours(args):
    import ctypes
    import torch
    import torch.distributed as dist
    import lvb200 as lv
    global C_uint64, lvcheck
    C_uint64 = ctypes.c_uint64

    def lvcheck(st, g):
        lv._capi.check(st, g._h)

    env = Env()
    env.lv = lv
    env.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    env.rank = rank = int(os.environ.get("RANK", "0"))
    env.local = local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: there is no CPU fallback for this path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    env.dev = dev = torch.device("cuda", local)
    # everything (torch ops, NCCL waits, the library's kernels and copies) runs on ONE explicit non-default stream
    env.stream = stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    env.barrier = barrier
    line_holder = {"line": None}

    def bail(what):
        def f():
            line = line_holder["line"]
            if rank == 0 and line is not None:
                line.setdefault("notes", []).append(f"{what} did not finish within its time limit")
                emit(line)
            os._exit(0)
        return f

    # ---- headline leg ------------------------------------------------------------------------------------------
    M = args.side
    if world == 1:
        res = leg_single(env, args, M, args.steps, args.warmup, args.c0, with_e2e=not args.no_e2e)
    else:
        res = leg_strips(env, args, M, args.scaling, args.steps, args.warmup, args.c0, with_e2e=not args.no_e2e, e2e_edges=True)
    line = None
    if rank == 0:
        line = headline_line(args, res, world)
        line_holder["line"] = line

    # ---- the same leg with plain CG, for the record (N = 1 only; a few steps) -------------------------------------------
    if world == 1 and args.krylov == "pcg":
        import copy
        a2 = copy.copy(args)
        a2.krylov = "cg"
        r2 = leg_single(env, a2, M, max(1, min(args.steps, 3)), 2, args.c0, with_e2e=False)
        line["submetrics"]["plain_cg"] = {"ms_per_step": r2["ms_max"] / r2["steps"], "krylov_iters_per_step": r2["iters_total"] // r2["steps"],
                                          "phase_ms_per_step": {k: v[0] / r2["steps"] for k, v in r2["prof"].items()}}

    # ---- worst case of Lagrangian drift: labels carry no spatial information (N = 1 only; a few steps) -----------------
    if world == 1 and not args.no_shuffle:
        r3 = leg_single(env, args, M, max(1, min(args.steps, 3)), 2, args.c0, with_e2e=False, shuffle_labels=True)
        line["submetrics"]["shuffled_labels"] = {
            "what": "same generators, labels randomly permuted (limit of Lagrangian drift: label order unrelated to position)",
            "ms_per_step": r3["ms_max"] / r3["steps"], "krylov_iters_per_step": r3["iters_total"] // r3["steps"],
            "phase_ms_per_step": {k: v[0] / r3["steps"] for k, v in r3["prof"].items()},
            "rows_eq_cells": r3["checks"].get("rows_eq_cells"), "euler_sum_deg_eq_6n": r3["checks"].get("euler_sum_deg_eq_6n")}

    # ---- CPU sample beside it (rank 0, N = 1 only) -------------------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_step_rate(args.cpu_side, args.c0, args.niter, args.seed, 1)
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line["cpu_baseline"]["detail"] = {k: r[k] for k in ("remesh_mcells_s", "krylov_mcell_iters_s", "s_per_step_sample")}

    # ---- strong-scaling leg: the fixed 64M-cell box over N strips (N = 1 included) ----------------------------------
    if not args.no_strong:
        dog = threading.Timer(420.0, bail("strong_64M leg"))
        dog.daemon = True
        dog.start()
        ss = max(1, min(args.steps, args.strong_steps))
        sw = max(1, min(args.warmup, 3))
        sres = leg_strips(env, args, args.strong_side, "strong", ss, sw, args.c0, with_e2e=not args.no_e2e, e2e_edges=False)
        dog.cancel()
        if rank == 0:
            line["submetrics"]["strong_64M"] = leg_summary(args, sres, world, args.strong_side)

    # ---- optional sweep: sizes and conditioning (N = 1) ---------------------------------------------------------
    if args.sweep and world == 1:
        import copy
        sweep = {}
        # c0 = 1000 (the stiffened taylorgreen config: kappa ~ 1e5) is run with plain CG and with the Jacobi preconditioner, which
        # does not pay there (the mass term it equilibrates is negligible); c0 = 10 uses --krylov
        for (m, c0, kry) in ((1024, args.c0, args.krylov), (2048, args.c0, args.krylov), (1024, 1000.0, "cg"), (2048, 1000.0, "cg"),
                             (4096, 1000.0, "cg"), (2048, 1000.0, "pcg")):
            a2 = copy.copy(args)
            a2.krylov = kry
            r = leg_single(env, a2, m, max(1, min(args.steps, 5)), 3, c0, with_e2e=False)
            sweep[f"M{m}_c0_{c0:g}_{kry}"] = leg_summary(a2, r, 1, m)
        if rank == 0:
            line["submetrics"]["sweep"] = sweep

    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def leg_summary(args, res, world, M):
    steps, ms = res["steps"], res["ms_max"]
    prof = res["prof"]
    out = {"cells_total": res["n_total"], "lattice": [res["M"], res["My"]], "n_gpus": world, "steps": steps, "warmup": res["warmup"],
           "ms_per_step": ms / steps, "mcell_steps_s": res["n_total"] * steps / (ms / 1e3) / 1e6,
           "krylov_iters_per_step": res["iters_total"] // steps,
           "phase_ms_per_step": dict({k: v[0] / steps for k, v in prof.items()},
                                     host_and_exchange=(ms - sum(v[0] for v in prof.values())) / steps),
           "wall_phase_ms_per_step_rank0": res.get("wall_phase_ms_per_step_rank0"),
           "gpu_launches": res["launches"], "checks": res["checks"], "clocks": res["clocks"], "parallelism": res["parallelism"]}
    mv_ms = prof["matvec"][0]
    mv_cnt = res["iters_total"] + args.niter * steps
    if mv_ms > 0:
        hbm, _ = peaks()
        bw = MATVEC_BYTES_PER_CELL * res["n"] / (mv_ms / mv_cnt * 1e-3) / 1e9
        out["matvec_gbs_rank0"] = bw
        out["matvec_frac_of_peak"] = bw / hbm
    if "e2e" in res:
        out["e2e"] = res["e2e"]
    return out


def headline_line(args, res, world):
    hbm, peak_src = peaks()
    steps, ms_max, prof, n, n_total = res["steps"], res["ms_max"], res["prof"], res["n"], res["n_total"]
    iters_total = res["iters_total"]
    mv_ms, mv_launched = prof["matvec"]
    # launches that did work: one per CG iteration plus the initial residual of each of the niter passes
    # (launches queued behind the convergence flag exit at once; their time stays in the numerator)
    mv_cnt = iters_total + args.niter * steps
    mv_avg = mv_ms / max(mv_cnt, 1)
    achieved = MATVEC_BYTES_PER_CELL * n / (mv_avg * 1e-3) / 1e9 if mv_avg > 0 else 0.0
    traffic = None
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:  # DRAM bytes per launch measured once with ncu --set full (profiles/), scaled to this run's cell count
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)["k_matvec"]
            traffic = (t["dram_bytes_read"] + t["dram_bytes_write"]) * n / t["cells"]
            break
        except Exception:
            pass
    rem_ms = prof["cells"][0] + prof["clip"][0]
    pr_ms = prof["assemble"][0] + prof["matvec"][0] + prof["vecops"][0]
    scaling = args.scaling if world > 1 else "weak"
    line = {
        "metric": METRIC, "value": n_total * steps / (ms_max / 1e3) / 1e6, "unit": UNIT, "n_gpus": world,
        "steps": steps, "warmup": res["warmup"], "ms_per_step": ms_max / steps, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(res["M"], res["My"], world, args.c0, args.niter),
                   "cells_total": n_total, "cells_rank0": n, "parallelism": res["parallelism"],
                   "l2": "inputs larger than L2 (no flush needed)", "krylov_iters_per_step": iters_total // steps,
                   "krylov_method": {"cg": "CG (north star)", "pcg": "CG with the Jacobi preconditioner 1/A_ii (--krylov cg for plain CG)",
                                     "minres": "MINRES (the reference's method)"}[args.krylov]
                                    + "; the CPU arm runs the reference's MINRES -- all stop at ||r|| <= atol + rtol ||r0||, rtol = atol = 1e-6"},
        "submetrics": {"remesh_mcells_s": 2 * n_total * steps / (rem_ms / 1e3) / 1e6 if rem_ms > 0 else None,
                       "cg_mcell_iters_s": n_total * iters_total / (pr_ms / 1e3) / 1e6 if pr_ms > 0 else None,
                       "s_per_step": ms_max / steps / 1e3, "wall_phase_ms_per_step_rank0": res.get("wall_phase_ms_per_step_rank0"),
                       "phase_ms_per_step": dict({k: v[0] / steps for k, v in prof.items()},
                                                 host_and_exchange=(ms_max - sum(v[0] for v in prof.values())) / steps),
                       "checks": res["checks"]},
        "roofline": {"kernel": "k_matvec (CSR Voronoi-Laplacian matvec + fused p.Ap)", "bound": "hbm", "achieved": achieved,
                     "peak": hbm, "unit": "GB/s", "frac": achieved / hbm if hbm else None, "peak_source": peak_src,
                     "frac_of_nominal_8TBs": achieved / 8000.0, "avg_launch_ms": mv_avg, "launches": mv_cnt, "launches_queued": mv_launched,
                     "algorithmic_bytes_per_launch": MATVEC_BYTES_PER_CELL * n, "traffic": traffic},
        "e2e": res.get("e2e"), "gpu_launches": res["launches"], "clocks": res["clocks"],
    }
    return line


def emit(line: dict) -> None:
    """The ONE JSON line goes to the process's real stdout; everything else (NCCL / torchrun banners that libraries write
    to fd 1) was diverted to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1
C_uint64 = None
lvcheck = None


def main():
    global _REAL_STDOUT, E2E_MODE
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # native libraries (NCCL prints its version banner on stdout) must not pollute the JSON channel
    a = parse()
    E2E_MODE = a.e2e_mode
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
