"""Multi-GPU parity check, run under torchrun on >= 2 GPUs (tests/test_multigpu.py launches it):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tests/multigpu_check.py

Every rank computes the single-GPU mesh and pressure of the whole problem on its own GPU and compares the rows
it owns in the strip decomposition: connectivity and vertices must be bit-identical (independent of the GPU
count), the converged pressure within 1e-8 relative."""
import os
os.environ.setdefault("LV_CHECK_PLAN", "1")  # synchronising sanity checks of the halo plan
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lvb200 as lv  # noqa: E402
from lvb200.distributed import StripGrid, StripSolver  # noqa: E402
from tests.conftest import make_points  # noqa: E402


def mesh_witness(grid, key, dev, world):
    """Order-independent hash of the (label, neighbour label) pairs of the rows this handle owns, summed over the ranks."""
    import ctypes as C
    from lvb200._capi import check, ptr
    hv = (C.c_uint64 * 6)()
    check(grid._L.lv_mesh_hash(grid._h, ptr(key) if key is not None else None, hv), grid._h)
    t = torch.tensor([int(x) for x in hv], dtype=torch.int64, device=dev)
    if world > 1 and key is not None:
        dist.all_reduce(t)
    return tuple(t.cpu().tolist())


def check_case(kind, n_side, xper, yper, rank, world, dev, peer=True, krylov="cg"):
    xy, dr, bmin, bmax = make_points(kind, n_side, 3)
    n = len(xy)
    dt = 0.1 * dr
    # ---- single-GPU reference on this rank's GPU
    g = lv.VoronoiGrid(lv.Rectangle(bmin, bmax), dr, xperiodic=xper, yperiodic=yper, device=dev.index)
    g.set_points(xy)
    lv.remesh(g)
    v, P = lv.synthetic.taylor_green_fields(xy)
    rho = np.where(xy[:, 0] > 0.5 * (bmin[0] + bmax[0]), 3.0, 1.0)
    g.rho[...] = rho; g.mass[...] = rho * lv.area(g); g.c2[...] = 400.0; g.v[...] = v; g.P[...] = P
    s = lv.PressureSolver(g, rtol=1e-12, atol=0.0, itmax=50000)
    vbc = np.array([[0.5, 0.0], [0.0, 0.0], [0.0, 0.0], [0.0, 0.0]])
    lv.find_pressure(s, dt, 3, boundary_velocity=vbc)
    P_ref = g.P.copy()
    # ---- strip decomposition
    sg = StripGrid(lv.Rectangle(bmin, bmax), dr, xperiodic=xper, yperiodic=yper, device=dev.index, use_peer_memory=peer)
    X = torch.from_numpy(xy).to(dev)
    lab = torch.arange(1, n + 1, dtype=torch.int64, device=dev)
    mine = sg.plan.owner(X[:, 1]) == rank
    sg.set_owned(X[mine], lab[mine])
    sg.remesh()
    rowptr, edges, area, cen = sg.mesh_download()
    lab_loc = sg.lab_loc.cpu().numpy()
    owned = sg.mask_loc.cpu().numpy().astype(bool)
    deg = np.diff(rowptr)
    assert (deg[~owned] == 0).all()
    nbad = 0
    for li in np.nonzero(owned)[0]:
        gl = lab_loc[li] - 1
        e_ref = g.edges[g.rowptr[gl]:g.rowptr[gl + 1]]
        e = edges[rowptr[li]:rowptr[li + 1]]
        if len(e) != len(e_ref):
            nbad += 1
            continue
        lbl = e["label"].copy()
        pos = lbl > 0
        lbl[pos] = lab_loc[lbl[pos] - 1]
        if not (np.array_equal(lbl, e_ref["label"]) and e["v1"].tobytes() == e_ref["v1"].tobytes() and e["v2"].tobytes() == e_ref["v2"].tobytes()):
            nbad += 1
    assert nbad == 0, f"{nbad} owned polygons differ from the single-GPU mesh"
    gl_owned = lab_loc[owned] - 1
    assert np.array_equal(area[owned], lv.area(g)[gl_owned])
    # every generator is owned exactly once
    cnt = torch.tensor([int(owned.sum())], device=dev)
    dist.all_reduce(cnt)
    assert int(cnt) == n
    assert mesh_witness(sg.grid, sg.key_loc, dev, world) == mesh_witness(g, None, dev, world)
    # ---- pressure
    gidx = lab_loc - 1
    f = {k: torch.from_numpy(np.ascontiguousarray(a[gidx])).to(dev) for k, a in
         (("mass", rho * lv.area(g)), ("rho", rho), ("c2", np.full(n, 400.0)), ("P", P), ("v", v))}
    ss = StripSolver(sg, rtol=1e-12, atol=0.0, itmax=50000, solver=krylov)
    ss.upload_fields(f["mass"], f["rho"], f["c2"], f["P"], f["v"], device=True)
    iters, relres = ss.find_pressure_dev(dt, 3, vbc_wall=vbc, want_relres=True)
    P_loc = ss.download_P()
    err = np.abs(P_loc[owned] - P_ref[gl_owned]).max() / np.abs(P_ref).max()
    assert krylov != "cg" or (relres < 1e-10).all(), relres
    assert err <= (1e-8 if krylov == "cg" else 1e-6), err
    # ---- move the generators (some change strips), migrate, remesh again: exercises the parity double-buffering of the
    # exchange areas and the sequence words; compared through the mesh witness
    rng = np.random.default_rng(11)
    for rep in range(3):
        xy2 = xy + 0.35 * dr * rng.standard_normal(xy.shape)
        xy2 = np.clip(xy2, np.asarray(bmin) + 1e-9, np.asarray(bmax) - 1e-9)
        g.x[...] = xy2
        lv.remesh(g, edges=False)
        own_now = sg.lab_own.cpu().numpy() - 1
        sg.set_owned(torch.from_numpy(xy2[own_now]).to(dev), sg.lab_own)
        sg.migrate()
        sg.remesh()
        w1, w0 = mesh_witness(sg.grid, sg.key_loc, dev, world), mesh_witness(g, None, dev, world)
        assert w1 == w0, (rep, w1, w0)
    if rank == 0:
        print(f"case {kind} n={n} per=({xper},{yper}) world={world} peer={sg.use_peer_memory} {krylov}: owned={int(owned.sum())} "
              f"local={len(lab_loc)} halo={sg.halo_counts} iters={iters.tolist()} P err={err:.2e}", flush=True)
    used_peer = sg.use_peer_memory
    del ss
    sg.close()
    return used_peer


def check_stepping(kind, n_side, xper, yper, rank, world, dev, two_phase=False, nsteps=4):
    """Device-resident stepping on strips vs one GPU: the canonical step! (examples/gresho.jl:100-114; with gravity and the
    multiphase projector of examples/rayleightaylor.jl:90-102 when two_phase) for a few steps.  Generators cross the strip
    boundaries and migrate with their fields; per-generator results are compared through the global labels."""
    S = lv.stepping
    xy, dr, bmin, bmax = make_points(kind, n_side, 5)
    n = len(xy)
    rng = np.random.default_rng(2)
    if two_phase:   # Rayleigh-Taylor style: heavy fluid above a wavy interface, walls in y, a smooth flow that vanishes at the walls
        dt = 0.1 * dr
        sy = np.sin(np.pi * (xy[:, 1] - bmin[1]) / (bmax[1] - bmin[1])) ** 2
        v = 0.4 * np.stack([np.sin(2 * np.pi * xy[:, 0]) * sy, np.cos(2 * np.pi * xy[:, 0]) * sy * np.cos(np.pi * xy[:, 1])], 1)
        up = xy[:, 1] > 0.5 * (bmin[1] + bmax[1]) + 0.05 * np.cos(2 * np.pi * xy[:, 0])
    else:
        dt = 0.2 * dr
        v = 0.6 * lv.synthetic.taylor_green_fields(xy)[0] + 0.05 * rng.standard_normal((n, 2))
        up = np.zeros(n, dtype=bool)
    rho = np.where(up, 1.8, 1.0)
    fields = {"v": v, "rho": rho, "P": 10.0 - 0.3 * rho * xy[:, 1], "mu": np.full(n, 2e-3), "phase": np.where(up, 0.0, 1.0),
              "quality": np.ones(n), "dv": np.zeros((n, 2)), "c2": np.full(n, 14.0)}

    def init_fields(area, sel):
        f = {k: a[sel].copy() for k, a in fields.items()}
        f["mass"] = f["rho"] * area
        f["e"] = 0.5 * (f["v"] ** 2).sum(1) + f["P"] / (f["rho"] * 0.4)
        return f

    def step(grid, solver, probe=None):
        ops = [("move", lambda: S.move(grid, dt))]
        if two_phase:
            ops.append(("gravity", lambda: S.gravity_step(grid, (0.0, -1.0), dt)))
        ops += [("eos", lambda: S.ideal_eos(grid, 1.4, 0.0)), ("find_pressure", lambda: S.find_pressure_resident(solver, dt)),
                ("pressure_step", lambda: S.pressure_step(grid, dt)), ("find_D", lambda: S.find_D(grid)),
                ("viscous", lambda: S.viscous_step(grid, dt, True)), ("find_dv", lambda: S.find_dv(grid, dt))]
        if two_phase:
            ops.append(("projection", lambda: S.multiphase_projection(grid, rtol=1e-11, atol=1e-11, itmax=3000)))
        ops.append(("relaxation", lambda: S.relaxation_step(grid, dt)))
        for name, op in ops:
            op()
            if probe is not None and name not in ("move", "relaxation"):
                bad = [nm for nm in ("v", "e", "P", "rho", "dv", "mass") if not np.isfinite(probe(nm)).all()]
                assert not bad, f"rank {rank}: non-finite {bad} after {name}"

    # ---- one GPU
    g = lv.VoronoiGrid(lv.Rectangle(bmin, bmax), dr, xperiodic=xper, yperiodic=yper, device=dev.index)
    g.set_points(xy)
    lv.remesh(g, edges=False)
    for k, a in init_fields(lv.area(g), slice(None)).items():
        getattr(g, k)[...] = a
    S.to_device(g)
    sol = lv.PressureSolver(g, rtol=1e-12, atol=0.0, itmax=50000)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(nsteps):
            step(g, sol, probe=(lambda nm: S.state_get(g, nm)) if os.environ.get("LV_DEBUG_STEP") else None)
    S.from_device(g)
    # ---- strips
    sg = StripGrid(lv.Rectangle(bmin, bmax), dr, xperiodic=xper, yperiodic=yper, device=dev.index)
    X = torch.from_numpy(xy).to(dev)
    lab = torch.arange(1, n + 1, dtype=torch.int64, device=dev)
    mine = sg.plan.owner(X[:, 1]) == rank
    sg.set_owned(X[mine], lab[mine])
    sg.remesh()
    sg.state_attach()
    own0 = sg.owned_labels().cpu().numpy() - 1
    _, _, area_loc, _ = sg.mesh_download(edges=False)
    for k, a in init_fields(area_loc[: len(own0)], own0).items():
        sg.state_set(k, a)
    lv._capi.check(sg._L.lv_state_remesh(sg.grid._h), sg.grid._h)     # migrate (nobody moves yet) + strip remesh on the resident state
    sg.refresh()
    ssol = StripSolver(sg, rtol=1e-12, atol=0.0, itmax=50000)
    moved = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(nsteps):
            before = set(sg.owned_labels().cpu().numpy().tolist())
            step(sg.grid, ssol, probe=sg.state_get if os.environ.get("LV_DEBUG_STEP") else None)
            sg.refresh()
            moved += len(set(sg.owned_labels().cpu().numpy().tolist()) - before)
    own = sg.owned_labels().cpu().numpy() - 1
    cnt = torch.tensor([len(own), moved], device=dev)
    dist.all_reduce(cnt)
    assert int(cnt[0]) == n, (int(cnt[0]), n)                          # every generator still has exactly one owner
    worst = 0.0
    for nm in ("x", "v", "e", "rho", "P", "mass"):
        a, b = sg.state_get(nm), getattr(g, nm)[own]
        err = np.abs(a - b).max() / np.abs(getattr(g, nm)).max()
        worst = max(worst, err)
        assert err <= (1e-5 if two_phase else 1e-8), (nm, err)   # the projector is singular: its MINRES stops a few digits short
    assert mesh_witness(sg.grid, sg.key_loc, dev, world) is not None
    if rank == 0:
        print(f"stepping {kind} n={n} per=({xper},{yper}) world={world} two_phase={two_phase}: {nsteps} steps, "
              f"{int(cnt[1])} generators changed strips, worst field error {worst:.2e}", flush=True)
    del ssol
    sg.close()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    peer_ok = check_case("jitter", 192, True, True, rank, world, dev)
    check_case("poisson", 160, False, False, rank, world, dev)
    check_case("rect2x1", 128, True, False, rank, world, dev)
    check_case("jitter", 128, True, True, rank, world, dev, krylov="minres")
    check_case("jitter", 128, True, True, rank, world, dev, peer=False)       # NCCL fallback path
    if peer_ok:
        check_stepping("jitter", 96, True, True, rank, world, dev)
        check_stepping("jitter", 80, True, False, rank, world, dev, two_phase=True)
    dist.barrier()
    if rank == 0:
        print(f"MULTIGPU OK (peer memory {'used' if peer_ok else 'UNAVAILABLE: NCCL fallback everywhere'})", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
