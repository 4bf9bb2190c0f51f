"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: strip partition, ghost-generator exchange,
migration and the variable-size exchange primitive.  The oracle's global mesh is the checker: every Voronoi
neighbour of an owned generator must be present on the owning rank after the ghost exchange."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from .conftest import make_points


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kind, n_side, xper, yper, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import lvb200
        from lvb200.distributed import StripPlan, exchange_ghosts, exchange_variable, migrate_generators
        from oracle import oracle as orc
        xy, dr, bmin, bmax = make_points(kind, n_side, 5)
        og = orc.OracleGrid(bmin, bmax, dr, xperiodic=xper, yperiodic=yper)
        og.set_points(xy)
        assert og.remesh() == 0
        rowptr, edges = og.mesh()
        info = og.info()
        plan = StripPlan(info, og.magic_path(), og.h, og.r_max ** 2, bmin, bmax, xper, yper, world, rank)
        # -- exchange_variable: ragged payloads, empty messages
        send = {qq: torch.full((rank + 2 * qq, 1), float(10 * rank + qq), dtype=torch.float64) for qq in range(world) if qq != rank}
        got = exchange_variable(send, world, rank, torch.device("cpu"), torch.float64, 1)
        for qq, t in got.items():
            assert t.shape[0] == qq + 2 * rank and (t == 10 * qq + rank).all()
        # -- ownership is a partition
        X = torch.from_numpy(xy)
        lab = torch.arange(1, len(xy) + 1, dtype=torch.int64)
        owner = plan.owner(X[:, 1])
        mine = owner == rank
        cnt = torch.tensor([int(mine.sum())])
        dist.all_reduce(cnt)
        assert int(cnt) == len(xy)
        # -- ghost exchange: all Voronoi neighbours of owned generators are present locally
        loc = exchange_ghosts(plan, X[mine], lab[mine])
        xy_loc, lab_loc, own_loc = loc.xy, loc.lab, loc.owner
        assert loc.n_own == int(mine.sum()) and bool((own_loc[: loc.n_own] == rank).all())   # owned first
        have = set(lab_loc.tolist())
        assert len(have) == lab_loc.numel()                                  # no generator twice
        for qq, (a, b) in loc.ghost_range.items():
            assert bool((own_loc[a:b] == qq).all())
        for i in torch.nonzero(mine).squeeze(1).tolist():
            nb = edges["label"][rowptr[i]:rowptr[i + 1]]
            assert all(int(j) in have for j in nb if j > 0), (rank, i)
        n_ghost = int((own_loc != rank).sum())
        assert n_ghost > 0
        assert np.array_equal(xy_loc[own_loc == rank].numpy(), xy[mine.numpy()])
        # -- migration after a move: every generator ends up at its owner, nothing lost
        rng = np.random.default_rng(7)
        moved = xy[mine.numpy()] + rng.normal(0, 3 * dr, (int(mine.sum()), 2))
        if yper:
            moved[:, 1] = bmin[1] + np.mod(moved[:, 1] - bmin[1], bmax[1] - bmin[1])
        else:
            moved[:, 1] = np.clip(moved[:, 1], bmin[1], bmax[1])
        x2, l2 = migrate_generators(plan, torch.from_numpy(moved), lab[mine])
        assert bool((plan.owner(x2[:, 1]) == rank).all())
        tot = torch.tensor([x2.shape[0], int(l2.sum())])
        dist.all_reduce(tot)
        assert int(tot[0]) == len(xy) and int(tot[1]) == len(xy) * (len(xy) + 1) // 2
        q.put((rank, "ok", n_ghost))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,n_side,xper,yper", [("jitter", 64, True, True), ("poisson", 56, False, False), ("rect2x1", 40, True, False)])
def test_strip_decomposition_world2(kind, n_side, xper, yper):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kind, n_side, xper, yper, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[1] == "ok", r[2]


def test_partition_and_halo_rows():
    from lvb200.distributed import halo_rows, partition_rows, owner_of_rows
    R = partition_rows(100, 6, 93, 4)
    assert R[0] == 0 and R[-1] == 100 and (np.diff(R) > 0).all()
    rows = torch.arange(0, 100)
    own = owner_of_rows(rows, R)
    assert int(own.min()) == 0 and int(own.max()) == 3
    for r in range(4):
        sel = rows[own == r]
        assert int(sel.min()) == R[r] and int(sel.max()) == R[r + 1] - 1
    i2 = np.array([0, -1, 1, 6, -6, 7]); rr = np.array([0.0, 0.0, 0.0, 25.0, 25.0, 36.0])
    assert halo_rows(i2, rr, 25.0) == 6


@pytest.mark.parametrize("world,yper,ymax", [(2, True, 1.0), (4, True, 1.03), (8, False, 2.0), (4, True, 0.517)])
def test_select_ghosts_prefilter_equals_window_test(world, yper, ymax):
    """select_ghosts looks only at generators near the strip edges; the result must equal the plain window test on
    every generator, also when the period is not a multiple of the bucket size and for generators that have left
    the strip since the last migration."""
    from lvb200.distributed import StripPlan, ghost_mask_for
    from oracle import oracle as orc
    dr = 1.0 / 96
    bmin, bmax = (0.0, 0.0), (1.0, ymax)
    og = orc.OracleGrid(bmin, bmax, dr, xperiodic=True, yperiodic=yper)
    info, path = og.info(), og.magic_path()
    rng = np.random.default_rng(world)
    X = torch.from_numpy(np.column_stack([rng.uniform(0, 1, 20000), rng.uniform(0, ymax, 20000)]))
    lab = torch.arange(1, 20001, dtype=torch.int64)
    for rank in range(world):
        plan = StripPlan(info, path, og.h, og.r_max ** 2, bmin, bmax, True, yper, world, rank)
        mine = plan.owner(X[:, 1]) == rank
        # a few strays that belong elsewhere by now (moved, not yet migrated)
        stray = torch.zeros_like(mine)
        stray[rng.integers(0, 20000, 50)] = True
        xy, lb = X[mine | stray], lab[mine | stray]
        out, sels = plan.select_ghosts(xy, lb)
        assert sorted(out) == plan.peers()
        for q in plan.peers():
            lo, hi = plan.window(q)
            ref = torch.nonzero(ghost_mask_for(xy[:, 1], plan.oy, plan.h, plan.yperiodic, plan.yperiod, lo, hi)).squeeze(1)
            assert torch.equal(sels[q], ref), (rank, q)
            assert torch.equal(out[q][:, 2].to(torch.int64), lb[sels[q]])
