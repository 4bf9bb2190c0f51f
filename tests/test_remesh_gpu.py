"""GPU parity tests of the remesh path (K1 + K2) through the C ABI against the oracle.

Bar (BASELINE.json north_star): neighbour connectivity bit-exact; areas, face lengths and centroids
within 1e-12 relative.  The kernel mirrors the reference's operation order, so vertices are in
fact compared bit for bit as well.
"""
import numpy as np
import pytest

from .conftest import make_points

pytestmark = pytest.mark.gpu

RTOL = 1e-12  # north_star tolerance for areas / face lengths / centroids


def _both(lv, oracle, kind, n_side, xper, yper, seed=0):
    xy, dr, bmin, bmax = make_points(kind, n_side, seed)
    og = oracle.OracleGrid(bmin, bmax, dr, xperiodic=xper, yperiodic=yper)
    og.set_points(xy)
    assert og.remesh() == 0
    g = lv.VoronoiGrid(lv.Rectangle(bmin, bmax), dr, xperiodic=xper, yperiodic=yper)
    g.set_points(xy)
    lv.remesh(g)
    return g, og, xy, dr


def _assert_mesh_equal(g, og, lv, bitwise=True):
    rowptr, edges = og.mesh()
    assert np.array_equal(g.rowptr, rowptr)                       # same degree per polygon
    assert np.array_equal(g.edges["label"], edges["label"])       # connectivity bit-exact, same edge order
    if bitwise:
        assert g.edges.tobytes() == edges.tobytes()               # vertices bit-identical
    a_ref, c_ref = og.area(), og.centroid()
    assert np.allclose(lv.area(g), a_ref, rtol=RTOL, atol=0)
    assert np.allclose(lv.centroid(g), c_ref, rtol=RTOL, atol=RTOL * np.abs(c_ref).max())
    ln = np.hypot(*(g.edges["v1"] - g.edges["v2"]).T)
    ln_ref = np.hypot(*(edges["v1"] - edges["v2"]).T)
    assert np.allclose(ln, ln_ref, rtol=RTOL, atol=0)


@pytest.mark.parametrize("kind,n_side,xper,yper,seed", [
    ("jitter", 64, True, True, 0),
    ("jitter", 256, True, True, 1),
    ("jitter", 37, True, True, 2),
    ("poisson", 96, True, True, 0),
    ("poisson", 64, False, False, 1),
    ("rect2x1", 48, False, False, 2),
    ("rect2x1", 40, True, False, 3),
    ("rect2x1", 40, False, True, 4),
    ("lattice", 32, True, True, 0),     # degenerate: decided by SIGNUM_EPS and cut order, still identical
    ("lattice", 32, False, False, 0),
])
def test_remesh_matches_oracle(lv, oracle, kind, n_side, xper, yper, seed):
    g, og, xy, dr = _both(lv, oracle, kind, n_side, xper, yper, seed)
    _assert_mesh_equal(g, og, lv)


def test_both_kernels_agree_and_fast_path_is_used(lv, oracle, monkeypatch):
    """The linked-slot kernel (default) and the edge-list kernel (LV_CLIP_MODE=plain) must give the same
    bytes; generic inputs must stay on the linked-slot kernel, degenerate lattices may be replayed."""
    for kind, n_side, per in (("jitter", 128, True), ("poisson", 96, False), ("lattice", 48, True)):
        xy, dr, bmin, bmax = make_points(kind, n_side, 11)
        out = {}
        for mode in ("fast", "plain"):
            monkeypatch.setenv("LV_CLIP_MODE", mode)
            g = lv.VoronoiGrid(lv.Rectangle(bmin, bmax), dr, xperiodic=per, yperiodic=per)
            g.set_points(xy); lv.remesh(g)
            out[mode] = (g.rowptr.copy(), g.edges.copy(), lv.area(g).copy(), lv.centroid(g).copy(), g.clip_info())
        monkeypatch.delenv("LV_CLIP_MODE")
        f, p = out["fast"], out["plain"]
        assert np.array_equal(f[0], p[0]) and f[1].tobytes() == p[1].tobytes()
        assert f[2].tobytes() == p[2].tobytes() and f[3].tobytes() == p[3].tobytes()
        assert p[4][0] >= 2
        if kind != "lattice":
            assert f[4] == (0, 0) or f[4] == (1, 0), f[4]     # no anomaly, linked-slot kernel


@pytest.mark.parametrize("M", [1024, 2048])
def test_remesh_invariants_at_scale(lv, M):
    """1M and 4M cells (the size of BASELINE's taylorgreen / rayleightaylor configs): size-independent properties, no
    oracle needed: torus Euler count, tiling, symmetric adjacency, idempotence."""
    xy = lv.synthetic.jittered_lattice(M, 0)
    g = lv.VoronoiGrid(lv.Rectangle((0, 0), (1, 1)), 1.0 / M, xperiodic=True, yperiodic=True)
    g.set_points(xy)
    lv.remesh(g)
    n = M * M
    assert g.rowptr[-1] == 6 * n
    assert abs(lv.area(g).sum() - 1.0) < 1e-10
    lab = g.edges["label"]
    assert lab.min() >= 1 and lab.max() <= n
    rows = np.repeat(np.arange(1, n + 1, dtype=np.int64), np.diff(g.rowptr))
    fwd = np.sort(rows * (n + 1) + lab)
    bwd = np.sort(lab * (n + 1) + rows)
    assert np.array_equal(fwd, bwd)                                # adjacency symmetric
    # idempotence: a second remesh of the same points reproduces the mesh bit for bit
    e0 = g.edges.copy(); r0 = g.rowptr.copy()
    lv.remesh(g)
    assert np.array_equal(g.rowptr, r0) and g.edges.tobytes() == e0.tobytes()


def test_empty_and_tiny_inputs(lv, oracle):
    g = lv.VoronoiGrid(lv.Rectangle((0, 0), (1, 1)), 0.25)       # r_max = 2.5 covers the box
    g.set_points(np.zeros((0, 2)))
    lv.remesh(g)
    assert g.rowptr.tolist() == [0] and g.edges.shape[0] == 0
    xy = np.array([[0.31, 0.42]])
    g.set_points(xy)
    lv.remesh(g)                                                    # one polygon = the whole box, 4 wall edges
    og = oracle.OracleGrid((0, 0), (1, 1), 0.25)
    og.set_points(xy); assert og.remesh() == 0
    _assert_mesh_equal(g, og, lv)
    assert sorted(g.edges["label"].tolist()) == [-4, -3, -2, -1]
    assert abs(lv.area(g)[0] - 1.0) < 1e-15
    # a lone cell in a box wider than r_max: the reference throws (voronoigrid.jl:63-65), so do both
    og2 = oracle.OracleGrid((0, 0), (1, 1), 0.1)
    og2.set_points(xy); assert og2.remesh() == 2
    g2 = lv.VoronoiGrid(lv.Rectangle((0, 0), (1, 1)), 0.1)
    g2.set_points(xy)
    with pytest.raises(lv.LvError, match="destroyed"):
        lv.remesh(g2)


def test_errors_match_reference(lv):
    g = lv.VoronoiGrid(lv.Rectangle((0, 0), (1, 1)), 1 / 50)
    g.set_points(np.array([[0.5, 0.5], [0.52, 0.5]]))
    with pytest.raises(lv.LvError, match="The Voronoi Mesh has been destroyed."):  # voronoigrid.jl:63-65
        lv.remesh(g)
    g.set_points(np.array([[0.5, np.nan], [0.52, 0.5]]))
    with pytest.raises(lv.LvError) as ei:
        lv.remesh(g)
    assert ei.value.status == 3
    # the handle stays usable after an error
    xy, dr, bmin, bmax = make_points("jitter", 20, 0)
    g2 = lv.VoronoiGrid(lv.Rectangle(bmin, bmax), dr, xperiodic=True, yperiodic=True)
    g2.set_points(xy); lv.remesh(g2)
    assert g2.rowptr[-1] == 6 * 400


def test_crowded_bucket_and_capacity_retry(lv, oracle):
    """Clustered generators: buckets with >16 labels (heap-sorted) and polygons that outgrow the
    first shared-memory capacity level, forcing the kernel's retry path."""
    rng = np.random.default_rng(5)
    dr = 1 / 16
    base = rng.random((200, 2))
    th = np.linspace(0, 2 * np.pi, 40, endpoint=False)
    ring = 0.5 + 0.07 * np.stack([np.cos(th), np.sin(th)], 1)     # 40 neighbours around one centre cell
    cluster = 0.25 + 0.01 * rng.random((60, 2))                     # 60 generators inside one bucket
    xy = np.concatenate([[[0.5, 0.5]], ring, cluster, base[(np.hypot(*(base - 0.5).T) > 0.12)]])
    og = oracle.OracleGrid((0, 0), (1, 1), dr)
    og.set_points(xy); assert og.remesh() == 0
    g = lv.VoronoiGrid(lv.Rectangle((0, 0), (1, 1)), dr)
    g.set_points(xy); lv.remesh(g)
    assert np.diff(g.rowptr).max() >= 30
    _assert_mesh_equal(g, og, lv)


def test_set_rects_like_piston(lv, oracle):
    """examples/piston.jl:43-47 moves a wall by mutating the rectangles without rebuilding the cell list."""
    xy, dr, bmin, bmax = make_points("poisson", 32, 7)
    xy = xy * np.array([0.8, 1.0])
    og = oracle.OracleGrid(bmin, bmax, dr); og.set_points(xy)
    og.set_rects((0, 0), (0.8, 1.0), (0, 0), (0.8, 1.0)); assert og.remesh() == 0
    g = lv.VoronoiGrid(lv.Rectangle(bmin, bmax), dr); g.set_points(xy)
    g.set_rects(lv.Rectangle((0, 0), (0.8, 1.0)), lv.Rectangle((0, 0), (0.8, 1.0)))
    lv.remesh(g)
    _assert_mesh_equal(g, og, lv)
    assert abs(lv.area(g).sum() - 0.8) < 1e-12


def test_lazy_edge_download_is_identical(lv):
    """lv_set_async_edges: the edge view arrives on a second stream; after wait_edges it is the same bytes."""
    xy, dr, bmin, bmax = make_points("jitter", 200, 4)
    g = lv.VoronoiGrid(lv.Rectangle(bmin, bmax), dr, xperiodic=True, yperiodic=True)
    g.set_points(xy)
    lv.remesh(g)
    ref = (g.rowptr.copy(), g.edges.copy())
    g.edges[...] = 0
    for _ in range(3):                                      # back-to-back lazy remeshes reuse the staging buffer safely
        lv.remesh(g, lazy=True)
    assert np.array_equal(g.rowptr, ref[0])                 # rowptr, area, centroid are synchronous
    lv.wait_edges(g)
    assert g.edges.tobytes() == ref[1].tobytes()
    lv.remesh(g)                                            # back to the synchronous mode
    assert g.edges.tobytes() == ref[1].tobytes()
    # everything lazy: rowptr, areas and centroids arrive in the background too
    area_ref, cen_ref = lv.area(g).copy(), lv.centroid(g).copy()
    g.edges[...] = 0; g.rowptr[...] = 0; g._area[...] = 0; g._centroid[...] = 0
    for _ in range(3):
        lv.remesh(g, lazy="all")
    lv.wait_edges(g)
    assert np.array_equal(g.rowptr, ref[0]) and g.edges.tobytes() == ref[1].tobytes()
    assert np.array_equal(lv.area(g), area_ref) and np.array_equal(lv.centroid(g), cen_ref)
    lv.remesh(g)
    assert g.edges.tobytes() == ref[1].tobytes()


def test_pipelined_mode_delivers_the_same_bytes(lv, monkeypatch):
    """lv_set_async_edges(h, 3): remesh returns with the clip kernel queued, the mesh crosses PCIe as 20 B/edge and host
    threads of the library expand it.  After wait_edges every output is the synchronous path's, byte for byte -- on a mesh
    of several wire chunks (> 2^18 edges), with walls, and through find_pressure (fields uploaded before the pending
    remesh is completed)."""
    for kind, n_side, per in (("jitter", 480, True), ("poisson", 64, False), ("rect2x1", 40, False)):
        xy, dr, bmin, bmax = make_points(kind, n_side, 4)
        g = lv.VoronoiGrid(lv.Rectangle(bmin, bmax), dr, xperiodic=per, yperiodic=per)
        g.set_points(xy)
        lv.remesh(g)
        ref = (g.rowptr.copy(), g.edges.copy(), lv.area(g).copy(), lv.centroid(g).copy())
        v, P = lv.synthetic.taylor_green_fields(xy)
        for name, val in (("rho", 1.0), ("mass", ref[2]), ("c2", 100.0), ("v", v), ("P", P)):
            getattr(g, name)[...] = val
        s = lv.PressureSolver(g, rtol=1e-10, atol=0.0, itmax=5000)
        lv.find_pressure(s, 0.1 * dr, 2)
        P_ref, it_ref = g.P.copy(), s.iters.copy()
        for rep in range(3):
            g.edges[...] = 0; g.rowptr[...] = 0; g._area[...] = 0; g._centroid[...] = 0
            g.P[...] = P
            lv.remesh(g, lazy="pipeline")
            assert g.edges is None                                # length unknown until the remesh is completed
            lv.remesh(g, lazy="pipeline")                         # completes the first, queues the second
            lv.find_pressure(s, 0.1 * dr, 2)                      # uploads first, then completes the second remesh
            lv.wait_edges(g)
            assert np.array_equal(g.rowptr, ref[0]) and g.edges.tobytes() == ref[1].tobytes()
            assert np.array_equal(lv.area(g), ref[2]) and np.array_equal(lv.centroid(g), ref[3])
            assert np.array_equal(s.iters, it_ref) and np.array_equal(g.P, P_ref)
        # an explicit download of the standing mesh takes the wire format too (lv_mesh_download in mode 3), also after
        # device-resident remeshes, and is settled before the next remesh replaces the mesh
        from lvb200._capi import check
        check(g._L.lv_set_async_edges(g._h, 3), g._h)
        g._lazy_edges = 3
        g.remesh_dev(__import__("torch").from_numpy(xy).cuda())
        rp, ed, ar, ce = g.mesh_download(len(xy))
        assert np.array_equal(rp, ref[0]) and ed.tobytes() == ref[1].tobytes() and np.array_equal(ar, ref[2])
        lv.remesh(g)                                              # back to the synchronous mode
        assert g.edges.tobytes() == ref[1].tobytes()


def test_pipelined_mode_survives_resizing_and_the_strip_api(lv):
    """One handle, generator sets of growing and shrinking size (the position buffers, the chunk-header buffers and the
    staging areas of the pipelined mode are re-allocated on the way), then the strip API's local mesh through the wire
    format: every delivery equals the synchronous one."""
    import torch
    from lvb200._capi import check
    from lvb200.distributed import StripGrid
    g = lv.VoronoiGrid(lv.Rectangle((0.0, 0.0), (1.0, 1.0)), 1 / 64, xperiodic=True, yperiodic=True)
    for M in (20, 64, 33, 330, 48):
        xy = lv.synthetic.jittered_lattice(M, M)
        g2 = lv.VoronoiGrid(lv.Rectangle((0.0, 0.0), (1.0, 1.0)), 1 / M, xperiodic=True, yperiodic=True)
        g2.set_points(xy)
        lv.remesh(g2)
        ref = (g2.rowptr.copy(), g2.edges.copy(), lv.area(g2).copy())
        # the pipelined handle keeps its cell list (dr = 1/64): same generators, another bucket size -> compare on its own terms
        g.set_points(xy)
        lv.remesh(g, lazy="pipeline")
        lv.remesh(g, lazy="pipeline")
        lv.wait_edges(g)
        got = (g.rowptr.copy(), g.edges.copy(), lv.area(g).copy())
        lv.remesh(g)
        assert np.array_equal(got[0], g.rowptr) and got[1].tobytes() == g.edges.tobytes() and np.array_equal(got[2], lv.area(g))
        if M == 64:
            assert np.array_equal(got[0], ref[0]) and got[1].tobytes() == ref[1].tobytes()
    # strip API on one rank: the local mesh of a StripGrid, downloaded in mode 0 and in mode 3
    M = 96
    xy = lv.synthetic.jittered_lattice(M, 3)
    sg = StripGrid(lv.Rectangle((0.0, 0.0), (1.0, 1.0)), 1 / M, xperiodic=True, yperiodic=True, device=0)
    sg.set_owned(xy, np.arange(1, len(xy) + 1))
    sg.remesh()
    rp0, e0, a0, c0 = sg.mesh_download()
    check(sg.grid._L.lv_set_async_edges(sg.grid._h, 3), sg.grid._h)
    sg.grid._lazy_edges = 3
    sg.remesh()
    rp1, e1, a1, c1 = sg.mesh_download()
    assert np.array_equal(rp0, rp1) and e0.tobytes() == e1.tobytes() and np.array_equal(a0, a1) and np.array_equal(c0, c1)
    sg.remesh()                                                    # settles the previous download before the mesh is replaced
    check(sg.grid._L.lv_set_async_edges(sg.grid._h, 0), sg.grid._h)
    sg.grid._lazy_edges = 0
    sg.close()
    torch.cuda.synchronize()


def test_pipelined_mode_on_degenerate_and_replayed_meshes(lv, monkeypatch):
    """The deferred remesh climbs the same ladder: anomalies replay with the edge-list kernel when the remesh is completed,
    meshes whose chains are not closed bit for bit come back as full records, and a destroyed mesh raises at the wait."""
    for name, xy in _degenerate_sets(3).items():
        g = lv.VoronoiGrid(lv.Rectangle((0, 0), (1, 1)), 1 / 24)
        g.set_points(xy)
        try:
            lv.remesh(g)
        except lv.LvError:                                        # this input destroys the mesh: so must the deferred remesh
            lv.remesh(g, lazy="pipeline")
            with pytest.raises(lv.LvError):
                lv.wait_edges(g)
            continue
        ref = (g.rowptr.copy(), g.edges.copy(), lv.area(g).copy())
        g.edges[...] = 0; g.rowptr[...] = 0
        lv.remesh(g, lazy="pipeline")
        lv.wait_edges(g)
        assert np.array_equal(g.rowptr, ref[0]) and g.edges.tobytes() == ref[1].tobytes(), name
        assert np.array_equal(lv.area(g), ref[2]), name
    monkeypatch.setenv("LV_CLIP_FORCE_ANOMALY", "1")
    xy, dr, bmin, bmax = make_points("jitter", 64, 5)
    g = lv.VoronoiGrid(lv.Rectangle(bmin, bmax), dr, xperiodic=True, yperiodic=True)
    g.set_points(xy)
    lv.remesh(g)
    ref = (g.rowptr.copy(), g.edges.copy())
    n_an = g.clip_info()[1]
    g.edges[...] = 0
    lv.remesh(g, lazy="pipeline")
    lv.wait_edges(g)
    assert g.clip_info()[1] == n_an + 1 and g.edges.tobytes() == ref[1].tobytes()
    monkeypatch.delenv("LV_CLIP_FORCE_ANOMALY")
    # voronoigrid.jl:63-65: an isolated generator cannot be closed within r_max
    g2 = lv.VoronoiGrid(lv.Rectangle((0, 0), (1, 1)), 1 / 50)
    g2.set_points(np.array([[0.5, 0.5], [0.52, 0.5]]))
    with pytest.raises(lv.LvError, match="destroyed"):
        lv.remesh(g2)
    lv.remesh(g2, lazy="pipeline")                              # queued: the error belongs to the completion
    with pytest.raises(lv.LvError, match="The Voronoi Mesh has been destroyed."):
        lv.wait_edges(g2)


def _degenerate_sets(seed):
    """Inputs on which the clipping decisions hinge on SIGNUM_EPS and the cut order (excluded from the north star's
    bit-exact claim, yet the GPU path must still reproduce the reference because its fast kernel detects what it cannot
    replay exactly and hands the remesh to the edge-list kernel)."""
    rng = np.random.default_rng(seed)
    out = {}
    g = (np.arange(24) + 0.5) / 24
    X, Y = np.meshgrid(g, g, indexing="ij")
    lat = np.stack([X.ravel(), Y.ravel()], 1)
    out["square lattice"] = lat
    sel = rng.random(len(lat)) < 0.7
    out["lattice with holes"] = lat[sel]
    out["lattice + 1e-13 noise"] = lat + 1e-13 * rng.standard_normal(lat.shape)
    th = np.linspace(0, 2 * np.pi, 48, endpoint=False)
    rings = [0.5 + r * np.stack([np.cos(th), np.sin(th)], 1) for r in (0.1, 0.2, 0.3, 0.4)]
    out["concentric co-circular rings"] = np.concatenate([[[0.5, 0.5]]] + rings)
    base = rng.random((300, 2)) * 0.9 + 0.05
    out["near-duplicate pairs"] = np.concatenate([base, base + 1e-9])
    a = (4 / 3) ** 0.25 / 24; b = (3 / 4) ** 0.25 / 24
    i, j = np.meshgrid(np.arange(-1, 30), np.arange(0, 32), indexing="ij")
    hexp = np.stack([((i + (j % 2) / 2) * a).ravel(), (j * b).ravel()], 1)
    out["hex lattice"] = hexp[(hexp >= 0).all(1) & (hexp <= 1).all(1)]
    out["collinear row + cloud"] = np.concatenate([np.stack([np.linspace(0.05, 0.95, 40), np.full(40, 0.5)], 1), rng.random((200, 2))])
    return out


@pytest.mark.parametrize("per", [False, True])
def test_degenerate_inputs_still_match_the_reference(lv, oracle, per):
    dr = 1 / 24
    kinds = {}
    for name, xy in _degenerate_sets(3).items():
        og = oracle.OracleGrid((0, 0), (1, 1), dr, xperiodic=per, yperiodic=per)
        og.set_points(xy)
        st = og.remesh()
        g = lv.VoronoiGrid(lv.Rectangle((0, 0), (1, 1)), dr, xperiodic=per, yperiodic=per)
        g.set_points(xy)
        if st != 0:
            with pytest.raises(lv.LvError):
                lv.remesh(g)
            continue
        lv.remesh(g)
        rowptr, edges = og.mesh()
        assert np.array_equal(g.rowptr, rowptr), name
        assert g.edges.tobytes() == edges.tobytes(), name
        assert lv.area(g).tobytes() == og.area().tobytes(), name
        kinds[name] = g.clip_info()[0]
    assert len(kinds) >= 5
    print("kernel level per degenerate input:", kinds)


def test_anomaly_replay_ladder(lv, oracle, monkeypatch):
    """When the linked-slot kernel reports a polygon it cannot replay exactly, the whole remesh is replayed by the
    edge-list kernel.  The hook LV_CLIP_FORCE_ANOMALY=1 makes the fast kernel report a few polygons."""
    xy, dr, bmin, bmax = make_points("poisson", 72, 6)
    og = oracle.OracleGrid(bmin, bmax, dr, xperiodic=True, yperiodic=False)
    og.set_points(xy); assert og.remesh() == 0
    monkeypatch.setenv("LV_CLIP_FORCE_ANOMALY", "1")
    g = lv.VoronoiGrid(lv.Rectangle(bmin, bmax), dr, xperiodic=True, yperiodic=False)
    g.set_points(xy)
    lv.remesh(g)
    level, anomalies = g.clip_info()
    assert level >= 2 and anomalies == 1
    _assert_mesh_equal(g, og, lv)
    monkeypatch.delenv("LV_CLIP_FORCE_ANOMALY")
    lv.remesh(g)                                          # the replay is per remesh: next time the fast kernel is back
    assert g.clip_info() == (0, 1) or g.clip_info() == (1, 1)
    _assert_mesh_equal(g, og, lv)
