"""Multi-GPU host logic: y-strip decomposition of the rectangle, one process per GPU.

The reference is shared-memory only; this is the decomposition of section 8(e):

* the global cell list (same origin, same ``h``) is cut into strips of bucket rows; a generator is owned by
  the rank whose rows contain its primary bucket; labels stay global;
* before a remesh every rank sends the generators whose buckets (primary or periodic image) fall into a
  neighbour's halo window -- ``H`` bucket rows either side, ``H`` = the largest row offset the reference's
  neighbour walk can reach before it throws (voronoigrid.jl:63-65) -- with ``torch.distributed`` send/recv;
* owned generators come first, ghosts are appended peer by peer in the order they were sent; the library orders
  every bucket by the global label it is given as key, so the local cell list finds labels in the same order as
  the single-GPU (= ``julia -t 1``) run and connectivity does not depend on the GPU count;
* the halo plan needs no negotiation: sender and receiver both use the order of the sent list (primary slots of
  ``sent_sel[q]`` on one side, primary slots of the ghost labels on the other).  Per remesh a rank talks to its
  strip neighbours only -- counts, ghost generators, and (peer-memory halo) one message with its slot addresses
  and CUDA IPC handles; there is no collective over the whole group;
* with peer memory (the default) all of this runs INSIDE the library (csrc/lv_strip.cu): ghost selection, the exchange of
  counts and ghost generators, the halo plan and every halo exchange are kernels that pull from the neighbours' exported
  exchange areas over NVLink; per remesh the host synchronises once for the counts and makes no torch / NCCL call.  The
  torch.distributed path below (``exchange_ghosts`` + ``_build_halo_plan``) is the fallback when CUDA IPC is unavailable
  and the reference for the CPU tests;
* inside the Krylov loop ghost values are pulled out of the neighbours' halo outboxes (packed by the kernel that produces the
  search direction) and the dot products are summed through peer mailboxes (NCCL send/recv + allreduce as the fallback).

Everything in this module is plain ``torch`` + ``torch.distributed`` and works on CPU tensors with the
``gloo`` backend (tests) as well as on CUDA tensors with ``nccl``; the compute calls go to liblvb200.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch
import torch.distributed as dist

from . import _capi
from ._capi import check, load_library, ptr
from .host import PressureSolver, Rectangle, VoronoiGrid

_CHECK_PLAN = bool(int(__import__("os").environ.get("LV_CHECK_PLAN", "0")))  # extra (synchronising) sanity checks


# ---------------------------------------------------------------------------------------------------
# pure planning helpers (device agnostic, unit-tested on CPU)
# ---------------------------------------------------------------------------------------------------
def halo_rows(i2: np.ndarray, rr: np.ndarray, rr_max: float) -> int:
    """Largest |row offset| of a path node the walk can actually visit (rr <= rr_max)."""
    vis = rr <= rr_max
    return int(np.abs(i2[vis]).max()) if vis.any() else 0


def partition_rows(n2: int, row_lo: int, row_hi: int, world: int) -> np.ndarray:
    """Row boundaries R[0..world]: rank r owns bucket rows [R[r], R[r+1]).  The rows [row_lo, row_hi] that
    hold the primary buckets of the domain are split evenly; rank 0 also owns the padding rows below,
    the last rank the padding rows above (they only ever hold periodic images)."""
    inner = row_hi - row_lo + 1
    cuts = [row_lo + (inner * r) // world for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, n2
    return np.asarray(cuts, dtype=np.int64)


def bucket_row(y: torch.Tensor, oy: float, h: float) -> torch.Tensor:
    """0-based bucket row of a y coordinate: floor((y - oy)/h), the arithmetic of findkey (neighborlist.jl:47-52)."""
    return torch.floor((y - oy) / h).to(torch.int64)


def owner_of_rows(rows: torch.Tensor, R: np.ndarray) -> torch.Tensor:
    """Rank owning each bucket row (rows outside [0, n2) are clamped to the first / last rank)."""
    cuts = torch.as_tensor(R[1:-1], dtype=torch.int64, device=rows.device)
    return torch.searchsorted(cuts, rows, right=True)


def ghost_mask_for(y: torch.Tensor, oy: float, h: float, yperiodic: bool, yperiod: float, lo: int, hi: int) -> torch.Tensor:
    """Generators with a primary or y-image bucket row inside [lo, hi) (a neighbour's halo window).
    x-images keep the row, so only the three y-images of insert_periodic! matter (voronoigrid.jl:130-147)."""
    m = torch.zeros_like(y, dtype=torch.bool)
    shifts = (0.0, yperiod, -yperiod) if yperiodic else (0.0,)
    for s in shifts:
        r = bucket_row(y + s * 1.0 if s else y, oy, h)
        m |= (r >= lo) & (r < hi)
    return m


def exchange_variable(send: dict, world: int, rank: int, device, dtype, width: int, group=None) -> dict:
    """Send ``send[q]`` (a [k, width] tensor, possibly empty) to every peer q in ``send`` and return what the
    peers sent to us.  Counts travel first (all_gather), payloads with batched isend/irecv."""
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    for q, t in send.items():
        counts[q] = t.shape[0]
    all_counts = [torch.zeros(world, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    ops, recv = [], {}
    for q in range(world):
        if q == rank:
            continue
        k_in = int(all_counts[q][rank])
        if k_in > 0:
            recv[q] = torch.empty((k_in, width), dtype=dtype, device=device)
            ops.append(dist.P2POp(dist.irecv, recv[q], q, group=group))
        k_out = int(counts[q])
        if k_out > 0:
            ops.append(dist.P2POp(dist.isend, send[q].contiguous(), q, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return recv


def exchange_with_peers(send: dict, peers, device, dtype, width: int, group=None, counts_in: dict = None) -> dict:
    """Neighbour-only variant of exchange_variable: every rank in ``peers`` gets ``send[q]`` ([k, width], possibly
    empty) and sends one back.  No collective over the whole group: counts travel as one-element messages between
    neighbours (skipped when the caller knows them: ``counts_in[q]``), payloads with batched isend/irecv."""
    peers = list(peers)
    if not peers:
        return {}
    if counts_in is None:
        cs = {q: torch.tensor([send[q].shape[0]], dtype=torch.int64, device=device) for q in peers}
        cr = {q: torch.zeros(1, dtype=torch.int64, device=device) for q in peers}
        ops = []
        for q in peers:
            ops.append(dist.P2POp(dist.irecv, cr[q], q, group=group))
            ops.append(dist.P2POp(dist.isend, cs[q], q, group=group))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        got = torch.cat([cr[q] for q in peers]).cpu().tolist()  # one device sync for all peers
        counts_in = {q: int(k) for q, k in zip(peers, got)}
    ops, recv = [], {}
    for q in peers:
        recv[q] = torch.empty((counts_in[q], width), dtype=dtype, device=device)
        if counts_in[q] > 0:
            ops.append(dist.P2POp(dist.irecv, recv[q], q, group=group))
        if send[q].shape[0] > 0:
            ops.append(dist.P2POp(dist.isend, send[q].contiguous(), q, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return recv


class StripPlan:
    """Geometry of the decomposition: who owns which bucket rows and which peers exist."""

    def __init__(self, info: dict, path, h: float, rr_max: float, bmin, bmax, xperiodic: bool, yperiodic: bool, world: int, rank: int):
        self.world, self.rank = world, rank
        self.n1, self.n2 = int(info["n1"]), int(info["n2"])
        self.oy = float(info["origin"][1])
        self.h = float(h)
        self.yperiodic = bool(yperiodic)
        self.yperiod = float(bmax[1] - bmin[1])
        i1, i2, rr = path
        self.H = halo_rows(np.asarray(i2), np.asarray(rr), rr_max)
        row_lo = int(math.floor((bmin[1] - self.oy) / self.h))
        row_hi = int(math.floor((bmax[1] - self.oy) / self.h))
        self.R = partition_rows(self.n2, row_lo, min(row_hi, self.n2 - 1), world)
        self.row_lo, self.row_hi = row_lo, min(row_hi, self.n2 - 1)
        if world > 1 and np.diff(self.R)[1:-1].min(initial=10 ** 9) < self.H:
            raise ValueError(f"strips thinner than the halo ({self.H} bucket rows): use fewer ranks for this grid")
        if world > 1 and min(self.R[1] - row_lo, row_hi + 1 - self.R[-2]) < self.H:
            raise ValueError(f"strips thinner than the halo ({self.H} bucket rows): use fewer ranks for this grid")

    def window(self, q: int):
        """Bucket rows whose entries rank q needs: its own rows plus H rows either side."""
        return max(int(self.R[q]) - self.H, 0), min(int(self.R[q + 1]) + self.H, self.n2)  # rows outside the grid hold nothing

    def peers(self):
        """Ranks whose windows can overlap my rows: the strip neighbours, wrapping when y is periodic."""
        if self.world == 1:
            return []
        cand = {self.rank - 1, self.rank + 1}
        if self.yperiodic:
            cand = {q % self.world for q in cand}
            # images of the top strip land in the bottom padding rows and vice versa
            if self.rank == 0:
                cand.add(self.world - 1)
            if self.rank == self.world - 1:
                cand.add(0)
        return sorted(q for q in cand if 0 <= q < self.world and q != self.rank)

    def owner(self, y: torch.Tensor) -> torch.Tensor:
        return owner_of_rows(bucket_row(y, self.oy, self.h), self.R)

    def near_bounds(self):
        """(A, B, everything): a generator of this strip can only be in a peer's window -- directly or through a
        y-image -- if v = (y - oy)/h satisfies v < A or v >= B.  Derived from the windows themselves (each shifted by
        0, +period, -period and widened by one row for rounding); ``everything`` is set when a window lies strictly
        inside the strip (thin strips), in which case no prefilter is possible."""
        e_lo = max(int(self.R[self.rank]), self.row_lo)
        e_hi = min(int(self.R[self.rank + 1]), self.row_hi + 1)
        shifts = (0.0, self.yperiod / self.h, -self.yperiod / self.h) if self.yperiodic else (0.0,)
        A, B, everything = -math.inf, math.inf, False
        for q in self.peers():
            lo, hi = self.window(q)
            for sr in shifts:
                a, b = lo - sr - 1.0, hi - sr + 1.0
                if a <= e_lo:
                    A = max(A, b)
                elif b >= e_hi:
                    B = min(B, a)
                else:
                    everything = True
        return A, B, everything

    def select_ghosts(self, xy: torch.Tensor, labels: torch.Tensor):
        """Per peer: the rows [x, y, label] of my generators that fall into its halo window, and their
        indices in my owned arrays (the halo plan refers to ghosts by their position in this list).

        Only generators near the strip's edges (``near_bounds``) can be in anybody's window, so one cheap pass over
        all generators picks those and the per-peer window tests run on that small subset."""
        out, sels = {}, {}
        peers = self.peers()
        if not peers:
            return out, sels
        y = xy[:, 1]
        v = (y - self.oy) / self.h                      # bucket_row(y) == floor(v); comparisons with integers need no floor
        A, B, everything = self.near_bounds()
        if everything or A >= B:
            near = torch.arange(y.numel(), device=y.device)
        else:
            near = torch.nonzero((v < A) | (v >= B), as_tuple=False).squeeze(1)
        y_near = y[near]
        for q in peers:
            lo, hi = self.window(q)
            m = ghost_mask_for(y_near, self.oy, self.h, self.yperiodic, self.yperiod, lo, hi)
            sel = near[m]
            payload = torch.empty((sel.numel(), 3), dtype=torch.float64, device=xy.device)
            payload[:, :2] = xy[sel]
            payload[:, 2] = labels[sel].to(torch.float64)  # labels < 2^53 travel exactly in a double
            out[q] = payload
            sels[q] = sel
        return out, sels


def merge_owned_and_ghosts(xy_own, lab_own, recv: dict, rank: int):
    """Local generator set in global-label order.  Returns (xy, labels, owner_rank) tensors."""
    xs, ls, os_ = [xy_own], [lab_own], [torch.full_like(lab_own, rank)]
    for q, t in sorted(recv.items()):
        xs.append(t[:, :2])
        lq = t[:, 2].to(torch.int64)
        ls.append(lq)
        os_.append(torch.full_like(lq, q))
    xy, lab, own = torch.cat(xs), torch.cat(ls), torch.cat(os_)
    order = torch.argsort(lab, stable=True)
    return xy[order].contiguous(), lab[order].contiguous(), own[order].contiguous()


class LocalSet:
    """Generators present on a rank after the ghost exchange: the owned ones first (in the order given), then
    the ghosts grouped by the peer that sent them (in that peer's send order)."""

    def __init__(self, xy, lab, owner, n_own, ghost_range, sent_sel):
        self.xy, self.lab, self.owner, self.n_own = xy, lab, owner, n_own
        self.ghost_range = ghost_range  # peer -> (start, stop) in the local arrays
        self.sent_sel = sent_sel        # peer -> indices (into my owned arrays) of what I sent there


def exchange_ghosts(plan: StripPlan, xy_own: torch.Tensor, lab_own: torch.Tensor, group=None) -> LocalSet:
    """One ghost-generator exchange (collective).  No sorting or merging: the library orders every bucket by
    the global label it is given as ordering key, so ghosts are simply appended."""
    n_own = int(xy_own.shape[0])
    if plan.world > 1:
        send, sels = plan.select_ghosts(xy_own, lab_own)
        recv = exchange_with_peers(send, plan.peers(), xy_own.device, torch.float64, 3, group)
        recv = {q: t for q, t in recv.items() if t.shape[0] > 0}
    else:
        sels, recv = {}, {}
    xs, ls, os_ = [xy_own], [lab_own], [torch.full_like(lab_own, plan.rank)]
    ranges, pos = {}, n_own
    for q in sorted(recv):
        t = recv[q]
        xs.append(t[:, :2])
        lq = t[:, 2].to(torch.int64)
        ls.append(lq)
        os_.append(torch.full_like(lq, q))
        ranges[q] = (pos, pos + int(t.shape[0]))
        pos += int(t.shape[0])
    return LocalSet(torch.cat(xs).contiguous(), torch.cat(ls).contiguous(), torch.cat(os_).contiguous(), n_own, ranges, sels)


def migrate_generators(plan: StripPlan, xy_own: torch.Tensor, lab_own: torch.Tensor, group=None):
    """After a move: hand generators that left the strip to their new owner (collective).  Returns the new
    owned set (xy, labels) in global-label order."""
    if plan.world == 1:
        return xy_own, lab_own
    own = plan.owner(xy_own[:, 1])
    send = {}
    for q in range(plan.world):
        if q == plan.rank:
            continue
        sel = torch.nonzero(own == q, as_tuple=False).squeeze(1)
        pay = torch.empty((sel.numel(), 3), dtype=torch.float64, device=xy_own.device)
        pay[:, :2] = xy_own[sel]
        pay[:, 2] = lab_own[sel].to(torch.float64)
        send[q] = pay
    recv = exchange_variable(send, plan.world, plan.rank, xy_own.device, torch.float64, 3, group)
    keep = own == plan.rank
    xy, lab, _ = merge_owned_and_ghosts(xy_own[keep], lab_own[keep], recv, plan.rank)
    return xy, lab


# ---------------------------------------------------------------------------------------------------
# the distributed grid / solver (GPU)
# ---------------------------------------------------------------------------------------------------
class _DevArray:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, p: int, n: int, typestr: str, shape=None):
        self.__cuda_array_interface__ = {"shape": shape or (n,), "typestr": typestr, "data": (p, False), "version": 3}


class StripGrid:
    """A VoronoiGrid decomposed into y-strips over the ranks of a torch.distributed process group."""

    def __init__(self, boundary_rect: Rectangle, dr: float, h=None, r_max=None, xperiodic=False, yperiodic=False,
                 device: int = 0, group=None, use_peer_memory: bool = True):
        self.group = group
        self.use_peer_memory = use_peer_memory  # False: NCCL send/recv for every halo (baseline path)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.grid = VoronoiGrid(boundary_rect, dr, h=h, r_max=r_max, xperiodic=xperiodic, yperiodic=yperiodic, device=device)
        g = self.grid
        self.dev = torch.device("cuda", device)
        self.plan = StripPlan(g.info(), g.magic_path(), g.h, g.rr_max, boundary_rect.xmin, boundary_rect.xmax, xperiodic,
                              yperiodic, self.world, self.rank)
        self._L = g._L
        if self.world > 1:
            idbuf = torch.zeros(128, dtype=torch.uint8)
            if self.rank == 0:
                arr = (C.c_uint8 * 128)()
                check(self._L.lv_comm_unique_id(arr), None)
                idbuf = torch.tensor(list(arr), dtype=torch.uint8)
            idbuf = idbuf.to(self.dev)
            dist.broadcast(idbuf, 0, group=group)
            raw = (C.c_uint8 * 128)(*idbuf.cpu().tolist())
            check(self._L.lv_comm_init(g._h, self.rank, self.world, raw), g._h)
            if use_peer_memory:  # allreduce mailboxes: every rank maps every other rank's mailbox once
                mine = (C.c_uint8 * 64)()
                check(self._L.lv_mailbox_export(g._h, mine), g._h)
                hb = torch.tensor(list(mine), dtype=torch.uint8, device=self.dev)
                allm = [torch.zeros(64, dtype=torch.uint8, device=self.dev) for _ in range(self.world)]
                dist.all_gather(allm, hb, group=group)
                flat = []
                for t in allm:
                    flat += t.cpu().tolist()
                st = self._L.lv_mailbox_plan(g._h, self.world, (C.c_uint8 * (64 * self.world))(*flat))
                self._agree_on_peer_memory(st == 0)
        self.xy_own = torch.zeros((0, 2), dtype=torch.float64, device=self.dev)
        self.lab_own = torch.zeros(0, dtype=torch.int64, device=self.dev)
        self._strip_ready = False        # library-side strip state (lv_strip_setup) allocated and peers mapped
        self._owned_dirty = True
        self.capacity_factor = 1.25      # head room of the local arrays and ghost outboxes over the first owned set

    def _strip_setup(self) -> None:
        """Collective, once: size and allocate the library's strip state, exchange the CUDA IPC handles of the exchange
        areas and map the neighbours' areas.  Capacities are fixed from here on (nothing is reallocated per remesh)."""
        g, L, plan = self.grid, self._L, self.plan
        peers = plan.peers()
        n_own = int(self.xy_own.shape[0])
        rows = max(1, min(int(plan.R[self.rank + 1]), plan.row_hi + 1) - max(int(plan.R[self.rank]), plan.row_lo))
        est = n_own / rows * (plan.H + 3) * (2 if len(peers) == 1 else 1)          # ghosts a neighbour takes from me
        stat = torch.tensor([float(n_own), float(est)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(stat, op=dist.ReduceOp.MAX, group=self.group)
        n_max, est_max = (float(v) for v in stat.cpu().tolist())
        capg = int(2.0 * self.capacity_factor * est_max) + 8192
        cap_loc = int(self.capacity_factor * n_max) + len(peers) * capg + 4096
        npeer = len(peers)
        # my index in each peer's (sorted) peer list -- the strip neighbourhood is symmetric
        idx_there = []
        for q in peers:
            other = StripPlan.__new__(StripPlan)
            other.__dict__.update(plan.__dict__)
            other.rank = q
            idx_there.append(other.peers().index(self.rank))
        win = [plan.window(q) for q in peers]
        arr = lambda v: (C.c_int32 * max(npeer, 1))(*v)  # noqa: E731
        mine = (C.c_uint8 * 64)()
        check(L.lv_strip_setup(g._h, npeer, arr(peers), arr(idx_there), arr([w[0] for w in win]), arr([w[1] for w in win]),
                               C.c_int64(capg), C.c_int64(cap_loc), mine), g._h)
        ok = True
        if self.world > 1:
            hb = torch.tensor(list(mine), dtype=torch.uint8, device=self.dev)
            allh = [torch.zeros(64, dtype=torch.uint8, device=self.dev) for _ in range(self.world)]
            dist.all_gather(allh, hb, group=self.group)
            flat = []
            for q in peers:
                flat += allh[q].cpu().tolist()
            st = L.lv_strip_map(g._h, (C.c_uint8 * max(64 * npeer, 1))(*flat))
            ok = st == 0
            self._agree_on_peer_memory(ok)
        self._strip_ready = bool(self.use_peer_memory or self.world == 1)
        self._capg, self._cap_loc = capg, cap_loc
        check(L.lv_strip_set_rows(g._h, self.world, (C.c_int32 * (self.world + 1))(*[int(v) for v in plan.R])), g._h)

    def _agree_on_peer_memory(self, ok: bool) -> None:
        """CUDA IPC mapping can be refused (container policy).  All ranks switch to the NCCL path together."""
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            if self.rank == 0:
                import warnings
                warnings.warn("peer-memory halo unavailable (cudaIpcOpenMemHandle failed on some rank): using NCCL send/recv")
            self.use_peer_memory = False
            check(self._L.lv_peer_disable(self.grid._h), self.grid._h)

    # -- generators -----------------------------------------------------------------------------------
    def set_owned(self, xy, labels) -> None:
        """Generators this rank owns (they must lie in its strip), global labels ascending."""
        self.xy_own = torch.as_tensor(xy, dtype=torch.float64, device=self.dev).contiguous()
        self.lab_own = torch.as_tensor(labels, dtype=torch.int64, device=self.dev).contiguous()
        if self.lab_own.numel() and int(self.lab_own.max()) >= 2 ** 30:
            raise ValueError("global labels must stay below 2^30 (the bucket sort packs key and image bit into 32 bits)")
        self._owned_dirty = True

    def set_owned_from_host(self, xy_host: np.ndarray, labels_dev: torch.Tensor) -> None:
        """Owned positions from (pinned) host memory, labels already on the device (end-to-end path)."""
        if self._strip_ready and labels_dev is self.lab_own and not self._owned_dirty:
            # straight into the library's local arrays: one host->device copy, labels unchanged
            n = int(xy_host.shape[0])
            check(self._L.lv_strip_set_owned(self.grid._h, n, ptr(xy_host), 1, None), self.grid._h)
            return
        self.xy_own = torch.from_numpy(xy_host).to(self.dev, non_blocking=True)
        self.lab_own = labels_dev
        self._owned_dirty = True

    def close(self) -> None:
        """Collective teardown: unmap all peer memory, wait for every rank, then free.  Memory exported over CUDA IPC
        must not be freed while another rank still has it mapped."""
        g = self.grid
        if g._h:
            torch.cuda.synchronize(self.dev)
            if self.world > 1:
                check(self._L.lv_peer_close(g._h), g._h)
                dist.barrier(group=self.group)
            self._L.lv_destroy(g._h)
            g._h = None

    def migrate(self) -> None:
        """After a move: hand generators that left the strip to their new owner."""
        self.xy_own, self.lab_own = migrate_generators(self.plan, self.xy_own, self.lab_own, self.group)
        self._owned_dirty = True

    # -- remesh!(grid) --------------------------------------------------------------------------------
    def remesh(self) -> None:
        """remesh!(grid) on the strips (collective).  With peer memory everything runs inside the library; otherwise the
        torch.distributed path below."""
        if self.use_peer_memory or self.world == 1:
            return self._remesh_library()
        return self._remesh_torch()

    def _remesh_library(self) -> None:
        g, L = self.grid, self._L
        sp = torch.cuda.current_stream(self.dev).cuda_stream
        if sp != getattr(self, "_stream_ptr", None):   # lv_set_stream synchronises: only when the stream really changes
            g.set_stream(sp)
            self._stream_ptr = sp
        if not self._strip_ready:
            self._strip_setup()
            if not (self.use_peer_memory or self.world == 1):   # CUDA IPC refused on some rank: all ranks fall back together
                return self._remesh_torch()
        if self._owned_dirty:
            n_own = int(self.xy_own.shape[0])
            self._key_own = self.lab_own.to(torch.int32).contiguous()
            check(L.lv_strip_set_owned(g._h, n_own, ptr(self.xy_own) if n_own else None, 0, ptr(self._key_own) if n_own else None), g._h)
            self._n_own = n_own
            self._owned_dirty = False
        counts = (C.c_int64 * 9)()
        check(L.lv_strip_remesh(g._h, counts), g._h)
        self.n_loc = int(counts[0])
        g._dev_n = self.n_loc
        peers = self.plan.peers()
        self.halo_counts = {q: (int(counts[1 + k]), int(counts[5 + k])) for k, q in enumerate(peers)}
        n_own = self._n_own
        # zero-copy views of the library's local arrays (owned first, then the ghosts peer by peer)
        self.xy_loc = self._dev_tensor(6, "<f8", torch.float64, ncomp=2)
        self.key_loc = self._dev_tensor(7, "<i4", torch.int32)
        self.lab_loc = self.key_loc.to(torch.int64)
        self.mask_loc = (torch.arange(self.n_loc, device=self.dev) < n_own).to(torch.uint8)
        ranges, pos = {}, n_own
        for k, q in enumerate(peers):
            ranges[q] = (pos, pos + int(counts[5 + k]))
            pos += int(counts[5 + k])
        self.local = LocalSet(self.xy_loc, self.lab_loc, None, n_own, ranges, {})

    def _remesh_torch(self) -> None:
        g, L = self.grid, self._L
        loc = exchange_ghosts(self.plan, self.xy_own, self.lab_own, self.group)
        self.local = loc
        self.xy_loc, self.lab_loc, self.owner_loc = loc.xy, loc.lab, loc.owner
        self.n_loc = int(loc.xy.shape[0])
        self.key_loc = loc.lab.to(torch.int32).contiguous()
        self.mask_loc = torch.zeros(self.n_loc, dtype=torch.uint8, device=self.dev)
        self.mask_loc[: loc.n_own] = 1
        stream = torch.cuda.current_stream(self.dev)
        g.set_stream(stream.cuda_stream)
        check(L.lv_remesh_owned_dev(g._h, self.n_loc, ptr(self.xy_loc), ptr(self.mask_loc), ptr(self.key_loc)), g._h)
        g._dev_n = self.n_loc
        if self.world > 1:
            self._build_halo_plan()

    def _dev_tensor(self, which: int, typestr: str, torch_dtype, ncomp: int = 1):
        p, n = C.c_void_p(), C.c_int64()
        check(self._L.lv_device_array(self.grid._h, which, C.byref(p), C.byref(n)), self.grid._h)
        if n.value == 0:
            return torch.zeros((0, ncomp) if ncomp > 1 else 0, dtype=torch_dtype, device=self.dev)
        return torch.as_tensor(_DevArray(p.value, n.value, typestr, shape=(n.value, ncomp) if ncomp > 1 else None), device=self.dev)

    def _build_halo_plan(self) -> None:
        """Ghost slots are filled from their owners.  Both sides use the order of the list the owner sent: the
        receiver's slots are the primary slots of its ghost labels (appended peer by peer in that order), the
        sender's slots are the primary slots of ``sent_sel[q]``.  Periodic-image entries of ghosts need no values
        (edges refer to the primary slot).  (NCCL fallback path; with peer memory the library builds the plan itself.)"""
        g, L, loc = self.grid, self._L, self.local
        prim = self._dev_tensor(1, "<i4", torch.int32)                             # primary slot of every local label
        peers = self.plan.peers()
        send_slots, recv_slots = {}, {}
        empty64 = torch.zeros(0, dtype=torch.int64, device=self.dev)
        for q in peers:
            a, b = loc.ghost_range.get(q, (0, 0))
            recv_slots[q] = prim[a:b]
            send_slots[q] = prim[loc.sent_sel.get(q, empty64)]
        npeer = len(peers)
        s_all = torch.cat([send_slots[q] for q in peers]).contiguous() if npeer else torch.zeros(0, dtype=torch.int32, device=self.dev)
        r_all = torch.cat([recv_slots[q] for q in peers]).contiguous() if npeer else torch.zeros(0, dtype=torch.int32, device=self.dev)
        if _CHECK_PLAN and (s_all.numel() and int(s_all.min()) < 0 or r_all.numel() and int(r_all.min()) < 0):
            raise RuntimeError("halo plan: a generator on the plan has no primary slot here")
        pr = (C.c_int32 * max(npeer, 1))(*peers)
        sc = (C.c_int64 * max(npeer, 1))(*[int(send_slots[q].numel()) for q in peers])
        rc = (C.c_int64 * max(npeer, 1))(*[int(recv_slots[q].numel()) for q in peers])
        self._halo_keep = (s_all, r_all)
        torch.cuda.current_stream(self.dev).synchronize()
        check(L.lv_halo_plan(g._h, npeer, pr, sc, ptr(s_all) if s_all.numel() else None, rc,
                             ptr(r_all) if r_all.numel() else None), g._h)
        self.halo_counts = {q: (int(send_slots[q].numel()), int(recv_slots[q].numel())) for q in peers}

    # -- device-resident stepping on strips (stepping.py functions take ``self.grid``) ---------------------------------
    def state_attach(self) -> None:
        """Make the owned generators the resident state of the lv_step_* sweeps (after ``remesh()``).  Fields are zero until
        ``state_set``.  Needs the library strip path (peer memory, or one rank)."""
        if not self._strip_ready:
            raise RuntimeError("device-resident stepping on strips needs the peer-memory strip exchange")
        check(self._L.lv_state_attach_strip(self.grid._h), self.grid._h)
        self.grid._resident = True

    def n_owned(self) -> int:
        p, n = C.c_void_p(), C.c_int64()
        check(self._L.lv_device_array(self.grid._h, 8, C.byref(p), C.byref(n)), self.grid._h)
        return int(n.value)

    def owned_labels(self) -> torch.Tensor:
        """Global labels of the owned generators in their current local order (it changes when generators migrate)."""
        return self._dev_tensor(8, "<i4", torch.int32).to(torch.int64)

    def state_set(self, name: str, arr) -> None:
        """One field for the owned generators, in the order of ``owned_labels()``."""
        a = np.ascontiguousarray(arr, dtype=np.float64)
        check(self._L.lv_state_set(self.grid._h, name.encode(), ptr(a), self.n_owned()), self.grid._h)

    def state_get(self, name: str) -> np.ndarray:
        nc = {"x": 2, "v": 2, "dv": 2, "momentum": 2, "D": 4}.get(name, 1)
        n = self.n_owned()
        pp, cnt = C.c_void_p(), C.c_int64()
        check(self._L.lv_state_ptr(self.grid._h, name.encode(), C.byref(pp), C.byref(cnt)), self.grid._h)
        if n == 0:
            return np.zeros((0, nc) if nc > 1 else 0)
        t = torch.as_tensor(_DevArray(pp.value, n * nc, "<f8"), device=self.dev)
        torch.cuda.current_stream(self.dev).synchronize()
        a = t.cpu().numpy()
        return a.reshape(n, nc) if nc > 1 else a

    def refresh(self) -> None:
        """After a sweep that moved generators (move / relaxation_step / lloyd): re-read the counts and the views."""
        self._n_own = self.n_owned()
        p, n = C.c_void_p(), C.c_int64()
        check(self._L.lv_device_array(self.grid._h, 7, C.byref(p), C.byref(n)), self.grid._h)
        self.n_loc = int(n.value)
        self.grid._dev_n = self.n_loc
        self.xy_loc = self._dev_tensor(6, "<f8", torch.float64, ncomp=2)
        self.key_loc = self._dev_tensor(7, "<i4", torch.int32)
        self.lab_loc = self.key_loc.to(torch.int64)
        self.mask_loc = (torch.arange(self.n_loc, device=self.dev) < self._n_own).to(torch.uint8)
        self.xy_own, self.lab_own = self.xy_loc[: self._n_own], self.lab_loc[: self._n_own]
        self._owned_dirty = False

    # -- results --------------------------------------------------------------------------------------
    def owned_index(self) -> torch.Tensor:
        """Positions of the owned generators inside the local (label-ordered) arrays."""
        return torch.nonzero(self.mask_loc, as_tuple=False).squeeze(1)

    def mesh_download(self, edges: bool = True):
        """(rowptr, edges, area, centroid) of the LOCAL generator list; ghost rows are empty.  Edge labels are
        1-based local indices; ``self.lab_loc[label-1]`` is the global label."""
        return self.grid.mesh_download(self.n_loc, edges=edges)


class StripSolver(PressureSolver):
    """PressureSolver on a StripGrid: fields are given for the local generator list (ghost entries are
    overwritten by the halo exchange), the solve runs collectively on all ranks."""

    def __init__(self, sgrid: StripGrid, **kw):
        self.sgrid = sgrid
        super().__init__(sgrid.grid, **kw)
