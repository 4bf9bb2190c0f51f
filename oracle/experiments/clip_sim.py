"""clip_sim.py -- development tool: replay the neighbour-walk traces of oracle/experiments/clip_trace.c under different
warp-scheduling policies of the GPU clipping kernel (K2, lv_clip_fast.cu) and count warp-level work with a cost model
calibrated on the kernel's own counters (LV_CLIP_STATS=1) and its ncu instruction attribution.

  python oracle/experiments/clip_sim.py /tmp/trace.bin tile            # the production tile kernel
  python oracle/experiments/clip_sim.py /tmp/trace.bin tile_v1         # the kernel before the scan-policy changes
  python oracle/experiments/clip_sim.py /tmp/trace.bin refill          # lane refill (old round), tile_merged, refill_merged

Measured after the scan-policy changes (QLOW=1, EV=4, BMAX=2): rounds/tile 21.02, alive/round 23.00, A 64.4 @ 18.50,
B 34.0 @ 16.75, C 17.4 @ 17.81 -- the simulator says 21.04 / 23.23 / 64.4 @ 18.70 / 34.2 @ 16.87 / 17.6 @ 17.85.

Measured on B200 at 16.8M cells (tile kernel): rounds/tile 16.99, alive/round 21.81, A 98.4 events/tile @ 12.17 lanes,
B 51.4 iterations/tile @ 11.70 lanes, C 15.1 cuts/tile @ 20.59 lanes, 970 warp instructions per cell
(A 26 %, B 29 %, C 35 %, emission 8 %).
"""
import sys

import numpy as np

EV_DT = np.dtype([("t", "<i4"), ("kind", "<i4"), ("d2", "<f8"), ("pb", "<f8"), ("pa", "<f8"), ("nv", "<i4"), ("pad", "<i4")])

# warp instructions per warp-level execution of a code section (from ncu shares / LV_CLIP_STATS counts)
COST = {"A": 82.0, "B": 175.0, "C": 720.0, "EMIT": 2480.0, "ROUND": 35.0,
        "B_POP": 55.0, "B_CLS": 120.0, "A_ADV": 40.0, "A_TEST": 42.0}


def load(path):
    raw = open(path, "rb").read()
    n, n1, n2, npath = np.frombuffer(raw, "<i8", 4, 0)
    off = 32
    polys = []
    for _ in range(int(n)):
        ip, bucket, ne, deg = np.frombuffer(raw, "<i8", 4, off)
        off += 32
        ev = np.frombuffer(raw, EV_DT, int(ne), off)
        off += EV_DT.itemsize * int(ne)
        polys.append((int(bucket), int(ip), ev, int(deg)))
    polys.sort(key=lambda p: (p[0], p[1]))  # slot order of the cell list: (bucket, label)
    return polys


class Lane:
    __slots__ = ("ev", "kind", "d2", "pa", "i", "prr", "q", "scan_done", "alive", "plus", "n", "deg")

    def __init__(self, ev, deg):
        self.kind = ev["kind"].tolist()
        self.d2 = ev["d2"].tolist()
        self.pa = ev["pa"].tolist()
        self.n = len(self.kind)
        self.i = 0               # scan cursor (next event)
        self.prr = float(ev["pb"][0])
        self.q = []
        self.scan_done = False
        self.alive = True
        self.plus = -1
        self.deg = deg

    # one scan event of phase A; returns (advanced, tested)
    def scan_event(self):
        adv = tst = False
        k = self.kind[self.i]
        if k == 0 or k >= 5:  # bucket exhausted: enter the next node
            adv = True
            if k >= 5:
                if k == 6 or self.d2[self.i] > self.prr:
                    self.scan_done = True
                # else: stale radius -- the real kernel would scan on into nodes the reference never saw; modelled as a stall
                return adv, tst
            if self.d2[self.i] > self.prr:  # cannot happen before the reference's own end (prr only shrinks)
                self.scan_done = True
                return adv, tst
            self.i += 1
            k = self.kind[self.i]
            if k == 0 or k >= 5:
                return adv, tst  # empty / out-of-bounds bucket: the event is spent
        tst = True
        if not (self.d2[self.i] > self.prr):
            self.q.append(self.i)
        self.i += 1
        return adv, tst


def sim_tile_policy(lanes, QCAP=4, EV=4, QLOW=1, BMAX=2, skip_self=True, stats=None):
    """The production kernel (lv_clip_round.inc): a scan pass runs only while some lane has <= QLOW candidates queued, EV
    events per pass, the generator's own entry never queued, at most BMAX pops per lane and round, then the cut."""
    cost = 0.0
    for l in lanes:
        l.plus = -1
    while any(l.alive for l in lanes):
        if any(l.alive and not l.scan_done and len(l.q) <= QLOW for l in lanes):
            for _ in range(EV):
                sc = [l for l in lanes if l.alive and not l.scan_done and len(l.q) < QCAP
                      and not (l.kind[l.i] == 5 and not l.d2[l.i] > l.prr)]
                if not sc or not any(len(l.q) <= QLOW for l in sc):
                    break
                for l in sc:
                    l.scan_event()
                    if skip_self and l.q and l.kind[l.q[-1]] == 1:
                        l.q.pop()
                stats["a_it"] += 1; stats["a_ln"] += len(sc)
                cost += COST["A"]
        b_max = b_sum = 0
        for l in lanes:
            nb = 0
            while l.alive and l.q and l.plus < 0 and nb < BMAX:
                j = l.q.pop(0)
                nb += 1
                k = l.kind[j]
                if k == 1 or l.d2[j] > l.prr:
                    continue
                if k == 4:
                    l.plus = j
            b_max = max(b_max, nb); b_sum += nb
            if l.alive and l.plus < 0 and l.scan_done and not l.q:
                l.alive = False
        stats["b_it"] += b_max; stats["b_ln"] += b_sum
        cost += COST["B"] * b_max
        stats["rounds"] += 1; stats["alive"] += sum(l.alive for l in lanes)
        cut = [l for l in lanes if l.alive and l.plus >= 0]
        if cut:
            for l in cut:
                l.prr = l.pa[l.plus]; l.plus = -1
            stats["c_it"] += 1; stats["c_ln"] += len(cut)
            cost += COST["C"]
        cost += COST["ROUND"]
    return cost + COST["EMIT"]


def sim_tile_kernel(lanes, QCAP=4, EV=8, stats=None):
    """The round-1 / first-session kernel: rounds of A (scan whenever a queue has room) / B (pop until a vertex is outside) / C."""
    cost = 0.0
    while any(l.alive for l in lanes):
        # phase A
        for _ in range(EV):
            sc = [l for l in lanes if l.alive and not l.scan_done and len(l.q) < QCAP]
            if not sc:
                break
            for l in sc:
                l.scan_event()
            stats["a_it"] += 1; stats["a_ln"] += len(sc)
            cost += COST["A"]
        # phase B
        b_max = b_sum = 0
        for l in lanes:
            l.plus = -1
            nb = 0
            while l.alive and l.q and l.plus < 0:
                j = l.q.pop(0)
                nb += 1
                k = l.kind[j]
                if k == 1 or l.d2[j] > l.prr:
                    continue
                if k == 4:
                    l.plus = j
            b_max = max(b_max, nb); b_sum += nb
            if l.alive and l.plus < 0 and l.scan_done and not l.q:
                l.alive = False
        stats["b_it"] += b_max; stats["b_ln"] += b_sum
        cost += COST["B"] * b_max
        stats["rounds"] += 1; stats["alive"] += sum(l.alive for l in lanes)
        # phase C
        cut = [l for l in lanes if l.plus >= 0]
        if cut:
            for l in cut:
                l.prr = l.pa[l.plus]
            stats["c_it"] += 1; stats["c_ln"] += len(cut)
            cost += COST["C"]
        cost += COST["ROUND"]
    cost += COST["EMIT"]
    return cost


def sim_refill_kernel(polys, QCAP=4, EV=8, THR=6, CHUNK=256, stats=None, merged=False, park=80.0, init=150.0, emit2=4000.0):
    """Lane refill (k_clip_refill): a warp owns a chunk of slots; lanes whose polygon is finished park it and take the next
    slot as soon as THR lanes are idle.  merged=True replaces phases A/B by a direct walk of the candidate stream."""
    cost = 0.0
    for c0 in range(0, len(polys), CHUNK):
        pool = [Lane(p[2], p[3]) for p in polys[c0:c0 + CHUNK]]
        pool.reverse()
        lanes = [None] * 32
        while True:
            idle = [k for k in range(32) if lanes[k] is None or not lanes[k].alive]
            if pool and (len(idle) >= THR or len(idle) == 32):
                for k in idle:
                    if pool:
                        lanes[k] = pool.pop()
                cost += park + init
            live = [l for l in lanes if l is not None and l.alive]
            if not live:
                if not pool:
                    break
                continue
            # rounds until enough lanes are free again
            while True:
                live = [l for l in lanes if l is not None and l.alive]
                if not live:
                    break
                if merged:
                    cost += round_merged(live, stats)
                else:
                    cost += round_abc(live, QCAP, EV, stats)
                nalive = sum(1 for l in lanes if l is not None and l.alive)
                if pool and 32 - nalive >= THR:
                    break
        cost += emit2 * (min(CHUNK, len(polys) - c0) / 32.0)
    return cost


def round_abc(lanes, QCAP, EV, stats):
    cost = 0.0
    for _ in range(EV):
        sc = [l for l in lanes if l.alive and not l.scan_done and len(l.q) < QCAP]
        if not sc:
            break
        for l in sc:
            l.scan_event()
        stats["a_it"] += 1; stats["a_ln"] += len(sc)
        cost += COST["A"]
    b_max = b_sum = 0
    for l in lanes:
        l.plus = -1
        nb = 0
        while l.alive and l.q and l.plus < 0:
            j = l.q.pop(0)
            nb += 1
            k = l.kind[j]
            if k == 1 or l.d2[j] > l.prr:
                continue
            if k == 4:
                l.plus = j
        b_max = max(b_max, nb); b_sum += nb
        if l.alive and l.plus < 0 and l.scan_done and not l.q:
            l.alive = False
    stats["b_it"] += b_max; stats["b_ln"] += b_sum
    cost += COST["B"] * b_max
    stats["rounds"] += 1; stats["alive"] += sum(l.alive for l in lanes)
    cut = [l for l in lanes if l.plus >= 0]
    if cut:
        for l in cut:
            l.prr = l.pa[l.plus]
        stats["c_it"] += 1; stats["c_ln"] += len(cut)
        cost += COST["C"]
    return cost + COST["ROUND"]


def round_merged(lanes, stats, STEP=45.0, ADV=40.0, CLS=120.0, MAXIT=64):
    """No queue, no scan-ahead: every lane walks its candidate stream (exact distance filter with the current radius)
    until one candidate has a vertex outside; then the cut.  One loop iteration = one stream step per searching lane."""
    cost = 0.0
    for l in lanes:
        l.plus = -1
    it = 0
    while it < MAXIT:
        srch = [l for l in lanes if l.alive and l.plus < 0]
        if not srch:
            break
        it += 1
        any_adv = any_cls = False
        for l in srch:
            k = l.kind[l.i]
            if k == 0 or k >= 5:  # node entry
                any_adv = True
                if k >= 5:
                    l.alive = False
                    continue
                l.i += 1
                k = l.kind[l.i]
                if k == 0 or k >= 5:
                    continue
            j = l.i
            l.i += 1
            if k == 1 or l.d2[j] > l.prr:
                continue
            any_cls = True
            if k == 4:
                l.plus = j
        stats["b_it"] += 1; stats["b_ln"] += len(srch)
        cost += STEP + (ADV if any_adv else 0.0) + (CLS if any_cls else 0.0)
    stats["rounds"] += 1; stats["alive"] += sum(l.alive for l in lanes)
    cut = [l for l in lanes if l.plus >= 0]
    if cut:
        for l in cut:
            l.prr = l.pa[l.plus]
        stats["c_it"] += 1; stats["c_ln"] += len(cut)
        cost += COST["C"]
    return cost + COST["ROUND"]


def report(name, stats, ntiles, cost):
    s = stats
    print(f"[{name}] tiles {ntiles} rounds/tile {s['rounds']/ntiles:.2f} alive/round {s['alive']/max(1,s['rounds']):.2f} | "
          f"A events/tile {s['a_it']/ntiles:.1f} lanes {s['a_ln']/max(1,s['a_it']):.2f} | B iters/tile {s['b_it']/ntiles:.1f} lanes "
          f"{s['b_ln']/max(1,s['b_it']):.2f} | C cuts/tile {s['c_it']/ntiles:.1f} lanes {s['c_ln']/max(1,s['c_it']):.2f} | "
          f"warp instr/cell {cost/ntiles/32:.0f}")


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else "/tmp/trace.bin"
    what = sys.argv[2] if len(sys.argv) > 2 else "tile"
    polys = load(path)
    ntiles = len(polys) // 32
    evs = np.array([len(p[2]) for p in polys])
    cuts = np.array([(p[2]["kind"] == 4).sum() for p in polys])
    passes = np.array([((p[2]["kind"] == 3) | (p[2]["kind"] == 4)).sum() for p in polys])
    cands = np.array([((p[2]["kind"] >= 1) & (p[2]["kind"] <= 4)).sum() for p in polys])
    print(f"polygons {len(polys)}: candidates {cands.mean():.1f}, pass the exact filter {passes.mean():.1f}, cuts {cuts.mean():.2f} "
          f"(max {cuts.max()}), events {evs.mean():.1f}")
    stats = dict(rounds=0, alive=0, a_it=0, a_ln=0, b_it=0, b_ln=0, c_it=0, c_ln=0)
    total = 0.0
    if what == "tile":
        for ti in range(ntiles):
            lanes = [Lane(p[2], p[3]) for p in polys[32 * ti: 32 * ti + 32]]
            total += sim_tile_policy(lanes, stats=stats)
        report("tile (production: QLOW=1 EV=4 BMAX=2)", stats, ntiles, total)
    elif what == "tile_v1":
        for ti in range(ntiles):
            lanes = [Lane(p[2], p[3]) for p in polys[32 * ti: 32 * ti + 32]]
            total += sim_tile_kernel(lanes, stats=stats)
        report("tile_v1 (scan whenever there is room, EV=8, pop until a cut)", stats, ntiles, total)
    elif what == "tile_merged":
        for ti in range(ntiles):
            lanes = [Lane(p[2], p[3]) for p in polys[32 * ti: 32 * ti + 32]]
            c = 0.0
            while any(l.alive for l in lanes):
                c += round_merged([l for l in lanes if l.alive], stats)
            total += c + COST["EMIT"]
        report("tile_merged", stats, ntiles, total)
    elif what.startswith("refill"):
        kw = {}
        for a in sys.argv[3:]:
            k, v = a.split("=")
            kw[k] = float(v) if "." in v else int(v)
        total = sim_refill_kernel(polys[: ntiles * 32], stats=stats, merged=what.endswith("merged"), **kw)
        report(what + str(kw), stats, ntiles, total)


if __name__ == "__main__":
    main()
