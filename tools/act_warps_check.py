"""Host-side model of the pair GEMM's ACTIVATION WARPS (gemm_tcgen05_2cta.cu, template parameter ACT): which 16-byte vectors of a
CTA's 128 x 256 part of a tile each lane reads and writes, and the two-mbarrier hand-over between the four epilogue warps and the two
activation warps.  It checks that
  * for every tile shape (rows clipped by M, columns clipped by N, N % 4 == 0) every vector of the part is visited exactly once,
    by the whole-width fast path (a batch = 8 rows, lane owns vectors `lane` and `lane + 32` of each row) or by the clipped-column
    path (generic index, four vectors in flight), and that no access leaves the part;
  * a warp-wide access touches 32 consecutive vectors of one row (512 contiguous bytes) on the fast path;
  * the hand-over protocol -- an epilogue warp announces tile k (arrives on z_done, count 4) only after it has seen phase k - 1 of
    act_done (count 2) complete; an activation warp works on tile k after phase k of z_done -- never lets a barrier run two phases
    ahead of a waiter (parity waits alias two phases apart) and never deadlocks, under any interleaving of a randomised scheduler.
Pure Python: run it anywhere.  Used by tests/test_tile_layouts.py."""
import random

BM, BN, ACT_WARPS, BATCH, EDGE_BATCH = 128, 256, 2, 16, 4
ROWS_PER_WARP = BM // ACT_WARPS


def warp_accesses(rows, vpr):
    """[(row, vector)] per warp instruction, in issue order, for one activation warp's part (rows x vpr vectors)"""
    out = []
    total = rows * vpr
    if vpr == BN // 4:
        for rb in range(0, rows, BATCH // 2):
            for i in range(0, BATCH, 2):
                r = rb + (i >> 1)
                if r < rows:
                    out.append([(r, lane) for lane in range(32)])
                    out.append([(r, lane + 32) for lane in range(32)])
    else:
        for base in range(0, total, EDGE_BATCH * 32):
            for i in range(EDGE_BATCH):
                acc = []
                for lane in range(32):
                    e = base + i * 32 + lane
                    if e < total:
                        r = e // vpr
                        acc.append((r, e - r * vpr))
                if acc:
                    out.append(acc)
    return out


def check_decomposition():
    problems = []
    for m_left in (1, 7, 63, 64, 65, 100, 127, 128, 300):           # rows of C left from the CTA's first row
        for n_left in (4, 36, 128, 252, 256, 1000):                  # columns left from the tile's first column
            vpr = min(n_left, BN) // 4
            for aw in range(ACT_WARPS):
                r0 = aw * ROWS_PER_WARP
                rows = max(0, min(m_left - r0, ROWS_PER_WARP))
                seen = {}
                for inst in warp_accesses(rows, vpr):
                    for (r, v) in inst:
                        if not (0 <= r < rows and 0 <= v < vpr):
                            problems.append(f"out of range: rows {rows} vpr {vpr}: {(r, v)}")
                        seen[(r, v)] = seen.get((r, v), 0) + 1
                    if vpr == BN // 4:
                        rs = {r for r, _ in inst}
                        vs = sorted(v for _, v in inst)
                        if len(rs) != 1 or vs != list(range(vs[0], vs[0] + 32)):
                            problems.append(f"fast path access is not 32 consecutive vectors of one row: {inst[:3]}")
                if len(seen) != rows * vpr or any(c != 1 for c in seen.values()):
                    problems.append(f"coverage: rows {rows} vpr {vpr}: {len(seen)} of {rows * vpr} vectors, max count {max(seen.values(), default=0)}")
    return problems


class Barrier:
    """mbarrier: `count` arrivals complete a phase; wait(parity) passes once the phase of that parity has completed, i.e. while the
    barrier's current phase has the OTHER parity -- two phases later it blocks again (the aliasing the protocol must avoid)"""

    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase + 1

    def passed(self, parity):
        return (self.phase & 1) != parity


def check_protocol(tiles=9, seeds=200):
    problems = []
    for seed in range(seeds):
        rng = random.Random(seed)
        z_done, act_done = Barrier(4), Barrier(ACT_WARPS)
        # agents: ("epi", q) announce tiles 0 .. tiles-1; ("act", a) consume them.  state = next tile index and a step within it
        epi = [{"tile": 0, "step": 0} for _ in range(4)]
        act = [{"tile": 0, "step": 0, "ph": 0} for _ in range(ACT_WARPS)]
        consumed = [[] for _ in range(ACT_WARPS)]
        announced_phase_at_consume = []
        for _ in range(100000):
            runnable = []
            for q, e in enumerate(epi):
                if e["tile"] < tiles:
                    if e["step"] == 0 and e["tile"] > 0 and not act_done.passed((e["tile"] - 1) & 1):
                        continue  # waits for the activation warps to have finished the tile before
                    runnable.append(("epi", q))
            for a, w in enumerate(act):
                if w["tile"] < tiles:
                    if w["step"] == 0 and not z_done.passed(w["ph"]):
                        continue
                    runnable.append(("act", a))
            if not runnable:
                break
            kind, i = rng.choice(runnable)
            if kind == "epi":
                e = epi[i]
                z_done.arrive()  # (wait passed above) announce
                e["tile"] += 1
            else:
                w = act[i]
                if w["step"] == 0:
                    # the wait has passed: the tile it thinks it is working on must be fully announced, and not overtaken twice
                    if z_done.phase < w["tile"] + 1:
                        problems.append(f"seed {seed}: activation warp {i} started tile {w['tile']} before all four epilogue warps announced it")
                    if z_done.phase > w["tile"] + 2:
                        problems.append(f"seed {seed}: z_done ran {z_done.phase - w['tile'] - 1} phases ahead of activation warp {i}")
                    w["ph"] ^= 1
                    w["step"] = 1
                else:
                    consumed[i].append(w["tile"])
                    act_done.arrive()
                    w["tile"] += 1
                    w["step"] = 0
        if any(e["tile"] != tiles for e in epi) or any(w["tile"] != tiles for w in act):
            problems.append(f"seed {seed}: deadlock at epilogue {[e['tile'] for e in epi]} activation {[w['tile'] for w in act]}")
        if any(c != list(range(tiles)) for c in consumed):
            problems.append(f"seed {seed}: tiles consumed out of order {consumed}")
    return problems


def check():
    return check_decomposition() + check_protocol()


if __name__ == "__main__":
    p = check()
    print("\n".join(p) if p else "activation warps: decomposition and hand-over protocol OK")
