"""Discrete-event model of the barrier protocol of the tensor-core scorer (nann_b200/csrc/scorer_mlp_tc8.cuh).

The kernel's correctness hangs on a handful of mbarrier hand-offs between four agents per cluster (the MMA warp and
the epilogue warps of each of the two CTAs) whose relative timing is arbitrary: a 4-slot A ring whose positions are
consumed in lockstep order by both CTAs but produced by different agents (x slabs locally, own h1 slabs locally +
a DSMEM copy that lands two ring positions later in the peer), `a_empty` phases collected from both CTAs on two
alternating barriers per slot, parity waits that are only sound while a waiter is never two phases away, and the rule
that a slot which is the SOURCE of a DSMEM copy may be reused only after the peer consumed the copy.  Two bugs of
exactly this kind were found on the GPU by wrong scores (a parity wait that was two phases early in the first
version's split epilogue groups; reusing a copy source).

This test restates the protocol -- same ring arithmetic, same barrier indices/parities, same wait lists as the
kernel -- with mbarrier semantics (arrival counts, phase parity, `try_wait.parity` passing iff the barrier's current
phase parity differs) and runs it under randomised, deliberately skewed timings, checking at every step that
  * the MMA warp reads from a slot exactly the slab it expects (tile, unit),
  * nothing is written into a slot that is still being read, not yet consumed, or the source of a copy in flight,
  * D1 / D2 are not overwritten while an epilogue reads them, and epilogues read the tile they expect,
  * every agent finishes (no deadlock).
It models no arithmetic and needs no GPU."""
import heapq
import random

import pytest

NA, UNITS = 4, 10                      # T8_NA, T8_UNITS


class Bar:
    def __init__(self, count):
        self.count, self.arrivals, self.phase, self.waiters = count, 0, 0, []

    def ready(self, parity):           # mbarrier.try_wait.parity: true iff the phase with this parity has completed
        return (self.phase & 1) != parity


class Sim:
    def __init__(self, seed, tiles, skew):
        self.rng = random.Random(seed)
        self.t, self.seq, self.heap, self.tiles, self.skew = 0.0, 0, [], tiles, skew
        self.finished = 0
        mk = lambda n, c: [[Bar(c) for _ in range(n)] for _ in range(2)]
        self.a_full, self.a_empty = mk(NA, 1), mk(2 * NA, 2)
        self.d1_full, self.d1_empty, self.d2_full, self.d2_empty = mk(1, 1), mk(1, 1), mk(1, 1), mk(1, 1)
        self.content = [[None] * NA for _ in range(2)]       # (tile, unit) tag of the slab in a slot
        self.consumed = [[True] * NA for _ in range(2)]      # its MMAs have completed
        self.reading = [[0] * NA for _ in range(2)]          # MMAs in flight on the slot
        self.copy_src = [[0] * NA for _ in range(2)]         # outgoing DSMEM copies reading the slot
        self.incoming = [[0] * NA for _ in range(2)]         # DSMEM copies being written into the slot
        self.d_tile = [[None, None] for _ in range(2)]       # tile whose D1 / D2 is complete in TMEM
        self.d_reading = [[False, False] for _ in range(2)]  # an epilogue is reading D1 / D2
        self.d_writing = [[0, 0] for _ in range(2)]          # MMAs in flight into D1 / D2

    # ---- kernel arithmetic (scorer_mlp_tc8.cuh: t8_consumed_idx / t8_consumed_par / t8_is_own_slab)
    @staticmethod
    def consumed_bar(q):
        return (q % NA) * 2 + ((q // NA) & 1), ((q // NA) >> 1) & 1

    @staticmethod
    def is_own(q):
        return q % UNITS in (2, 3, 6, 7)

    # ---- engine
    def at(self, dt, fn):
        self.seq += 1
        heapq.heappush(self.heap, (self.t + dt, self.seq, fn))

    def arrive(self, bar):
        bar.arrivals += 1
        if bar.arrivals == bar.count:
            bar.arrivals = 0
            bar.phase += 1
            waiters, bar.waiters = bar.waiters, []
            for parity, cont in waiters:
                if bar.ready(parity):
                    self.at(0, cont)
                else:
                    bar.waiters.append((parity, cont))

    def run_agent(self, gen):
        def step():
            try:
                op = next(gen)
            except StopIteration:
                self.finished += 1
                return
            if op[0] == "sleep":
                self.at(op[1], step)
            else:                                           # ("wait", bar, parity)
                _, bar, parity = op
                if bar.ready(parity):
                    self.at(0, step)
                else:
                    bar.waiters.append((parity, step))
        self.at(0, step)

    def dur(self, lo, hi, r, kind):
        return self.rng.uniform(lo, hi) * self.skew.get((kind, r), 1.0)

    # ---- agents
    def mma(self, r):
        a_slot = a_par = ae_sel = 0
        last_done = 0.0
        for i in range(self.tiles):
            for u in range(UNITS):
                if u == 0 and i > 0:
                    yield ("wait", self.d1_empty[r][0], (i - 1) & 1)
                if u == 2 and i > 0:
                    yield ("wait", self.d2_empty[r][0], (i - 1) & 1)
                yield ("wait", self.a_full[r][a_slot], a_par)
                d = 0 if u < 2 else 1
                assert self.content[r][a_slot] == (i, u), f"cta {r}: unit ({i},{u}) found {self.content[r][a_slot]} in slot {a_slot}"
                assert not self.consumed[r][a_slot] and self.incoming[r][a_slot] == 0
                assert not self.d_reading[r][d], f"cta {r}: MMA into D{d + 1} of tile {i} while an epilogue reads it"
                self.reading[r][a_slot] += 1
                self.d_writing[r][d] += 1
                t_exec = self.dur(1300, 1900, r, "mma")
                done = max(self.t, last_done) + t_exec
                last_done = done

                def complete(slot=a_slot, sel=ae_sel, i=i, u=u, d=d):
                    self.reading[r][slot] -= 1
                    self.consumed[r][slot] = True
                    self.d_writing[r][d] -= 1
                    for c in (0, 1):                                     # tcgen05.commit ... multicast, mask 0b11
                        self.arrive(self.a_empty[c][slot * 2 + sel])
                    if u == 1:
                        self.d_tile[r][0] = i
                        self.arrive(self.d1_full[r][0])
                    if u == UNITS - 1:
                        self.d_tile[r][1] = i
                        self.arrive(self.d2_full[r][0])
                self.at(done - self.t, complete)
                yield ("sleep", t_exec * self.rng.uniform(0.5, 0.95))    # the issue loop blocks on the ~2-deep MMA queue
                a_slot += 1
                if a_slot == NA:
                    a_slot, a_par, ae_sel = 0, a_par ^ 1, ae_sel ^ 1

    def write_slot(self, r, slot, tag):
        assert self.reading[r][slot] == 0, f"cta {r}: slot {slot} written while MMAs read it ({tag})"
        assert self.consumed[r][slot], f"cta {r}: slot {slot} overwritten before {self.content[r][slot]} was consumed ({tag})"
        assert self.copy_src[r][slot] == 0, f"cta {r}: slot {slot} overwritten while it is the source of a DSMEM copy ({tag})"
        assert self.incoming[r][slot] == 0
        self.content[r][slot], self.consumed[r][slot] = tag, False

    def x_write(self, r, j):
        for half in (0, 1):
            nn = j * UNITS + half
            if nn >= NA:
                m = nn - NA
                q = m + 2 if self.is_own(m) else m
                idx, par = self.consumed_bar(q)
                yield ("wait", self.a_empty[r][idx], par)
            yield ("sleep", self.dur(100, 1500, r, "xw"))
            self.write_slot(r, nn % NA, (j, half))
            self.arrive(self.a_full[r][nn % NA])

    def epilogue(self, r):
        peer = r ^ 1
        yield from self.x_write(r, 0)
        for i in range(self.tiles):
            yield ("wait", self.d1_full[r][0], i & 1)
            assert self.d_tile[r][0] == i and self.d_writing[r][0] == 0
            self.d_reading[r][0] = True
            for js in range(4):
                n = i * UNITS + 2 + (js >> 1) * 4 + (js & 1)
                nr = n + 2
                for q in ((n - NA) if n >= NA else None, (nr - NA) if nr >= NA else None):
                    if q is not None:
                        idx, par = self.consumed_bar(q)
                        yield ("wait", self.a_empty[r][idx], par)
                if js == 3:                                              # last tcgen05.ld of D1 has landed
                    assert self.d_tile[r][0] == i
                    self.d_reading[r][0] = False
                    self.arrive(self.d1_empty[r][0])
                yield ("sleep", self.dur(300, 1500, r, "epi1"))
                slot, rslot = n % NA, nr % NA
                self.write_slot(r, slot, (i, n - i * UNITS))
                self.arrive(self.a_full[r][slot])
                # DSMEM copy: reads the local slot, writes the peer's slot of ring position nr
                assert self.reading[peer][rslot] == 0 and self.consumed[peer][rslot] and self.copy_src[peer][rslot] == 0, \
                    f"cta {r}: copy into the peer's slot {rslot} which still holds {self.content[peer][rslot]}"
                self.copy_src[r][slot] += 1
                self.incoming[peer][rslot] += 1

                def landed(slot=slot, rslot=rslot, tag=(i, nr - i * UNITS)):
                    self.copy_src[r][slot] -= 1
                    self.incoming[peer][rslot] -= 1
                    assert self.reading[peer][rslot] == 0
                    self.content[peer][rslot], self.consumed[peer][rslot] = tag, False
                    self.arrive(self.a_full[peer][rslot])               # complete_tx on the peer's barrier
                self.at(self.dur(1200, 6000, r, "copy"), landed)
            if i + 1 < self.tiles:
                yield ("sleep", self.dur(0, 800, r, "gather"))
                yield from self.x_write(r, i + 1)
            yield ("wait", self.d2_full[r][0], i & 1)
            assert self.d_tile[r][1] == i and self.d_writing[r][1] == 0
            self.d_reading[r][1] = True
            yield ("sleep", self.dur(400, 2500, r, "epi2"))
            self.d_reading[r][1] = False
            self.arrive(self.d2_empty[r][0])

    def run(self):
        for r in (0, 1):
            self.run_agent(self.mma(r))
            self.run_agent(self.epilogue(r))
        steps = 0
        while self.heap:
            self.t, _, fn = heapq.heappop(self.heap)
            fn()
            steps += 1
            assert steps < 2_000_000
        assert self.finished == 4, f"deadlock: {self.finished} of 4 agents finished at t={self.t:.0f}"


SKEWS = [
    {},                                                       # balanced
    {("mma", 0): 0.3},                                        # CTA 0's tensor core far ahead
    {("mma", 1): 3.0},                                        # CTA 1's far behind
    {("epi1", 0): 4.0, ("xw", 1): 5.0},                       # slow epilogues, asymmetric
    {("copy", 0): 6.0},                                       # DSMEM very slow in one direction
    {("copy", 0): 0.05, ("copy", 1): 0.05, ("mma", 0): 0.2, ("mma", 1): 0.2},   # everything but the epilogues is instant
    {("epi2", 1): 8.0, ("mma", 0): 0.5},
]


@pytest.mark.parametrize("skew", range(len(SKEWS)))
def test_protocol_holds_under_skewed_timings(skew):
    for seed in range(40):
        Sim(seed * 7 + skew, tiles=7, skew=SKEWS[skew]).run()


def test_model_detects_broken_protocols():
    """The model is only worth something if it fails when the protocol is broken: (1) reusing the slot of an own slab
    as soon as position n - NA was consumed, while the DSMEM copy out of it may still be in flight -- the bug that
    corrupted ~1 row in 10^4 on the GPU; (2) not waiting for the PEER's slot (position n + 2 - NA) before copying."""
    class ReuseCopySourceEarly(Sim):
        @staticmethod
        def is_own(q):
            return False

    def fails(make):
        for skew in SKEWS:
            for seed in range(25):
                try:
                    make(seed, skew).run()
                except AssertionError:
                    return True
        return False

    assert fails(lambda seed, skew: ReuseCopySourceEarly(seed, tiles=6, skew=skew))

    def ignore_peer(seed, skew):
        s = Sim(seed, tiles=6, skew=skew)
        real = s.consumed_bar
        calls = {"n": 0}

        def every_other(q):                       # epi1 asks for (n - NA) then (nr - NA): answer the second like the first
            calls["n"] += 1
            return real(q - 2) if (calls["n"] % 2 == 0 and q >= 2 + NA) else real(q)
        s.consumed_bar = every_other
        return s
    assert fails(ignore_peer)
