"""Discrete-event models of the two multi-GPU exchange protocols (csrc/lib_shard.inl, csrc/lib_dist.inl), run under
randomised, skewed timings on the CPU: no GPU needed, so the N > 1 hand-off logic is exercised by `-m "not gpu"` too.

What is modelled: per rank the in-order streams and their operations (kernels with a duration, one-warp flag waits,
flag publications into the peers' windows), the double-buffered / per-round windows with a version tag per write, and --
for the sharded form -- the fact that a kernel which fills every SM keeps other kernels of the same GPU from starting.
What is checked: nobody reads a window region while it is being rewritten or before the data it expects is there, every
rank finishes (no deadlock), and the two bugs found on the GPU (back-pressure wait inside the top-k CTAs; missing
back-pressure) are caught by the model when re-introduced."""
import heapq
import random

import pytest


class Sim:
    """ranks x streams of generator-based operations; yield ("work", dt, exclusive) | ("wait", predicate) | ("call", fn)"""

    def __init__(self, seed):
        self.rng = random.Random(seed)
        self.now = 0.0
        self.heap = []          # (time, n, stream)
        self.n = 0
        self.waiting = []       # (predicate, stream)
        self.streams = []
        self.busy_exclusive = {}  # gpu -> number of running kernels that fill the whole GPU

    def add(self, gpu, gen):
        st = {"gpu": gpu, "gen": gen, "done": False}
        self.streams.append(st)
        self._advance(st)

    def _push(self, t, st):
        self.n += 1
        heapq.heappush(self.heap, (t, self.n, st))

    def _advance(self, st, value=None):
        try:
            op = st["gen"].send(value)
        except StopIteration:
            st["done"] = True
            return
        if op[0] == "work":
            _, dt, exclusive = op
            if self.busy_exclusive.get(st["gpu"], 0) > 0:          # a GPU-filling kernel is running: this one cannot start
                self.waiting.append((lambda g=st["gpu"]: self.busy_exclusive.get(g, 0) == 0, st, op))
                return
            if exclusive:
                self.busy_exclusive[st["gpu"]] = self.busy_exclusive.get(st["gpu"], 0) + 1
                st["excl"] = True
            self._push(self.now + dt, st)
        elif op[0] == "wait":
            if op[1]():
                self._advance(st)
            else:
                self.waiting.append((op[1], st, None))
        elif op[0] == "call":
            op[1]()
            self._advance(st)

    def run(self, limit=10**6):
        steps = 0
        while True:
            progressed = True
            while progressed:                                       # release every waiter whose condition holds
                progressed = False
                for w in list(self.waiting):
                    if w[0]():
                        self.waiting.remove(w)
                        if w[2] is not None:                       # a deferred kernel start
                            st = w[1]
                            st["gen"] = _prepend(w[2], st["gen"])
                        self._advance(w[1])
                        progressed = True
            if not self.heap:
                break
            self.now, _, st = heapq.heappop(self.heap)
            if st.pop("excl", False):
                self.busy_exclusive[st["gpu"]] -= 1
            self._advance(st)
            steps += 1
            assert steps < limit
        return all(s["done"] for s in self.streams)


def _prepend(op, gen):
    def g():
        v = yield op
        while True:
            try:
                v = yield gen.send(v)
            except StopIteration:
                return
    it = g()
    return it


# ------------------------------------------------------------------------------------------------------------------
# sharded form: push into every window (final top-k epilogue), wait + merge on the group's stream, depth-2 windows
# ------------------------------------------------------------------------------------------------------------------
def run_shard_model(G, n_seq, seed, backpressure="kernel", depth=2):
    """backpressure: "kernel" = one-warp kernel before the final top-k (the shipped protocol);
    "in_cta" = waited inside the GPU-filling top-k kernel (the N=4 deadlock); "none" = no back-pressure at all."""
    sim = Sim(seed)
    rng = sim.rng
    arrive = [[[0] * G for _ in range(depth)] for _ in range(G)]      # arrive[owner][slot][src]
    done = [[0] * G for _ in range(G)]                                 # done[owner][src]
    content = [[[None] * G for _ in range(depth)] for _ in range(G)]  # window[owner][slot][src] = seq written
    reading = [[False] * depth for _ in range(G)]                      # owner is merging this slot
    pushed = [[-1] * G for _ in range(G)]
    errors = []
    skew = [rng.uniform(0.5, 3.0) for _ in range(G)]

    def main(r):
        for s in range(n_seq):
            slot = s % depth
            yield ("work", rng.uniform(5, 10) * skew[r], False)                       # the search itself
            need = s + 1 - depth
            free = lambda r=r, need=need: need <= 0 or all(done[r][p] >= need for p in range(G))   # noqa: E731
            if backpressure == "kernel":
                yield ("wait", free)
                yield ("work", 0.3, True)                                              # final top-k: fills the GPU, never waits
            elif backpressure == "in_cta":
                started = {"t": False}

                def begin(started=started, r=r):
                    started["t"] = True
                    sim.busy_exclusive[r] = sim.busy_exclusive.get(r, 0) + 1           # CTAs resident and spinning
                yield ("call", begin)
                yield ("wait", free)
                yield ("call", lambda r=r: sim.busy_exclusive.__setitem__(r, sim.busy_exclusive[r] - 1))
                yield ("work", 0.3, False)
            else:
                yield ("work", 0.3, True)

            def push(r=r, s=s, slot=slot):
                for p in range(G):
                    if reading[p][slot]:
                        errors.append(f"rank {r} overwrote slot {slot} of rank {p} during its merge (seq {s})")
                    content[p][slot][r] = s
                    arrive[p][slot][r] = s + 1
                pushed[r][r] = s
            yield ("call", push)

    def merger(r):
        for s in range(n_seq):
            slot = s % depth
            yield ("wait", lambda r=r, s=s: pushed[r][r] >= s)                         # cudaStreamWaitEvent(ev_push)
            yield ("wait", lambda r=r, s=s, slot=slot: all(arrive[r][slot][p] >= s + 1 for p in range(G)))
            yield ("call", lambda r=r, slot=slot: reading[r].__setitem__(slot, True))
            yield ("work", rng.uniform(0.2, 2.0) * skew[r], False)                     # merge kernel (needs a free SM slot)

            def finish(r=r, s=s, slot=slot):
                if any(content[r][slot][p] != s for p in range(G)):
                    errors.append(f"rank {r} merged seq {s} from slot contents {content[r][slot]}")
                reading[r][slot] = False
                for p in range(G):
                    done[p][r] = s + 1
            yield ("call", finish)

    for r in range(G):
        sim.add(r, main(r))
        sim.add(r, merger(r))
    finished = sim.run()
    return finished, errors


@pytest.mark.parametrize("G", [2, 4, 8])
def test_shard_exchange_protocol_is_safe_and_live(G):
    for seed in range(40):
        finished, errors = run_shard_model(G, n_seq=7, seed=seed)
        assert finished, f"deadlock (G={G}, seed={seed})"
        assert not errors, errors[:3]


def test_shard_model_catches_the_bugs_found_on_the_gpu():
    # (i) waiting for the done flags inside the top-k CTAs, which fill the GPU: this GPU's own merge can never start
    assert any(not run_shard_model(4, 7, seed, backpressure="in_cta")[0] for seed in range(40))
    # (ii) no back-pressure: a fast rank overwrites a slot a slow rank is still merging (or has not merged yet)
    assert any(run_shard_model(4, 7, seed, backpressure="none")[1] for seed in range(40))


# ------------------------------------------------------------------------------------------------------------------
# distributed scoring: per call one user-state broadcast, per round ids out -> score -> scores back, ONE stream per rank
# ------------------------------------------------------------------------------------------------------------------
def run_dist_model(G, n_calls, seed, wait_scores=True, rounds=5):
    sim = Sim(seed)
    rng = sim.rng
    hu_flag = [[0] * G for _ in range(G)]
    ids_flag = [[0] * G for _ in range(G)]
    sc_flag = [[0] * G for _ in range(G)]
    hu = [[None] * G for _ in range(G)]           # hu[owner][src] = call
    req = [[None] * G for _ in range(G)]          # req[owner][src] = (call, round)
    resp = [[None] * G for _ in range(G)]         # resp[src][owner] = (call, round)
    scoring = [False] * G                         # owner reads req + hu
    unbucketing = [False] * G                     # source reads resp
    errors = []
    skew = [rng.uniform(0.5, 3.0) for _ in range(G)]

    def rank(r):
        e_hu = e_ids = e_sc = 0
        for c in range(n_calls):
            yield ("work", rng.uniform(0.1, 0.5) * skew[r], False)                     # users copy, init, hoist
            e_hu += 1

            def bcast(r=r, c=c, e=e_hu):
                for p in range(G):
                    if scoring[p]:
                        errors.append(f"rank {r} rewrote its user state on rank {p} while {p} was scoring")
                    hu[p][r] = c
                    hu_flag[p][r] = e
            yield ("call", bcast)
            yield ("wait", lambda r=r, e=e_hu: all(hu_flag[r][p] >= e for p in range(G)))
            for k in range(rounds):
                yield ("work", rng.uniform(0.2, 1.0) * skew[r], False)                 # expand + filter / previous top-k
                e_ids += 1

                def bucket(r=r, c=c, k=k, e=e_ids):
                    for o in range(G):
                        if scoring[o]:
                            errors.append(f"rank {r} rewrote its request on rank {o} while {o} was scoring (call {c} round {k})")
                        req[o][r] = (c, k)
                        ids_flag[o][r] = e
                yield ("call", bucket)
                yield ("wait", lambda r=r, e=e_ids: all(ids_flag[r][p] >= e for p in range(G)))
                yield ("call", lambda r=r: scoring.__setitem__(r, True))
                yield ("work", rng.uniform(1.0, 3.0) * skew[r], True)                  # the scorer kernel fills the GPU

                def scored(r=r, c=c, k=k):
                    if any(req[r][p] != (c, k) for p in range(G)) or any(hu[r][p] != c for p in range(G)):
                        errors.append(f"rank {r} scored call {c} round {k} from requests {req[r]} / user states {hu[r]}")
                    scoring[r] = False
                yield ("call", scored)
                e_sc += 1

                def give_back(r=r, c=c, k=k, e=e_sc):
                    for src in range(G):
                        if unbucketing[src]:
                            errors.append(f"rank {r} rewrote scores on rank {src} while {src} was reading them")
                        resp[src][r] = (c, k)
                        sc_flag[src][r] = e
                yield ("call", give_back)
                if wait_scores:
                    yield ("wait", lambda r=r, e=e_sc: all(sc_flag[r][p] >= e for p in range(G)))
                yield ("call", lambda r=r: unbucketing.__setitem__(r, True))
                yield ("work", rng.uniform(0.05, 0.3) * skew[r], False)

                def unbucket(r=r, c=c, k=k):
                    if any(resp[r][o] != (c, k) for o in range(G)):
                        errors.append(f"rank {r} read scores of {resp[r]} in call {c} round {k}")
                    unbucketing[r] = False
                yield ("call", unbucket)

    for r in range(G):
        sim.add(r, rank(r))
    finished = sim.run()
    return finished, errors


@pytest.mark.parametrize("G", [2, 4, 8])
def test_distributed_scoring_protocol_is_safe_and_live(G):
    for seed in range(30):
        finished, errors = run_dist_model(G, n_calls=3, seed=seed)
        assert finished, f"deadlock (G={G}, seed={seed})"
        assert not errors, errors[:3]


def test_dist_model_catches_a_missing_wait():
    assert any(run_dist_model(4, 3, seed, wait_scores=False)[1] for seed in range(30))
