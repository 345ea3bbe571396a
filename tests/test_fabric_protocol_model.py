"""Race / deadlock check of the multi-GPU fabric's barrier protocol on a model (CPU): every rank is a coroutine that
follows the sequence of publications, inbox reads and fabric barriers of k_team_run (csrc/team_run.cuh: F1 after the
Nesterov phase, split F2 "inbox consumed", F3 at the leader's turn; one barrier for plain RBCD; two per tick under the
parallel schedule), a random scheduler interleaves them, and two invariants are asserted:

  * freshness  -- a robot assembling G in iteration k reads, for every neighbour, exactly the poses that neighbour
                  published for iteration k (not the previous ones, not the next ones);
  * no WAR     -- nobody stores into an inbox while its owner is still reading it.

The negative controls drop one barrier each and must trip an invariant, so the model has teeth.  (The device
implementation of arrive / wait -- flags in peer memory, release / acquire at system scope -- is exercised on real
hardware by tests/test_gpu_fabric*.py; this file checks the protocol those primitives are used in.)"""
import random

import pytest


class Violation(AssertionError):
    pass


class Fabric:
    def __init__(self, world, robots, neighbors, rank_of):
        self.world, self.robots, self.nbrs, self.rank_of = world, robots, neighbors, rank_of
        self.flags = [[0] * world for _ in range(world)]       # flags[dst][src]
        self.inbox = {}                                         # (dst robot, src robot, kind) -> version
        self.reading = {}                                       # dst robot -> expected {(src, kind): version} while reading

    def publish(self, src, kind, version):
        for dst in self.nbrs[src]:
            want = self.reading.get(dst)
            if want is not None and (src, kind) in want and want[(src, kind)] != version:
                raise Violation(f"robot {src} overwrites {kind} in robot {dst}'s inbox while it is being read")
            self.inbox[(dst, src, kind)] = version

    def begin_read(self, dst, kind, version):
        want = {(src, kind): version for src in self.nbrs[dst]}
        for (src, k), v in want.items():
            got = self.inbox.get((dst, src, k))
            if got != v:
                raise Violation(f"robot {dst} reads {k} of robot {src}: version {got}, expected {v}")
        self.reading[dst] = want

    def end_read(self, dst):
        self.reading.pop(dst, None)

    def arrive(self, rank, seq):
        for dst in range(self.world):
            if dst != rank:
                self.flags[dst][rank] = seq

    def ready(self, rank, upto):
        return all(self.flags[rank][s] >= upto for s in range(self.world) if s != rank)


def rank_program(F, rank, iters, mode, restart_every=0, drop=None):
    """Generator: yields ("wait", seq) at fabric waits and ("step",) inside a read window (so others can interleave)."""
    local = [a for a in range(F.robots) if F.rank_of[a] == rank]
    N, seq, wait_to = F.robots, 0, 0
    for k in range(iters):
        sel = k % N
        restart = restart_every and (k + 1) % restart_every == 0
        if mode == "parallel":
            for a in local:                      # gradient of every local robot against the previous tick's poses
                F.begin_read(a, "reg", k)
                yield ("step",)
                F.end_read(a)
            seq += 1
            if drop != "consumed":
                F.arrive(rank, seq)
                yield ("wait", seq)
            else:
                F.arrive(rank, seq)
            for a in local:
                F.publish(a, "reg", k + 1)
            seq += 1
            F.arrive(rank, seq)
            yield ("wait", seq)
            continue
        if mode == "accel":
            yield ("wait", wait_to)              # peers finished reading their inboxes (F2 of the previous iteration)
            for a in local:                      # Nesterov phase
                if restart:
                    if a != sel:
                        F.publish(a, "aux", k)
                        F.publish(a, "reg", k)
                else:
                    F.publish(a, "aux", k)
                    if a != sel:
                        F.publish(a, "reg", k)
            seq += 1
            F.arrive(rank, seq)
            yield ("wait", seq)                  # F1
        else:                                    # plain RBCD: X+ of the previous iteration has arrived
            seq += 1
            F.arrive(rank, seq)
            yield ("wait", seq)
        if sel in local:
            kind = "reg" if (mode != "accel" or restart) else "aux"
            if mode == "accel":
                want = k
            else:
                want = None                      # plain RBCD: every neighbour's latest X+ (checked below)
            if want is not None:
                F.begin_read(sel, kind, want)
            else:
                F.reading[sel] = {(src, "reg"): F.inbox.get((sel, src, "reg")) for src in F.nbrs[sel]}
                for src in F.nbrs[sel]:          # latest step of src happened at the last iteration it was selected
                    last = max([j for j in range(k) if j % N == src], default=None)
                    exp = "init" if last is None else last + 0.5
                    if F.inbox.get((sel, src, "reg")) != exp:
                        raise Violation(f"robot {sel} reads X of robot {src}: {F.inbox.get((sel, src, 'reg'))}, expected {exp}")
            yield ("step",)
            F.end_read(sel)
        if mode == "accel" and drop != "consumed":
            seq += 1
            F.arrive(rank, seq)                  # F2: my inbox is consumed (or I had nothing to read)
            wait_to = seq
        if sel in local:
            F.publish(sel, "reg", k + 0.5 if mode != "accel" else k)   # X+ (accel: re-published next Nesterov phase)
            if mode == "accel" and restart:
                F.publish(sel, "aux", k)
        if sel == 0 and drop != "leader":        # leader's turn: ready bits
            seq += 1
            F.arrive(rank, seq)
            yield ("wait", seq)
            wait_to = max(wait_to, seq)


def simulate(world, robots, mode, iters, seed, restart_every=0, drop=None, ring=True):
    rng = random.Random(seed)
    nbrs = {a: sorted({(a - 1) % robots, (a + 1) % robots} - {a}) if ring else
            sorted(b for b in (a - 1, a + 1) if 0 <= b < robots) for a in range(robots)}
    per = robots // world
    rank_of = {a: min(a // per, world - 1) for a in range(robots)}
    F = Fabric(world, robots, nbrs, rank_of)
    for a in range(robots):                      # INITIALIZE: everybody has everybody's poses
        for b in nbrs[a]:
            F.inbox[(a, b, "reg")] = "init" if mode == "plain" else 0
            F.inbox[(a, b, "aux")] = 0
    if mode == "parallel":
        for a in range(robots):
            for b in nbrs[a]:
                F.inbox[(a, b, "reg")] = 0
    progs = [rank_program(F, r, iters, mode, restart_every, drop) for r in range(world)]
    state = [next(p) for p in progs]             # each rank runs to its first yield
    alive = set(range(world))
    for _ in range(200000):
        runnable = [r for r in alive if state[r][0] == "step" or F.ready(r, state[r][1])]
        if not alive:
            return True
        if not runnable:
            raise Violation(f"deadlock: {[(r, state[r]) for r in alive]}")
        r = rng.choice(runnable)
        try:
            state[r] = next(progs[r])
        except StopIteration:
            alive.discard(r)
    raise Violation("did not finish")


@pytest.mark.parametrize("mode", ["accel", "plain", "parallel"])
@pytest.mark.parametrize("world,robots", [(2, 8), (4, 8), (8, 8), (3, 6)])
def test_protocol_is_race_and_deadlock_free(mode, world, robots):
    for seed in range(40):
        assert simulate(world, robots, mode, iters=3 * robots + 2, seed=seed, restart_every=5 if mode == "accel" else 0)


def test_accel_without_the_consumed_barrier_races():
    """Negative control: without F2 a fast rank's next Nesterov phase overwrites an inbox that is still being read."""
    with pytest.raises(Violation):
        for seed in range(200):
            simulate(4, 8, "accel", iters=30, seed=seed, drop="consumed")


def test_parallel_without_the_consumed_barrier_races():
    with pytest.raises(Violation):
        for seed in range(200):
            simulate(4, 8, "parallel", iters=30, seed=seed, drop="consumed")
