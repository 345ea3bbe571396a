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
        self.prog = [[0] * world for _ in range(world)]        # prog[dst][src]: point-to-point progress words
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

    # point-to-point progress words (round 2): a rank tells ONLY the ranks that host a neighbour of one of its robots
    def post(self, rank, to_ranks, value):
        for dst in to_ranks:
            self.prog[dst][rank] = value

    def ready_p(self, rank, from_ranks, value):
        return all(self.prog[rank][s] >= value for s in from_ranks)

    def nbr_ranks(self, robots_):
        out = set()
        for a in robots_:
            for b in self.nbrs[a]:
                out.add(self.rank_of[b])
        return out


def rank_program_p2p(F, rank, iters, mode, restart_every=0, drop=None):
    """The round-2 protocol of k_team_run (csrc/team_run.cuh): no all-to-all barrier on the critical path.  Step k
    (1-based, monotone across launches) of rank q posts, into the windows of its NEIGHBOUR ranks only,
        2k      "my Nesterov-phase publications of step k have landed"
        2k + 1  "my inbox of step k is consumed"   (posted at once by a rank that does not hold the selected robot;
                                                    plain RBCD: "my step k -- X+ included -- is over")
    and waits  * before it stores into a neighbour's inbox in step k:  2(k-1)+1 from its neighbour ranks,
               * before the selected robot reads its inbox:            2k (plain RBCD: 2(k-1)+1) from that robot's
                                                                       neighbour ranks.
    The leader's turn keeps the global barrier (ready bits of every rank)."""
    local = [a for a in range(F.robots) if F.rank_of[a] == rank]
    mine = F.nbr_ranks(local) - {rank}
    N, seq = F.robots, 0
    for it in range(iters):
        k = it + 1
        sel = it % N
        restart = restart_every and (it + 1) % restart_every == 0
        if mode == "accel":
            if k > 1 and drop != "consumed":
                yield ("waitp", mine, 2 * (k - 1) + 1)
            for a in local:                      # Nesterov phase
                if restart:
                    if a != sel:
                        F.publish(a, "aux", it)
                        F.publish(a, "reg", it)
                else:
                    F.publish(a, "aux", it)
                    if a != sel:
                        F.publish(a, "reg", it)
            F.post(rank, mine, 2 * k if sel in local else 2 * k + 1)
            if sel in local:
                if drop != "gate":
                    yield ("waitp", F.nbr_ranks([sel]) - {rank}, 2 * k)
                F.begin_read(sel, "reg" if restart else "aux", it)
                yield ("step",)
                F.end_read(sel)
                F.post(rank, mine, 2 * k + 1)
                F.publish(sel, "reg", it)
                if restart:
                    F.publish(sel, "aux", it)
        else:                                    # plain RBCD
            if sel in local:
                if k > 1 and drop != "gate":
                    yield ("waitp", F.nbr_ranks([sel]) - {rank}, 2 * (k - 1) + 1)
                F.reading[sel] = {(src, "reg"): F.inbox.get((sel, src, "reg")) for src in F.nbrs[sel]}
                for src in F.nbrs[sel]:
                    last = max([j for j in range(it) if j % N == src], default=None)
                    exp = "init" if last is None else last + 0.5
                    if F.inbox.get((sel, src, "reg")) != exp:
                        raise Violation(f"robot {sel} reads X of robot {src}: {F.inbox.get((sel, src, 'reg'))}, expected {exp}")
                yield ("step",)
                F.end_read(sel)
                F.publish(sel, "reg", it + 0.5)
            F.post(rank, mine, 2 * k + 1)
            yield ("step",)
        if sel == 0:                             # leader's turn: ready bits of every rank (global)
            seq += 1
            F.arrive(rank, seq)
            yield ("wait", seq)


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


def simulate(world, robots, mode, iters, seed, restart_every=0, drop=None, ring=True, p2p=False):
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
    prog_fn = rank_program_p2p if p2p else rank_program
    progs = [prog_fn(F, r, iters, mode, restart_every, drop) for r in range(world)]
    state = [next(p) for p in progs]             # each rank runs to its first yield
    alive = set(range(world))
    for _ in range(200000):
        runnable = [r for r in alive if state[r][0] == "step" or
                    (state[r][0] == "wait" and F.ready(r, state[r][1])) or
                    (state[r][0] == "waitp" and F.ready_p(r, state[r][1], state[r][2]))]
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


@pytest.mark.parametrize("mode", ["accel", "plain"])
@pytest.mark.parametrize("ring", [True, False])
@pytest.mark.parametrize("world,robots", [(2, 8), (4, 8), (8, 8), (3, 6)])
def test_point_to_point_protocol_is_race_and_deadlock_free(mode, world, robots, ring):
    """Round 2: neighbour-only progress words instead of the all-to-all barriers (chain and ring neighbour graphs)."""
    for seed in range(60):
        assert simulate(world, robots, mode, iters=3 * robots + 2, seed=seed, restart_every=5 if mode == "accel" else 0,
                        ring=ring, p2p=True)


@pytest.mark.parametrize("mode,drop", [("accel", "consumed"), ("accel", "gate"), ("plain", "gate")])
def test_point_to_point_protocol_negative_controls(mode, drop):
    """Dropping the write-after-read gate or the selected robot's read gate must trip an invariant."""
    with pytest.raises(Violation):
        for seed in range(300):
            simulate(4, 8, mode, iters=30, seed=seed, restart_every=5 if mode == "accel" else 0, drop=drop, p2p=True)
