"""Race check of the per-robot driver's three-barrier step on a model (CPU).

dpgo_b200_sync_driver_run / _run_shm (csrc/sync_driver.cu) replay PGOAgentROS's synchronous call sequence with one
thread per robot.  Round 2 removed the barrier at the end of a step: mailboxes and statuses are double-buffered by
the parity of the step, and a robot may pack the next iteration's poses (or, without acceleration, run a whole next
step up to its single barrier) while a slower one still delivers this iteration's.  Here every robot is a coroutine
that follows exactly that sequence, a random scheduler interleaves them, and the invariants are:

  * freshness  -- a delivery reads, from every mailbox, the message packed for exactly this step;
  * no tearing -- nobody packs into a mailbox (or posts into a status slot) while somebody still reads it;
  * status     -- the leader reads every robot's status of exactly this step.

Negative controls (single-buffered mailboxes / statuses without the end-of-step barrier) must trip an invariant, so
the model has teeth.  The real thing is exercised on the GPU by tests/test_gpu_parity.py::test_native_sync_driver_* and
tests/test_gpu_fabric_multi.py (iterates identical to the oracle's)."""
import random

import pytest


class Violation(AssertionError):
    pass


class Board:
    """Mailboxes out[p][a] (one per sender, read by every neighbour) and status slots, each tagged with the step it
    was written for; readers register while they read."""

    def __init__(self, n, buffers):
        self.n, self.buffers = n, buffers
        self.box = {}        # (p, sender) -> step
        self.status = {}     # (p, robot) -> step
        self.readers = {}    # ("box" | "status", p, index) -> number of readers inside

    def _slot(self, step):
        return step % self.buffers

    def write(self, kind, who, step):
        key = (kind, self._slot(step), who)
        if self.readers.get(key, 0):
            raise Violation(f"{kind} of robot {who} rewritten for step {step} while it is being read")
        (self.box if kind == "box" else self.status)[(self._slot(step), who)] = step

    def begin_read(self, kind, who, step):
        got = (self.box if kind == "box" else self.status).get((self._slot(step), who))
        if got != step:
            raise Violation(f"{kind} of robot {who}: read version {got}, expected {step}")
        key = (kind, self._slot(step), who)
        self.readers[key] = self.readers.get(key, 0) + 1

    def end_read(self, kind, who, step):
        self.readers[(kind, self._slot(step), who)] -= 1


class Barrier:
    def __init__(self, n):
        self.n, self.count, self.gen = n, 0, 0


def robot(a, n, steps, accelerated, board, bar, end_barrier):
    """One thread of the driver.  Yields ("work",) between actions and ("wait", generation) at a barrier."""

    def barrier():
        g = bar.gen
        bar.count += 1
        if bar.count == bar.n:
            bar.count = 0
            bar.gen += 1
        else:
            while bar.gen == g:
                yield ("wait",)

    def read_all(kind, sources, step):
        for s in sources:
            board.begin_read(kind, s, step)
            yield ("work",)          # another thread may run while this one copies
            board.end_read(kind, s, step)

    for s in range(steps):
        sel = s % n
        others = [b for b in range(n) if b != a]
        if accelerated:
            if a != sel:
                yield ("work",)                                    # iterate(false)
                board.write("box", a, s)                           # pack: Y of this step
                yield ("work",)
            yield from barrier()
            yield from read_all("box", [b for b in others if b != sel], s)
            yield from barrier()
        elif a != sel:
            yield ("work",)                                        # iterate(false): counts only
        if a == sel:
            yield ("work",)                                        # iterate(true)
            board.write("box", a, s)                               # pack: X+ (and Y) of this step
            yield ("work",)
        board.write("status", a, s)
        yield from barrier()
        if a != sel:
            yield from read_all("box", [sel], s)
        if a == 0:
            yield from read_all("status", [b for b in range(1, n)], s)
        if end_barrier:
            yield from barrier()


def run(n, steps, accelerated, buffers, end_barrier, seed):
    rng = random.Random(seed)
    board, bar = Board(n, buffers), Barrier(n)
    threads = {a: robot(a, n, steps, accelerated, board, bar, end_barrier) for a in range(n)}
    waiting = set()
    while threads:
        runnable = [a for a in threads if a not in waiting] or list(threads)
        a = rng.choice(runnable)
        try:
            what = next(threads[a])
        except StopIteration:
            del threads[a]
            waiting.discard(a)
            continue
        if what[0] == "wait":
            waiting.add(a)
            if all(t in waiting for t in threads):   # everybody spins: let them all look again
                waiting.clear()
        else:
            waiting.discard(a)
            waiting.clear()


@pytest.mark.parametrize("accelerated", [True, False])
@pytest.mark.parametrize("n", [2, 3, 8])
def test_three_barrier_step_with_double_buffers(n, accelerated):
    for seed in range(40):
        run(n, 4 * n + 3, accelerated, buffers=2, end_barrier=False, seed=seed)


@pytest.mark.parametrize("accelerated", [True, False])
def test_four_barrier_step_with_single_buffers(accelerated):
    """Round 1's form: single buffers are fine as long as the step ends with a barrier."""
    for seed in range(20):
        run(4, 19, accelerated, buffers=1, end_barrier=True, seed=seed)


@pytest.mark.parametrize("accelerated", [True, False])
def test_negative_control_single_buffers_without_end_barrier(accelerated):
    tripped = 0
    for seed in range(200):
        try:
            run(4, 19, accelerated, buffers=1, end_barrier=False, seed=seed)
        except Violation:
            tripped += 1
    assert tripped > 0, "the model did not notice the missing barrier"
