"""Schedule-fuzzing model of the int8 TEAM form's partial-sum hand-over (csrc/megakernel.cu, gemv_phase,
`ph.team`): kTeam consumer warps share a ring stage, each sums a slice of the columns of the stage's rows;
members 1..kTeam-1 write their row partials into shared-memory scratch, the whole team meets at a named
barrier (`bar.sync 2 + team, 32 * kTeam`), member 0 adds the partials in member order and runs the epilogues.

The scratch is double buffered by the team's task parity.  Why two buffers suffice and one does not: a member
that leaves the barrier of task k may race ahead to task k + 1 and write its partial while member 0 is still
reading the partials of task k -- into the OTHER buffer.  It cannot reach task k + 2 (the same buffer again)
before member 0 arrives at the barrier of task k + 1, which member 0 only does after it has finished reading
task k.  The model runs the protocol with one and with two buffers under random and adversarial schedules;
every partial carries its task number, so a read of the wrong task's value is detected.
"""
import random

import pytest


class TeamError(AssertionError):
    pass


class NamedBarrier:
    """bar.sync with a fixed arrival count: a thread leaves once the generation it arrived in is complete."""

    def __init__(self, count):
        self.count, self.arrived, self.generation = count, 0, 0

    def arrive(self):
        gen = self.generation
        self.arrived += 1
        if self.arrived == self.count:
            self.arrived = 0
            self.generation += 1
        return gen

    def passed(self, gen):
        return self.generation > gen


def member(m, team, tasks, buffers, scratch, bar, results):
    for k in range(tasks):
        buf = k % buffers
        yield "sum my slice"
        if m != 0:
            scratch[buf][m] = (k, m)           # this member's partial of task k
        yield "arrive"
        gen = bar.arrive()
        while not bar.passed(gen):
            yield "wait"
        if m == 0:
            total = []
            for other in range(1, team):        # member order, one shared-memory read at a time
                yield "read"
                got = scratch[buf][other]
                if got != (k, other):
                    raise TeamError(f"task {k}: member 0 read {got} from member {other}'s slot")
                total.append(got)
            results.append(k)


def run(team, tasks, buffers, pick):
    scratch = [[None] * team for _ in range(buffers)]
    bar = NamedBarrier(team)
    results = []
    gens = [member(m, team, tasks, buffers, scratch, bar, results) for m in range(team)]
    alive = list(range(team))
    steps = 0
    while alive:
        m = pick(alive, steps)
        try:
            next(gens[m])
        except StopIteration:
            alive.remove(m)
        steps += 1
        if steps > 200000:
            raise TeamError("no progress (deadlock)")
    assert results == list(range(tasks))


def random_pick(rng):
    return lambda alive, steps: rng.choice(alive)


def starve_member0(rng):
    # member 0 runs only when nobody else can make progress without it, or rarely: the others race ahead
    def pick(alive, steps):
        others = [m for m in alive if m != 0]
        if others and rng.random() < 0.97:
            return rng.choice(others)
        return rng.choice(alive)
    return pick


@pytest.mark.parametrize("team", [2, 4])
@pytest.mark.parametrize("seed", range(8))
def test_double_buffered_scratch_is_safe(team, seed):
    rng = random.Random(seed)
    run(team, 40, 2, random_pick(rng))
    run(team, 40, 2, starve_member0(rng))


@pytest.mark.parametrize("team", [2, 4])
def test_single_buffer_is_caught(team):
    """Negative control: with ONE scratch buffer a member that races ahead overwrites a partial member 0 has not
    read yet -- the model must see it (otherwise it could not vouch for the double-buffered form)."""
    caught = 0
    for seed in range(40):
        try:
            run(team, 40, 1, starve_member0(random.Random(seed)))
        except TeamError:
            caught += 1
    assert caught > 0
