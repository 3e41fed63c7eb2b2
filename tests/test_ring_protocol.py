"""Schedule-fuzzing model of the megakernel's shared-memory weight ring (csrc/megakernel.cu: one
producer warp, 8 consumer warps, `num_stages` stages, a `full` and an `empty` mbarrier per stage).

An mbarrier only exposes the PARITY of its phase: `try_wait.parity p` is true while the most recently
completed phase has parity p.  A waiter that falls two phases behind therefore sees a stale answer
("phase aliasing").  The ring is safe because of one rule: EVERY consumer warp waits on EVERY fill in
order and arrives on its `empty` barrier (arrival count = number of consumer warps), even for stages
whose rows belong to other warps.  Then the producer cannot start fill n + S of a stage before every
warp has passed fill n, so no waiter is ever more than one phase behind.

The model runs that rule -- and the tempting optimisation that broke an earlier version of the kernel
(warps skip the stages they own no rows of, the owner alone arrives) -- under random and adversarial
schedules.  Data carries its fill number, so reading a stage that was refilled early, or refilling a
stage that is still being read, is detected, as is a warp that hangs on an aliased phase.
"""
import random

import pytest


class RingError(AssertionError):
    pass


class MBarrier:
    """Phase-counting barrier; waiters may only ask about the parity of the last completed phase."""

    def __init__(self, arrivals):
        self.arrivals, self.pending, self.completed = arrivals, arrivals, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.completed += 1
            self.pending = self.arrivals

    def test_wait(self, parity):
        # phase k (k = 0, 1, ...) has parity k & 1; before anything completed, the "previous" phase has
        # parity 1 -- which is what lets the producer's first wait on `empty` fall through
        last = (self.completed - 1) & 1
        return last == parity


def producer(S, fills, full, empty, stage_data, reading):
    slot, parity = 0, 0
    for n in range(fills):
        while not empty[slot].test_wait(parity ^ 1):
            yield "wait"
        if reading[slot]:
            raise RingError(f"fill {n} overwrites stage {slot} while warps {sorted(reading[slot])} still read it")
        yield "copy"
        stage_data[slot] = n  # the bulk copy lands, then completes the transaction on `full`
        full[slot].arrive()
        slot += 1
        if slot == S:
            slot, parity = 0, parity ^ 1


def consumer(w, W, S, fills, full, empty, stage_data, reading, all_wait):
    slot, parity = 0, 0
    for n in range(fills):
        owner = n % W  # (the kernel spreads a stage's rows over warps; one owner is enough for the model)
        if all_wait or owner == w:
            spins = 0
            while not full[slot].test_wait(parity):
                spins += 1
                yield "wait"
            reading[slot].add(w)
            yield "read"
            if stage_data[slot] != n:
                raise RingError(f"warp {w} expected fill {n} in stage {slot}, found fill {stage_data[slot]}")
            yield "read"
            if stage_data[slot] != n:
                raise RingError(f"warp {w}: stage {slot} was refilled (fill {stage_data[slot]}) under fill {n}")
            reading[slot].discard(w)
            if all_wait or owner == w:
                empty[slot].arrive()
        slot += 1
        if slot == S:
            slot, parity = 0, parity ^ 1


def run(W, S, fills, all_wait, seed, bias):
    rng = random.Random(seed)
    full = [MBarrier(1) for _ in range(S)]
    empty = [MBarrier(W if all_wait else 1) for _ in range(S)]
    stage_data = [None] * S
    reading = [set() for _ in range(S)]
    procs = {"p": producer(S, fills, full, empty, stage_data, reading)}
    for w in range(W):
        procs[w] = consumer(w, W, S, fills, full, empty, stage_data, reading, all_wait)
    fast = rng.choice(list(procs))
    slow = rng.choice([k for k in procs if k != fast])
    idle = 0
    while procs:
        keys = list(procs)
        if rng.random() < bias and fast in procs:
            k = fast
        else:
            k = rng.choice([x for x in keys if x != slow] or keys) if rng.random() < bias else rng.choice(keys)
        try:
            op = next(procs[k])
        except StopIteration:
            del procs[k]
            idle = 0
            continue
        idle = idle + 1 if op == "wait" else 0
        if idle > 200000:
            raise RingError("every remaining warp waits forever (a waiter missed its phase)")


@pytest.mark.parametrize("W,S", [(8, 6), (8, 2), (4, 3), (3, 16)])
def test_every_warp_waits_on_every_fill_is_safe(W, S):
    for seed in range(10):
        for bias in (0.0, 0.8, 0.97):
            run(W, S, fills=10 * S + 3, all_wait=True, seed=seed, bias=bias)


def test_skipping_stages_you_do_not_own_is_not():
    """The shortcut (only the owner waits and arrives) lets a warp fall two phases behind on a stage it
    skipped: it then reads an aliased parity -- stale data, an early refill, or a hang."""
    caught = 0
    for seed in range(30):
        for bias in (0.0, 0.8, 0.97):
            try:
                run(8, 2, fills=64, all_wait=False, seed=seed, bias=bias)
            except RingError:
                caught += 1
    assert caught > 0
