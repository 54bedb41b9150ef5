"""Executable model of the row hand-over between two chained layers (reve_b200/csrc/conv_umma.cu,
conv3x3_chain_kernel): the protocol, not the CUDA code.  One sender CTA (two epilogue groups, their staging buffers and
couriers) feeds one receiver CTA (loader thread, 4-stage A ring, MMA issuer) through two scratch rings of S slots guarded
by the `published` / `consumed` counters.  Every actor advances under a random scheduler and every asynchronous effect
(a TMA store landing in the ring, a TMA load landing in the A ring) takes a random time, so the model explores
interleavings the GPU only produces rarely.  Checked: every row arrives exactly once, in order, with the right content
(no slot is overwritten before it was read, none is read before it was written and announced), and nothing deadlocks --
for several ring depths, announcement cadences and stream lengths, including the cadence the kernel ships with (a
gpu-scope release every second row, never while a row is waiting, always before blocking on `consumed`).

The model also shows what the fix of round 1 was about: with `announce_before_landing=True` (a flag that may overtake
the data, which is what a relaxed store after cp.async.bulk.wait_group allowed) the content check fails.
"""
import random

import pytest

STAGES = 4          # A ring of the receiver
RETIRE_LAG = 2      # the loader reports a slot as consumed two steps after it issued the load


def interleave(n0, n1):
    """Order in which the receiver walks the rows of its two streams (Sequencer in the kernel)."""
    out, k = [], 1
    while k <= n0 or k <= n1:
        if k <= n0:
            out.append((0, k))
        if k <= n1:
            out.append((1, k))
        k += 1
    return out


def simulate(n_rows, slots, publish_every, seed, announce_before_landing=False, max_ticks=2_000_000):
    rng = random.Random(seed)
    now = 0
    ring = [[None] * slots for _ in range(2)]          # scratch ring content: (stream, row) or None
    pub = [0, 0]                                       # rows announced (1-based count)
    cons = [0, 0]                                      # rows the receiver has finished loading
    pending = []                                       # (time, fn): asynchronous effects
    errors = []

    # ---- sender side -------------------------------------------------------------------------
    class Group:
        def __init__(self, g):
            self.g, self.written, self.stg_full, self.stg_free = g, 0, -1, -1   # staging holds row stg_full

        def step(self):          # epilogue: write the next row into the staging buffer
            if self.written < n_rows[self.g] and self.stg_free >= self.written - 1 and self.stg_full < self.written:
                self.stg_full = self.written
                self.written += 1
                return True
            return False

    class Courier:
        def __init__(self, grp):
            self.grp, self.g = grp, grp.g
            self.r, self.published, self.landed, self.issued = 0, 0, 0, 0
            self.state = "wait_full"

        def announce(self, n):
            if n > self.published:
                if not announce_before_landing and self.landed < self.issued:
                    return False                      # wait_group 0: every store issued so far must have landed
                self.published = n
                pub[self.g] = n
            return True

        def step(self):
            g, r = self.g, self.r
            if self.state == "done":
                return False
            if r >= n_rows[g]:
                if self.announce(n_rows[g]):
                    self.state = "done"
                    return True
                return False
            if self.state == "wait_full":
                if self.grp.stg_full < r:
                    return False
                self.state = "cons"
                return True
            if self.state == "cons":
                if cons[g] < r + 1 - slots:
                    self.announce(r)                  # everything stored so far, before blocking on the consumer
                    return False
                self.state = "store"
                return True
            if self.state == "store":
                self.issued += 1
                row = r

                def land(row=row):
                    ring[g][row % slots] = (g, row)
                    self.landed += 1
                pending.append((now + rng.randint(1, 40), land))
                self.grp.stg_free = r                 # (read-out of the staging buffer is folded into this step)
                self.state = "publish"
                return True
            if self.state == "publish":
                next_waiting = self.grp.stg_full >= r + 1
                if r + 1 - self.published >= publish_every and not next_waiting:
                    if not self.announce(r + 1):
                        return False
                self.r += 1
                self.state = "wait_full"
                return True
            raise AssertionError(self.state)

    # ---- receiver side -----------------------------------------------------------------------
    order = interleave(*n_rows)
    got = []                                           # rows as the MMA issuer consumed them

    class Loader:
        def __init__(self):
            self.i, self.landed, self.consumed_steps = 0, [False] * len(order), 0

        def step(self):
            i = self.i
            if i >= len(order):
                return False
            if i >= RETIRE_LAG:
                o = i - RETIRE_LAG
                if not self.landed[o]:
                    return False
                s, k = order[o]
                cons[s] = max(cons[s], k)
            s, k = order[i]
            if pub[s] < k:
                return False
            if i - self.consumed_steps >= STAGES:      # A ring full
                return False

            def land(i=i, s=s, k=k):
                content = ring[s][(k - 1) % slots]
                if content != (s, k - 1):
                    errors.append((i, s, k, content))
                self.landed[i] = content
            pending.append((now + rng.randint(1, 40), land))
            self.i += 1
            return True

        def mma(self):
            j = self.consumed_steps
            if j < len(order) and self.landed[j]:
                got.append(self.landed[j])
                self.consumed_steps += 1
                return True
            return False

    groups = [Group(0), Group(1)]
    couriers = [Courier(groups[0]), Courier(groups[1])]
    loader = Loader()
    actors = [groups[0].step, groups[1].step, couriers[0].step, couriers[1].step, loader.step, loader.mma]
    idle = 0
    while loader.consumed_steps < len(order):
        now += 1
        if now > max_ticks:
            return {"deadlock": True, "errors": errors, "got": got}
        due = [p for p in pending if p[0] <= now]
        for p in due:
            pending.remove(p)
            p[1]()
        progressed = rng.choice(actors)()
        idle = 0 if (progressed or due) else idle + 1
        if idle > 20000 and not pending:               # nobody can move and nothing is in flight
            if not any(a() for a in actors):
                return {"deadlock": True, "errors": errors, "got": got}
            idle = 0
    # the loader still owes the last RETIRE_LAG `consumed` updates; the sender does not need them
    return {"deadlock": False, "errors": errors, "got": got}


@pytest.mark.parametrize("slots", [2, 3, 4, 8])
@pytest.mark.parametrize("publish_every", [1, 2, 3])
def test_handover_delivers_every_row_once_in_order(slots, publish_every):
    for seed, n_rows in enumerate([(37, 37), (40, 33), (1, 9), (0, 12), (25, 26)]):
        res = simulate(n_rows, slots, publish_every, seed)
        assert not res["deadlock"], (n_rows, slots, publish_every)
        assert res["errors"] == []
        assert res["got"] == [(s, k - 1) for (s, k) in interleave(*n_rows)]


def test_the_shipped_cadence_with_the_shipped_ring_depth():
    for seed in range(5):
        res = simulate((300, 297), slots=8, publish_every=2, seed=100 + seed)
        assert not res["deadlock"] and res["errors"] == [] and len(res["got"]) == 597


def test_a_flag_that_overtakes_the_data_is_caught():
    """What round 1's race was: the counter could become visible before the row it announces."""
    bad = 0
    for seed in range(10):
        res = simulate((60, 60), slots=8, publish_every=1, seed=seed, announce_before_landing=True)
        bad += bool(res["errors"])
    assert bad > 0
