"""Executable model of the step schedule of the row-streaming first conv (reve_b200/csrc/conv0_rows.cu): which MMAs a
step issues, into which TMEM slot, with which accumulate flag, and when a pair of output rows is handed to the epilogue.
The kernel's loop is restated here line by line (same conditions, same rotation of the slot registers); the checks are
the properties the kernel relies on:
  * every output row of a segment receives exactly the three vertical taps, from the right input rows, the first of
    them overwriting (no TMEM zero-fill exists), and nothing else;
  * no MMA targets a row outside the segment (rows of other CTAs, rows beyond the canvas);
  * a pair is committed to the epilogue after its last contribution and before its slot is opened again, and at most
    kSlots pairs are between 'opened' and 'drained' when the issuer honours the acc_empty barrier;
  * all roles walk the same sequence of segments, steps and events for any split of the canvas over CTAs.
CPU only; the GPU tests check the numbers, this checks the bookkeeping for every (ya, n) a CTA range can produce."""
import random

K_SLOTS = 4
K_STAGES = 6
K_GROUPS = 3


def segments(total_rows, n_strips, grid):
    """Walk::next for every CTA: (cta, strip, ya, n)."""
    total = n_strips * total_rows
    out = []
    for b in range(grid):
        p, hi = total * b // grid, total * (b + 1) // grid
        while p < hi:
            strip, ya = divmod(p, total_rows)
            n = min(total_rows - ya, hi - p)
            out.append((b, strip, ya, n))
            p += n
    return out


def issue_segment(ya, n, state, log):
    """The MMA thread's loop over one segment.  state: slot / d_open / d_done / stage as the kernel carries them."""
    pairs = (n + 1) // 2
    for s in range(pairs + 1):
        open0, open1 = 2 * s < n, 2 * s + 1 < n
        done1 = s >= 1 and 2 * s - 1 < n
        if open0:
            log.append(("wait_empty", state["slot"]))
        log.append(("wait_full", state["stage"]))
        ra, rb = ya + 2 * s - 1, ya + 2 * s            # the two input rows of the step
        d_open, d_done = state["d_open"], state["d_done"]
        if open0:
            log.append(("mma", d_open, 0, ra, 0, False))     # (slot, row in pair, input row, tap, accumulate)
        if done1:
            log.append(("mma", d_done, 1, ra, 1, True))
        if s >= 1:
            log.append(("mma", d_done, 0, ra, 2, True))
        if open1:
            log.append(("mma", d_open, 1, rb, 0, False))
        if open0:
            log.append(("mma", d_open, 0, rb, 1, True))
        if done1:
            log.append(("mma", d_done, 1, rb, 2, True))
        log.append(("commit_empty", state["stage"]))
        if s >= 1:
            log.append(("commit_full", state["f_done"]))
        state["d_done"], state["f_done"] = d_open, state["slot"]
        if open0:
            state["slot"] = (state["slot"] + 1) % K_SLOTS
            state["d_open"] = state["slot"]
        state["stage"] = (state["stage"] + 1) % K_STAGES


def check_cta(segs):
    """segs: the (ya, n) segments of one CTA, in order."""
    state = {"slot": 0, "d_open": 0, "d_done": None, "f_done": None, "stage": 0}
    log = []
    bounds = []
    for ya, n in segs:
        start = len(log)
        issue_segment(ya, n, state, log)
        bounds.append((start, len(log), ya, n))
    # replay: slots hold (segment index, pair index); rows collect their contributions
    events = 0
    for si, (a, b, ya, n) in enumerate(bounds):
        pairs = (n + 1) // 2
        slot_of = {}                      # slot -> pair index currently accumulating there
        contrib = {y: [] for y in range(ya, ya + n)}
        opened, completed = [], []
        for rec in log[a:b]:
            if rec[0] == "mma":
                _, slot, h, r, tap, acc = rec
                if not acc and h == 0:    # the first row of a pair opens the slot
                    slot_of[slot] = len(opened)
                    opened.append(slot)
                pair = slot_of[slot]
                y = ya + 2 * pair + h
                assert ya <= y < ya + n, ("MMA outside the segment", ya, n, rec)
                assert r == y + tap - 1, ("wrong input row for the tap", y, rec)
                assert acc == (len(contrib[y]) > 0), ("first contribution must overwrite, later ones accumulate", y, rec)
                contrib[y].append(tap)
            elif rec[0] == "commit_full":
                pair = slot_of[rec[1]]
                for h in (0, 1):
                    y = ya + 2 * pair + h
                    if y < ya + n:
                        assert contrib[y] == [0, 1, 2], ("pair handed over before it is complete", y, contrib[y])
                completed.append(pair)
        assert completed == list(range(pairs)), (completed, pairs)
        assert all(v == [0, 1, 2] for v in contrib.values())
        assert opened == [(events + i) % K_SLOTS for i in range(pairs)]      # the epilogue's e % kSlots sees the same slots
        events += pairs
    # steps: one wait_full / commit_empty per step, stages in ring order
    stages = [r[1] for r in log if r[0] == "wait_full"]
    assert stages == [i % K_STAGES for i in range(len(stages))]
    assert len(stages) == sum((n + 1) // 2 + 1 for _, n in segs)
    # a slot is re-opened only after the pair that used it before was committed: with the acc_empty wait in front
    # of every opening, pairs between "opened" and "drained" never exceed kSlots; in issue order that is
    # opened - committed <= 2 (the pair being opened and the one being completed)
    live = 0
    for rec in log:
        if rec[0] == "wait_empty":
            live += 1
            assert live <= 2
        elif rec[0] == "commit_full":
            live -= 1
    return len(stages), events


def test_every_segment_shape_gets_exactly_its_taps():
    for ya in (0, 1, 2, 7):
        for n in range(1, 40):
            check_cta([(ya, n)])


def test_consecutive_segments_share_the_rings():
    rng = random.Random(5)
    for _ in range(200):
        segs = [(rng.randrange(0, 50), rng.randrange(1, 30)) for _ in range(rng.randrange(1, 5))]
        check_cta(segs)


def test_all_roles_count_the_same_steps_and_events_for_any_grid():
    """Producers count steps (group = step % 3), the epilogue counts events (group = event % 3): both derive their
    counters from the segment list alone, so the lists must cover every (strip, row) exactly once."""
    for rows, strips, grid in ((31, 1, 31), (1205, 17, 148), (4823, 17, 148), (542, 6, 148), (3, 2, 5), (100, 3, 7)):
        seen = set()
        per_cta = {}
        for b, strip, ya, n in segments(rows, strips, grid):
            assert n >= 1 and ya + n <= rows
            for y in range(ya, ya + n):
                assert (strip, y) not in seen
                seen.add((strip, y))
            per_cta.setdefault(b, []).append((ya, n))
        assert len(seen) == rows * strips
        for segs in per_cta.values():
            steps, events = check_cta(segs)
            # every producer group and every epilogue group gets work whenever there are at least three steps / events
            if steps >= K_GROUPS:
                assert {j % K_GROUPS for j in range(steps)} == set(range(K_GROUPS))
            assert events == sum((n + 1) // 2 for _, n in segs)
