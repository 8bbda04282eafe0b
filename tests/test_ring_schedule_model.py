"""Host-side model of the schedule of the streaming interaction backward (csrc/interact_warp.cu:
bwd_ring_pass / bwd_ring2_pass): a ring of RS shared-memory slots per warp filled by cp.async groups, two passes
over the F rows of T, the second pass's first rows requested at the end of the first.  The kernel relies on three
facts that are index arithmetic, not CUDA: (1) when iteration j waits for "all but the RS - 1 youngest groups", the
group that carries row j of the current pass has completed; (2) the slot it reads then holds exactly that row;
(3) a slot is only refilled after the iteration that consumed it.  This model replays the schedule and asserts
them, for the shipped geometry (F = 27, RS = 12) and every other (F, RS) the launcher's guard admits."""
import pytest


def replay(F: int, RS: int):
    slots = [None] * RS            # (pass, row) a landed copy has written, or None
    inflight = []                  # committed groups, oldest first: each a list of (slot, pass, row) copies (maybe empty)
    landed_groups = 0
    consumed = []

    def commit(copies):
        inflight.append(list(copies))

    def wait_all_but(n):
        nonlocal landed_groups
        while len(inflight) > n:   # cp.async groups complete in commit order
            for slot, p, r in inflight.pop(0):
                slots[slot] = (p, r)
            landed_groups += 1

    for r in range(RS):            # prime: one group per row
        commit([(r, 0, r)])
    for p, slot0, refill in ((0, 0, True), (1, F % RS, False)):
        for j in range(F):
            slot = (slot0 + j) % RS
            wait_all_but(RS - 1)
            assert slots[slot] == (p, j), (F, RS, p, j, slot, slots[slot])     # facts (1) and (2)
            consumed.append((p, j))
            slots[slot] = None                                                  # fact (3): nothing may land here before now
            if j + RS < F:
                commit([(slot, p, j + RS)])
            elif refill:
                commit([(slot, p + 1, j + RS - F)])
            else:
                commit([])                                                      # empty group: the count stays uniform
    wait_all_but(0)
    assert all(s is None for s in slots), "a copy landed that nobody consumed"
    return consumed


@pytest.mark.parametrize("F,RS", [(27, 12), (27, 9), (27, 27), (12, 12), (13, 12), (32, 12), (27, 8), (27, 1)])
def test_ring_schedule_delivers_every_row_in_order(F, RS):
    consumed = replay(F, RS)
    assert consumed == [(0, j) for j in range(F)] + [(1, j) for j in range(F)]


def test_row_paired_pass_split_covers_every_output_row_once():
    """bwd_ring2: output rows are split into a first pass of H1 rows and a second of H2 (padding rows included), both
    multiples of 4 (one LDS.128 of S = four output rows), inside the padded row pitch FP4."""
    for F in range(12, 33):
        FP4 = (F + 3) & ~3
        H1 = ((FP4 // 2) + 3) & ~3
        H2 = FP4 - H1
        assert H1 % 4 == 0 and H2 % 4 == 0 and H2 >= 0 and H1 + H2 == FP4 >= F
        rows = [f for f in range(0, H1) if f < F] + [H1 + f for f in range(0, H2) if H1 + f < F]
        assert rows == list(range(F))
        assert (F * FP4 + 4) * 4 % 16 == 0      # the ring behind S stays 16-byte aligned (cp.async 16)
