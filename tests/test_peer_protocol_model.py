"""Model check of the two-flag protocol that keeps a multi-rank memory bank consistent without collectives
(csrc/peer.cuh, bank_tc.cu, peer.cu::bank_enqueue_peer_kernel).

Per step t (1-based) every rank r runs, in stream order,
    K3(t):   wait  enq_done[s] >= t-1  for every peer s;  READ the ring (every shard / its own copy);
             publish smooth_done[r] = t on every peer
    ENQ(t):  (side stream, joined before K3(t+1))  wait smooth_done[s] >= t for every peer s;
             WRITE its rows of step t into the owning shards / every copy;  publish enq_done[r] = t
The reference semantics (code/comatch.py:179-196) need, for every rank d that holds ring memory:
    (WAR) no rank writes step-t rows into d's memory before d's READ of step t has ended;
    (RAW) d's READ of step t+1 starts only after every rank's step-t rows are in d's memory.
The simulator interleaves the ranks' atomic actions at random (many seeds) and checks both, plus progress;
a mutated protocol (one wait removed) must be caught."""
import random

import pytest


def run(world, steps, seed, skip_wait=None):
    rng = random.Random(seed)
    enq_done = [[0] * world for _ in range(world)]       # enq_done[d][s]: flag in d's arena, written by s
    smooth_done = [[0] * world for _ in range(world)]
    reading = [0] * world                                # step whose READ is in progress on rank d (0 = none)
    read_ended = [0] * world                             # last step whose READ has ended on rank d
    written = [[0] * world for _ in range(world)]        # written[d][s]: last step whose rows from s are in d's memory
    violations = []

    def rank_program(r):
        for t in range(1, steps + 1):
            if skip_wait != "enq":
                yield ("wait", lambda t=t: all(enq_done[r][s] >= t - 1 for s in range(world) if s != r))
            # stream order on the own rank: ENQ(t-1) was joined before K3(t)
            yield ("do", lambda t=t: begin_read(r, t))
            yield ("do", lambda t=t: end_read(r, t))
            yield ("do", lambda t=t: publish(smooth_done, r, t))
            if skip_wait != "smooth":
                yield ("wait", lambda t=t: all(smooth_done[r][s] >= t for s in range(world) if s != r))
            for d in rng.sample(range(world), world):     # remote stores land in any order
                yield ("do", lambda t=t, d=d: write(r, d, t))
            yield ("do", lambda t=t: publish(enq_done, r, t))

    def begin_read(d, t):
        reading[d] = t
        for s in range(world):                            # RAW: all rows of step t-1 must be here
            if written[d][s] < t - 1:
                violations.append(("RAW", d, s, t))

    def end_read(d, t):
        reading[d], read_ended[d] = 0, t

    def write(s, d, t):
        if read_ended[d] < t:                             # WAR: d has not finished reading the pre-enqueue ring of step t
            violations.append(("WAR", s, d, t))
        written[d][s] = t

    def publish(flags, r, t):
        for d in range(world):
            flags[d][r] = t

    progs = [rank_program(r) for r in range(world)]
    pending = [next(p) for p in progs]
    done = [False] * world
    for _ in range(100000):
        runnable = [r for r in range(world) if not done[r] and (pending[r][0] == "do" or pending[r][1]())]
        if not runnable:
            break
        r = rng.choice(runnable)
        if pending[r][0] == "do":
            pending[r][1]()
        try:
            pending[r] = next(progs[r])
        except StopIteration:
            done[r] = True
    return all(done), violations


@pytest.mark.parametrize("world", [2, 3, 8])
def test_two_flag_protocol_has_no_hazard_and_no_deadlock(world):
    for seed in range(300 if world < 8 else 60):
        finished, violations = run(world, steps=4, seed=seed)
        assert finished, f"deadlock with seed {seed}"
        assert not violations, violations[:3]


@pytest.mark.parametrize("skip,kind", [("smooth", "WAR"), ("enq", "RAW")])
def test_each_wait_is_necessary(skip, kind):
    """Dropping the wait on "reads done" lets a fast rank overwrite rows a slow rank still needs (WAR); dropping the
    wait on "rows in" lets a rank smooth against a ring that misses a peer's rows (RAW)."""
    caught = set()
    for seed in range(200):
        _, violations = run(3, steps=4, seed=seed, skip_wait=skip)
        caught |= {v[0] for v in violations}
    assert kind in caught
