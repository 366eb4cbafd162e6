"""Executable model of the stage-ring protocol of the experimental update kernel (harmony_b200/csrc/
update_kernel3.cuh): an in-order producer refills D ring slots guarded by full/empty mbarriers (waited on by
phase parity), consumer warps draw the stages of a block step from a ticket counter and meet at a barrier at the
end of the step.  Random schedules check the two properties the kernel relies on:
  * no deadlock and every stage consumed exactly once;
  * a warp never reads a slot that still holds an older stage — which requires that at most D warps draw tickets
    (a parity wait cannot tell the first refill of a slot from the second).  The model also shows that the rule
    is needed: without it the aliasing is found within a few schedules."""
import random

import pytest


def _try_wait(phase, parity):
    # mbarrier.try_wait.parity: true iff the phase with this parity is the immediately preceding (completed) one
    return (phase & 1) != parity


def simulate(D, NW, steps_stages, cap_by_depth, seed):
    rnd = random.Random(seed)
    full_ph, empty_ph, ring = [0] * D, [0] * D, [None] * D
    total = sum(steps_stages)
    produced = 0
    takes = [(w < D) if cap_by_depth else True for w in range(NW)]
    cons = [{"step": 0, "state": "draw", "g": None} for _ in range(NW)]
    tick = [0] * len(steps_stages)
    gbase = [0]
    for n in steps_stages:
        gbase.append(gbase[-1] + n)
    arrived = [0] * len(steps_stages)
    consumed, idle = [], 0
    while not (all(c["step"] >= len(steps_stages) for c in cons) and produced >= total):
        progressed = False
        actor = rnd.randrange(NW + 1)
        if actor == NW:  # producer: stage g -> slot g % D after the slot's previous use was released
            if produced < total:
                slot, use = produced % D, produced // D
                if use == 0 or _try_wait(empty_ph[slot], (use - 1) & 1):
                    ring[slot] = produced
                    full_ph[slot] += 1
                    produced += 1
                    progressed = True
        else:
            c = cons[actor]
            s = c["step"]
            if s < len(steps_stages):
                if c["state"] == "draw":
                    i = 10 ** 9
                    if takes[actor]:
                        i = tick[s]
                        tick[s] += 1
                    if i >= steps_stages[s]:
                        c["state"] = "barrier"
                        arrived[s] += 1
                    else:
                        c["g"] = gbase[s] + i
                        c["state"] = "wait"
                    progressed = True
                elif c["state"] == "wait":
                    slot, use = c["g"] % D, c["g"] // D
                    if _try_wait(full_ph[slot], use & 1):
                        if ring[slot] != c["g"]:
                            return "alias"
                        consumed.append(c["g"])
                        c["state"] = "release"
                        progressed = True
                elif c["state"] == "release":
                    empty_ph[c["g"] % D] += 1
                    c["state"] = "draw"
                    progressed = True
                elif arrived[s] == NW:  # end-of-step barrier of the group
                    c["step"] += 1
                    c["state"] = "draw"
                    progressed = True
        idle = 0 if progressed else idle + 1
        if idle > 20000:
            return "deadlock"
    return "ok" if sorted(consumed) == list(range(total)) else "lost"


@pytest.mark.parametrize("D", [2, 3, 5, 8])
@pytest.mark.parametrize("NW", [2, 6, 8])
def test_ring_protocol_is_safe_with_depth_capped_consumers(D, NW):
    for seed in range(12):
        rnd = random.Random(1000 * D + 10 * NW + seed)
        steps = [rnd.randrange(0, 14) for _ in range(5)]
        assert simulate(D, NW, steps, True, seed) == "ok"


def test_ring_protocol_needs_the_depth_cap():
    outcomes = {simulate(2, 8, [12, 9, 13], False, seed) for seed in range(20)}
    assert "alias" in outcomes
