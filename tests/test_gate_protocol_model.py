"""Exhaustive interleaving model of the device-side weight gate (bagua_b200/csrc/peer_kernels.cu: phase 3 of ``async_average_kernel``,
``gate_acquire_kernel``, ``gate_release_kernel``).

The kernels cannot run here (no GPU), but the protocol is five words and a handful of atomics, small enough to enumerate EVERY
interleaving of the trainer stream (acquire → touch the weights → release, several steps) with the CTAs of several consecutive
averaging kernels (arrive, the last one takes the gate or gives up after its bounded wait, siblings follow its decision, apply,
last one out re-arms the words).  Each memory operation of the CUDA code is one atomic step of the model.  Checked over all
reachable states:

* mutual exclusion: the trainer never touches the weights while a CTA applies the average;
* all CTAs of one kernel take the same decision (nobody applies half an average);
* no trap states: a state where everything has finished is reachable from every reachable state (no deadlock / livelock);
* once everything has finished the words are back to their armed values (state free, nobody waiting, counters zero).

The reference serialises the same two parties with a host mutex held across a stream synchronize
(/root/reference/bagua/torch_api/algorithms/async_model_average.py:120-147); the properties are the ones that mutex gives.
"""
from collections import deque

import pytest

STATE, WANT, DONE, ARRIVED, DECISION = range(5)


def _set(t, i, v):
    return t[:i] + (v,) + t[i + 1:]


def _trainer_steps(gate, tr):
    """tr = (pc, rounds_left, in_critical).  Mirrors gate_acquire_kernel / the step / gate_release_kernel; the timeout branch of the
    acquire (a wedged averaging kernel) is a fault path and not part of the model."""
    pc, left, crit = tr
    if pc == "done":
        return
    if pc == "want":          # while (gate[1] != 0 && gate[0] != 1) ...  -- first load
        yield gate, (("cas" if gate[WANT] == 0 else "state"), left, crit)
    elif pc == "state":       # second load of the loop condition
        yield gate, (("cas" if gate[STATE] == 1 else "want"), left, crit)
    elif pc == "cas":         # atomicCAS(&gate[0], 0, 1): done when the old value was 0 or 1
        if gate[STATE] == 0:
            yield _set(gate, STATE, 1), ("work", left, True)
        elif gate[STATE] == 1:
            yield gate, ("work", left, True)
        else:
            yield gate, ("cas", left, crit)
    elif pc == "work":        # forward / backward / optimizer step on the weights
        yield gate, ("release", left, True)
    elif pc == "release":     # atomicCAS(&gate[0], 1, 0)
        g = _set(gate, STATE, 0) if gate[STATE] == 1 else gate
        yield g, (("want" if left > 1 else "done"), left - 1, False)


def _cta_steps(gate, cta, ncta, may_time_out):
    """cta = (pc, got, applying)."""
    pc, got, applying = cta
    if pc == "done":
        return
    if pc == "arrive":        # atomicAdd(&gate[3], 1) + 1 == gridDim.x ?
        n = gate[ARRIVED]
        yield _set(gate, ARRIVED, n + 1), (("announce" if n + 1 == ncta else "follow"), got, applying)
    elif pc == "announce":    # atomicExch(&gate[1], 1)
        yield _set(gate, WANT, 1), ("take", got, applying)
    elif pc == "take":        # CAS free -> averaging, bounded
        if gate[STATE] == 0:
            yield _set(gate, STATE, 2), ("decide", 1, applying)
        else:
            yield gate, ("take", got, applying)
            if may_time_out:
                yield gate, ("decide", 0, applying)
    elif pc == "decide":      # atomicExch(&gate[4], got ? 1 : 2)
        yield _set(gate, DECISION, 1 if got else 2), ("apply?", got, applying)
    elif pc == "follow":      # spin on gate[4]
        d = gate[DECISION]
        yield (gate, ("follow", got, applying)) if d == 0 else (gate, ("apply?", 1 if d == 1 else 0, applying))
    elif pc == "apply?":
        yield gate, (("applying", got, True) if got else ("leave", got, False))
    elif pc == "applying":
        yield gate, ("leave", got, False)
    elif pc == "leave":       # atomicAdd(&gate[2], 1) == gridDim.x - 1 ?
        n = gate[DONE]
        yield _set(gate, DONE, n + 1), (("rearm2" if n == ncta - 1 else "done"), got, applying)
    elif pc == "rearm2":
        yield _set(gate, DONE, 0), ("rearm3", got, applying)
    elif pc == "rearm3":
        yield _set(gate, ARRIVED, 0), ("rearm4", got, applying)
    elif pc == "rearm4":
        yield _set(gate, DECISION, 0), ("handback", got, applying)
    elif pc == "handback":    # atomicCAS(&gate[0], 2, 0)
        yield (_set(gate, STATE, 0) if gate[STATE] == 2 else gate), ("clearwant", got, applying)
    elif pc == "clearwant":   # atomicExch(&gate[1], 0)
        yield _set(gate, WANT, 0), ("done", got, applying)


def _explore(ncta, trainer_rounds, kernel_rounds, may_time_out):
    fresh_ctas = tuple(("arrive", 0, False) for _ in range(ncta))
    start = ((0, 0, 0, 0, 0), ("want", trainer_rounds, False), fresh_ctas, kernel_rounds)
    seen, edges, queue = {start}, {}, deque([start])
    while queue:
        s = queue.popleft()
        gate, tr, ctas, kleft = s
        # ---- invariants of every reachable state
        assert not (tr[2] and any(c[2] for c in ctas)), f"trainer and averaging touch the weights together: {s}"
        decided = {c[1] for c in ctas if c[0] in ("applying", "leave", "done", "rearm2", "rearm3", "rearm4", "handback", "clearwant")}
        assert len(decided) <= 1, f"CTAs of one kernel disagree about applying: {s}"
        if any(c[2] for c in ctas):
            assert gate[STATE] == 2, f"average applied without holding the gate: {s}"
        if tr[2]:
            assert gate[STATE] == 1, f"trainer in its critical section without the gate: {s}"
        nxt = []
        for g2, tr2 in _trainer_steps(gate, tr):
            nxt.append((g2, tr2, ctas, kleft))
        if all(c[0] == "done" for c in ctas):
            if kleft > 1:     # the next averaging kernel on the same stream starts only after the previous one has retired
                nxt.append((gate, tr, fresh_ctas, kleft - 1))
        else:
            for i, c in enumerate(ctas):
                for g2, c2 in _cta_steps(gate, c, ncta, may_time_out):
                    nxt.append((g2, tr, ctas[:i] + (c2,) + ctas[i + 1:], kleft))
        edges[s] = nxt
        for n in nxt:
            if n not in seen:
                seen.add(n)
                queue.append(n)
    return seen, edges


def _finished(s):
    gate, tr, ctas, kleft = s
    return tr[0] == "done" and kleft == 1 and all(c[0] == "done" for c in ctas)


@pytest.mark.parametrize("ncta,trainer_rounds,kernel_rounds,may_time_out", [(1, 2, 2, False), (2, 2, 2, False), (3, 2, 1, False), (2, 2, 2, True), (3, 1, 2, True)])
def test_weight_gate_protocol_all_interleavings(ncta, trainer_rounds, kernel_rounds, may_time_out):
    seen, edges = _explore(ncta, trainer_rounds, kernel_rounds, may_time_out)
    finals = [s for s in seen if _finished(s)]
    assert finals, "no interleaving finishes"
    for gate, *_ in finals:
        assert gate == (0, 0, 0, 0, 0), f"gate words not re-armed at the end: {gate}"
    # no trap states: walk the graph backwards from the finished states
    back = {}
    for s, outs in edges.items():
        for n in outs:
            back.setdefault(n, []).append(s)
    alive, queue = set(finals), deque(finals)
    while queue:
        for p in back.get(queue.popleft(), ()):
            if p not in alive:
                alive.add(p)
                queue.append(p)
    stuck = seen - alive
    assert not stuck, f"{len(stuck)} reachable states can never finish, e.g. {next(iter(stuck))}"
    assert len(seen) > 50


def test_weight_gate_model_detects_a_broken_protocol():
    """The checker is not vacuous: without the fairness word's precondition on the trainer CAS — i.e. a trainer that writes
    state = 1 unconditionally — mutual exclusion is violated and the exploration says so."""
    orig = _trainer_steps

    def broken(gate, tr):
        pc, left, crit = tr
        if pc == "cas":
            yield _set(gate, STATE, 1), ("work", left, True)
        else:
            yield from orig(gate, tr)

    globals()["_trainer_steps"] = broken
    try:
        with pytest.raises(AssertionError):
            _explore(2, 2, 1, False)
    finally:
        globals()["_trainer_steps"] = orig
