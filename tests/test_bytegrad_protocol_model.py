"""Schedule-exploring model of the synchronisation skeleton of ``bytegrad_kernel`` (bagua_b200/csrc/bytegrad_kernels.cu): one launch
does  A min/max → grid barrier → B quantise + store into the OWNER's inbox → peer barrier (per CTA row) + grid barrier → C reduce the
P received copies → grid barrier → D re-quantise + store into EVERY peer's outbox → peer barrier + grid barrier → E decode.

Two things make it subtle.  (1) The peer barrier is per CTA row, but the data a row reads was written by a DIFFERENT row of the
peer: in B the CTAs ``cj·bpc … cj·bpc+bpc−1`` of rank X write rank ``cj``'s inbox, in C all rows of the owner read it striped — so
the per-row barrier alone proves nothing and the grid-wide rendezvous after it is load-bearing.  (2) inbox and outbox are single
buffers reused by consecutive launches; a fast rank may be a whole launch ahead of a slow one.

The model runs random interleavings of all CTAs of 2-3 ranks over several launches (stream order per rank, grid barriers local,
peer barriers per row) and checks every read against the launch it belongs to; the kernel without its grid rendezvous after the
peer barriers must be caught."""
import random

import pytest


def run(nranks, bpc, launches, rng, grid_after_peer=True):
    nb = nranks * bpc                                    # the host guarantees nb % P == 0
    inbox = [[[None] * nb for _ in range(nranks)] for _ in range(nranks)]     # [owner][source slot][stripe cell] = launch id
    outbox = [[[None] * nb for _ in range(nranks)] for _ in range(nranks)]
    flags = [[[0] * nranks for _ in range(nb)] for _ in range(nranks)]        # [owner][row][source] = epoch
    grid_count = [0] * nranks                                                 # arrivals at the rank-local grid barrier (monotonic)
    launch = [0] * nranks
    program = ["A", "grid", "B", "peer_arrive", "peer_wait"] + (["grid"] if grid_after_peer else []) + \
              ["C", "grid", "D", "peer_arrive", "peer_wait"] + (["grid"] if grid_after_peer else []) + ["E"]

    def fresh():
        return [{"ip": 0, "grids": 0, "peers": 0, "waiting": None} for _ in range(nb)]

    ctas = [fresh() for _ in range(nranks)]
    grids_per_launch = program.count("grid")
    steps = 0
    while any(x < launches for x in launch):
        steps += 1
        assert steps < 400_000, "model made no progress"
        runnable = []
        for r in range(nranks):
            if launch[r] >= launches:
                continue
            for b, cta in enumerate(ctas[r]):
                if cta["ip"] >= len(program):
                    continue
                op = program[cta["ip"]]
                if op == "grid" and cta["waiting"] is not None and grid_count[r] < cta["waiting"]:
                    continue
                if op == "peer_wait" and not all(flags[r][b][p] >= cta["epoch"] for p in range(nranks)):
                    continue
                runnable.append((r, b))
        r, b = rng.choice(runnable)
        cta, c = ctas[r][b], launch[r]
        op = program[cta["ip"]]
        cj, sb = b // bpc, b % bpc
        if op == "grid":
            if cta["waiting"] is None:          # arrive: the target is "everybody of this rank has arrived at this barrier instance"
                grid_count[r] += 1
                cta["waiting"] = (c * grids_per_launch + cta["grids"] + 1) * nb
                continue
            cta["waiting"] = None
            cta["grids"] += 1
        elif op == "B":                          # my share (tiles sb, sb+bpc, …) of chunk cj → rank cj's inbox, slot = me
            for cell in range(nb):
                if cell % bpc == sb:
                    inbox[cj][r][cell] = c
        elif op == "peer_arrive":
            cta["peers"] += 1
            cta["epoch"] = 2 * c + cta["peers"]
            for p in range(nranks):
                flags[p][b][r] = cta["epoch"]
        elif op == "C":                          # striped over ALL rows: row b reads cell b of every slot
            for s in range(nranks):
                if inbox[r][s][b] != c:
                    return f"C: rank {r} launch {c} row {b} read inbox slot {s} holding launch {inbox[r][s][b]}"
        elif op == "D":                          # my stripe of the reduced chunk → every peer's outbox, slot = me
            for p in range(nranks):
                outbox[p][r][b] = c
        elif op == "E":
            for s in range(nranks):
                if outbox[r][s][b] != c:
                    return f"E: rank {r} launch {c} row {b} read outbox slot {s} holding launch {outbox[r][s][b]}"
        cta["ip"] += 1
        if all(x["ip"] >= len(program) for x in ctas[r]):
            launch[r] += 1
            ctas[r] = fresh()
    return None


@pytest.mark.parametrize("nranks,bpc", [(2, 1), (2, 2), (3, 1), (3, 2)])
def test_bytegrad_buffers_are_race_free_across_launches(nranks, bpc):
    for seed in range(120):
        assert run(nranks, bpc, launches=3, rng=random.Random(seed)) is None


def test_the_grid_rendezvous_after_the_peer_barrier_is_load_bearing():
    """Per-row peer barriers alone do not order a row's reads after the OTHER rows' remote writes."""
    found = None
    for seed in range(300):
        found = found or run(2, 2, launches=2, rng=random.Random(seed), grid_after_peer=False)
    assert found is not None


# ---------------------------------------------------------------------------------------------------------------------
# lpdec_ring_kernel: ONE peer barrier per launch, the three-slot box double-buffered by launch parity
# ---------------------------------------------------------------------------------------------------------------------
def run_ring(nranks, nb, launches, rng, double_buffered=True):
    """Launch c: every CTA quantises its stripe of the difference and stores it into the LEFT neighbour's "from-right" slot, the RIGHT
    neighbour's "from-left" slot and the own slot of box[parity(c)]; per-row peer barrier + grid rendezvous; every CTA reads its stripe
    of the three slots of its own box.  There is no closing barrier: a fast rank enters launch c+1 while a neighbour still reads
    launch c's slots — which is why the box has two halves."""
    halves = 2 if double_buffered else 1
    box = [[[[None] * nb for _ in range(3)] for _ in range(halves)] for _ in range(nranks)]     # [owner][half][slot][stripe]
    flags = [[[0] * nranks for _ in range(nb)] for _ in range(nranks)]
    grid_count, launch = [0] * nranks, [0] * nranks
    program = ["send", "peer_arrive", "peer_wait", "grid", "receive"]

    def fresh():
        return [{"ip": 0, "waiting": None} for _ in range(nb)]

    ctas = [fresh() for _ in range(nranks)]
    steps = 0
    while any(x < launches for x in launch):
        steps += 1
        assert steps < 300_000
        runnable = []
        for r in range(nranks):
            if launch[r] >= launches:
                continue
            for b, cta in enumerate(ctas[r]):
                if cta["ip"] >= len(program):
                    continue
                op = program[cta["ip"]]
                if op == "grid" and cta["waiting"] is not None and grid_count[r] < cta["waiting"]:
                    continue
                if op == "peer_wait" and not all(flags[r][b][p] >= launch[r] + 1 for p in range(nranks)):
                    continue
                runnable.append((r, b))
        r, b = rng.choice(runnable)
        cta, c = ctas[r][b], launch[r]
        op = program[cta["ip"]]
        half = (c & 1) if double_buffered else 0
        left, right = (r - 1) % nranks, (r + 1) % nranks
        if op == "send":
            box[left][half][1][b] = (r, c)       # the left neighbour's "from-right" slot
            box[right][half][0][b] = (r, c)      # the right neighbour's "from-left" slot
            box[r][half][2][b] = (r, c)
        elif op == "peer_arrive":
            for p in range(nranks):
                flags[p][b][r] = c + 1
        elif op == "grid":
            if cta["waiting"] is None:
                grid_count[r] += 1
                cta["waiting"] = (c + 1) * nb
                continue
            cta["waiting"] = None
        elif op == "receive":
            want = ((left, c), (right, c), (r, c))
            got = tuple(box[r][half][s][b] for s in range(3))
            if got != want:
                return f"rank {r} launch {c} row {b} received {got}, expected {want}"
        cta["ip"] += 1
        if all(x["ip"] >= len(program) for x in ctas[r]):
            launch[r] += 1
            ctas[r] = fresh()
    return None


@pytest.mark.parametrize("nranks,nb", [(2, 1), (2, 3), (3, 2), (4, 2)])
def test_ring_box_halves_make_the_single_barrier_sufficient(nranks, nb):
    for seed in range(120):
        assert run_ring(nranks, nb, launches=4, rng=random.Random(seed)) is None


def test_a_single_buffered_ring_box_is_caught():
    found = None
    for seed in range(300):
        found = found or run_ring(3, 2, launches=3, rng=random.Random(seed), double_buffered=False)
    assert found is not None
