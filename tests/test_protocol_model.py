"""Model check of the hand-off protocol of the real-d DMMA kernel (wignerd/jl_b200/csrc/wd_k2_dmma.cuh,
k2_real_kernel): the two compute warps of an SM sub-partition pass ONE token back and forth (named barriers
bar_me / bar_other), every loop iteration consumes it once and emits it once, and unit indices come from a counter
shared by all eight warps of the CTA; the trig tables of item k+1 are made during item k (trig_ready / trig_free
mbarriers, 8 arrivals each).  The kernel grabs the index WHILE HOLDING the token; grabbing before the
wait (which would hide the atomic's latency) deadlocks on small items -- that variant hung the GPU once, and is
kept here as the negative control.  Pure Python, random interleavings; no GPU, no library."""
import random

import pytest


def simulate(units_per_item, grab_under_token, seed, npairs=4):
    """returns True if every warp terminates, False on deadlock"""
    rng = random.Random(seed)
    nwarps = 2 * npairs
    tokens = [0] * nwarps                    # pending arrivals on each warp's barrier (bar_me)
    counters = list(units_per_item)          # remaining units per item (the shared-memory counter, one per item)

    def partner(w):
        return w ^ npairs                    # (w, w + npairs) share a sub-partition

    nitems = len(counters)
    trig_ready = [0] * (nitems + 1)          # arrivals: the tables of item k are complete at 2 * npairs
    trig_free = [0] * (nitems + 1)           # arrivals: every warp has finished item k

    def prefetch(kn):                        # this warp's share of the trig tables of item kn (other buffer)
        while kn >= 2 and trig_free[kn - 2] < nwarps:
            yield "blocked"
        trig_ready[kn] += 1

    def warp(w):
        if w >= npairs:                      # second-turn warps hand the first turn to their partners
            tokens[partner(w)] += 1
        yield from prefetch(0)
        for k in range(nitems):
            while trig_ready[k] < nwarps:    # mbar_wait(trig_ready)
                yield "blocked"
            pre = k + 1 >= nitems
            while True:
                have = None
                if not grab_under_token:     # negative control: index grabbed before the token wait
                    have = counters[k] > 0
                    counters[k] -= 1
                    yield "ready"
                while tokens[w] == 0:        # named_bar_sync(bar_me)
                    yield "blocked"
                tokens[w] -= 1
                if grab_under_token:
                    have = counters[k] > 0
                    counters[k] -= 1
                if have:
                    yield "ready"            # main loop, first part
                    tokens[partner(w)] += 1  # release at ~2/3 of the main loop
                    yield "ready"            # rest of the main loop + epilogue
                else:
                    tokens[partner(w)] += 1  # leaving the item: pass the token on
                if not pre:                  # after the first unit, or when there is none
                    yield from prefetch(k + 1)
                    pre = True
                if not have:
                    break
            trig_free[k] += 1
        return

    gens = {w: warp(w) for w in range(nwarps)}
    state = {w: "ready" for w in gens}
    stuck = 0
    while gens:
        w = rng.choice(list(gens))
        try:
            state[w] = next(gens[w])
            stuck = stuck + 1 if state[w] == "blocked" else 0
        except StopIteration:
            del gens[w]
            del state[w]
            stuck = 0
        if stuck > 5000:                     # every remaining warp has been polling without progress
            return False
    return True


@pytest.mark.parametrize("units", [[0], [1], [2], [5], [8], [11], [1, 1, 1], [3, 0, 7, 1], [26, 26, 26], [9, 202, 1]])
def test_grab_under_token_always_terminates(units):
    for seed in range(300):
        assert simulate(units, True, seed), (units, seed)


def test_grab_before_token_can_deadlock():
    # the negative control: with one unit in the item a second-turn warp can win it after its partner has left
    assert any(not simulate([1], False, seed) for seed in range(300))


def simulate_stream(tiles, stages, seed, nwarps=8):
    """Stage pipeline of the streaming kernel (wd_k2_stream.cuh): chunk g of the CTA's running chunk sequence is
    requested by warp g % 8 after its work on chunk g - (stages-1) (or in the tile's prologue), consumers wait on
    full[g] and release with one arrival each on empty[stage][use].  True if every warp terminates."""
    rng = random.Random(seed)
    full = set()
    empty = {}                                # (stage, use) -> arrivals

    def produce(g):
        si, use = g % stages, g // stages
        while use > 0 and empty.get((si, use - 1), 0) < nwarps:     # mbar_wait(sempty): previous chunk of the stage
            yield "blocked"
        full.add(g)

    def warp(w):
        gch = 0
        for nck in tiles:
            for i in range(min(stages - 1, nck)):
                if w == (gch + i) % nwarps:
                    yield from produce(gch + i)
            for kc in range(nck):
                g = gch + kc
                while g not in full:          # mbar_wait(sfull)
                    yield "blocked"
                yield "ready"                 # the DMMAs of the chunk
                empty[(g % stages, g // stages)] = empty.get((g % stages, g // stages), 0) + 1
                nx = kc + stages - 1
                if nx < nck and w == (gch + nx) % nwarps:
                    yield from produce(gch + nx)
            gch += nck
            yield "ready"                     # epilogue

    gens = {w: warp(w) for w in range(nwarps)}
    stuck = 0
    while gens:
        w = rng.choice(list(gens))
        try:
            st = next(gens[w])
            stuck = stuck + 1 if st == "blocked" else 0
        except StopIteration:
            del gens[w]
            stuck = 0
        if stuck > 20000:                     # every remaining warp has been polling without progress
            return False
    return True


@pytest.mark.parametrize("tiles", [[1], [2], [3, 1, 2], [4, 4, 4], [17] * 3, [8, 1, 9, 2, 18], [5] * 7])
@pytest.mark.parametrize("stages", [2, 3, 4, 5])
def test_stream_pipeline_terminates(tiles, stages):
    for seed in range(40):
        assert simulate_stream(tiles, stages, seed), (tiles, stages, seed)
