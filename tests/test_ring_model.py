"""Host-side model of the weight-ring protocol of csrc/tc_conv_patch.cu (no GPU): one producer filling the ring slots in order,
two MMA issuers on alternate accumulation chains, mbarrier PARITY waits.  A parity wait only tells a barrier's current phase from
the previous one, so an issuer that waits for fill n + 1 of a slot without having seen fill n can be satisfied by fill n - 1 when
fill n straggles ("a weight-ring hazard").  The model reproduces that failure for the configurations that
showed it or could show it, and checks that the kernel's guard -- restated here with the kernel's own bookkeeping (gtap, own_hist,
tap_seen) -- removes it."""
import random

import pytest


def simulate(ring, chains, guard, seed, straggle=0.05, steps=4000):
    """chains: tap counts of the accumulation chains in ring order, repeated; chain c belongs to issuer c % 2.
    guard: False, True (checked before every tap, FUSE8) or "chain" (checked once before a chain's first tap, the 8-slot rings).
    Returns (taps issued, taps issued on a slot that did not hold their tile)."""
    rng = random.Random(seed)
    done_fills = [0] * ring                 # completed fills per slot = completed barrier phases
    released = [0] * ring                   # taps of the slot whose MMAs are complete
    in_flight = {}                          # slot -> completion time of the fill in flight
    next_fill = 0
    tap_seen = [-1, -1]
    # per issuer: position in the chain sequence and the kernel's bookkeeping
    st = [dict(chain=i, k=0, gtap=0, hist=0xffffffff, busy_until=0, stall_until=0) for i in range(2)]
    first_tap = []                          # first tap index of chain c (grown on demand)

    def chain_len(c):
        return chains[c % len(chains)]

    def first_of(c):
        while len(first_tap) <= c:
            first_tap.append((first_tap[-1] + chain_len(len(first_tap) - 1)) if first_tap else 0)
        return first_tap[c]

    issued = bad = 0
    pending_release = []                    # (time, slot)
    for t in range(steps):
        # fills complete, MMAs complete
        for s in list(in_flight):
            if in_flight[s] <= t:
                done_fills[s] += 1
                del in_flight[s]
        for item in [p for p in pending_release if p[0] <= t]:
            released[item[1]] += 1
            pending_release.remove(item)
        # producer: fill tap next_fill when its slot's previous tap is released (one fill per slot at a time)
        s = next_fill % ring
        if released[s] >= next_fill // ring and s not in in_flight and done_fills[s] == next_fill // ring:
            delay = rng.randint(8, 14) + (rng.randint(60, 160) if rng.random() < straggle else 0)
            in_flight[s] = t + delay
            next_fill += 1
        # issuers
        for me in (0, 1):
            x = st[me]
            if t < x["busy_until"] or t < x["stall_until"]:
                continue
            c = x["chain"]
            # bookkeeping for the other issuer's chains skipped since my last chain (as the kernel does on the skip path)
            g = first_of(c) + x["k"]
            if x["k"] == 0 and x["gtap"] != first_of(c):
                skipped = first_of(c) - x["gtap"]
                x["hist"] = (x["hist"] << skipped) & 0xffffffff
                x["gtap"] = first_of(c)
            slot, want = g % ring, (g // ring) & 1
            if guard == "chain":
                if x["k"] == 0:
                    need = -1
                    for k in range(chain_len(c)):
                        if g + k >= ring and not (x["hist"] >> (ring - 1 - k)) & 1:
                            need = g + k - ring
                    if tap_seen[me ^ 1] < need:
                        continue                                    # spin before the chain starts
            elif guard and g >= ring and not (x["hist"] >> (ring - 1 - x["k"])) & 1:
                if tap_seen[me ^ 1] < g - ring:
                    continue                                        # spin
            if (done_fills[slot] & 1) == want:
                continue                                            # parity wait not satisfied
            # the wait passed: which tile does the slot hold?
            issued += 1
            if done_fills[slot] - 1 != g // ring or slot in in_flight:
                bad += 1
            tap_seen[me] = g
            mma = 6
            x["busy_until"] = t + 1
            pending_release.append((t + mma, slot))
            x["k"] += 1
            if x["k"] == chain_len(c):                              # chain done: own taps enter the history, next own chain
                n = chain_len(c)
                x["hist"] = ((x["hist"] << n) | ((1 << n) - 1)) & 0xffffffff
                x["gtap"] = first_of(c) + n
                x["chain"], x["k"] = c + 2, 0
                x["stall_until"] = t + rng.randint(0, 40)          # TMEM slot / epilogue back-pressure
    return issued, bad


DCONV7 = [3, 1, 3, 3, 3, 3, 3, 3, 3]        # chains of the four phases of dconv7 (4, 6, 6, 9 taps)


def test_five_slot_ring_needs_the_guard():
    """FUSE8: 5 slots, two three-tap chains in flight.  Without the guard a straggling fill lets a wait pass on a stale tile."""
    bad_without = sum(simulate(5, DCONV7, False, s)[1] for s in range(20))
    assert bad_without > 0
    for s in range(20):
        issued, bad = simulate(5, DCONV7, True, s)
        assert issued > 300 and bad == 0, s


def test_six_tap_chains_on_the_eight_slot_ring_are_unsafe():
    """The dconv1 variant that was withdrawn: two six-tap chains span 12 taps of an 8-slot ring."""
    assert sum(simulate(8, [6, 3, 3, 3, 6, 3, 3, 3], False, s, straggle=0.1)[1] for s in range(20)) > 0
    for s in range(10):
        assert simulate(8, [6, 3, 3, 3, 6, 3, 3, 3], True, s, straggle=0.1)[1] == 0


@pytest.mark.parametrize("chains", [[3, 3, 3], [3, 3, 3, 3, 3], DCONV7])
@pytest.mark.parametrize("guard", [True, "chain"])
def test_guard_never_deadlocks_and_never_reads_a_stale_tile(chains, guard):
    for ring in (5, 8):
        for s in range(10):
            issued, bad = simulate(ring, chains, guard, 100 + s, straggle=0.2)
            assert issued > 200 and bad == 0, (ring, s)


def test_three_tap_chains_on_the_eight_slot_ring_without_a_guard():
    """The round-1 protocol: safe only as long as no fill straggles past several later ones.  The model (exaggerated stragglers)
    finds the stale wait; never observed on hardware."""
    assert sum(simulate(8, [3, 3, 3], False, s)[1] for s in range(20)) > 0
    assert sum(simulate(8, [3, 3, 3], "chain", s)[1] for s in range(20)) == 0
