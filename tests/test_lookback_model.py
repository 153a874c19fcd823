"""Model check of the compact-descriptor look-back protocol (megahit_b200/csrc/mhb_sort3.cuh, bit 15 of the variant
field): a randomised interleaving simulation, no GPU.  Every tile is a coroutine that (1) publishes its digit count
early, (2) walks back exactly as the CUDA code does - 16-byte group snapshots that may be stale, re-poll on
unpublished words, stop at a status-3 word and fetch that tile's inclusive prefix, which may become visible AFTER the
flag - and (3) publishes its own inclusive prefix and flag.  A random scheduler interleaves the steps of a bounded
number of in-flight tiles (persistent CTAs taking tickets in order).  Every tile's prefix must equal the exact sum."""
import random

import pytest


def run(seed, T, maxc=50):
    rnd = random.Random(seed)
    counts = [rnd.randint(0, maxc) for _ in range(T)]
    EP = 5
    part = [0] * (4 * ((T + 3) // 4))      # status<<30 | ep<<22 | count
    incl = [None] * T                       # inclusive prefix or None
    result = [None] * T
    started = [False] * T

    def word(status, c):
        return (status << 30) | (EP << 22) | c

    def tile(t):
        # early publish
        if t == 0:
            incl[0] = counts[0]
            yield
            part[0] = word(3, counts[0])
            result[0] = 0
            return
        part[t] = word(1, counts[t])
        yield
        # prefetch group of t-1 (snapshot)
        tcur = t - 1
        g = tcur >> 2
        cur = part[4 * g:4 * g + 4]
        for _ in range(rnd.randint(0, 6)):   # ranking etc. happens here
            yield
        prefix = 0
        while True:
            e = tcur & 3
            w = cur[e]
            while (w >> 30) == 0 or ((w >> 22) & 255) != EP:
                yield
                g = tcur >> 2
                cur = part[4 * g:4 * g + 4]
                w = cur[e]
            if (w >> 30) == 3:
                while incl[tcur] is None:
                    yield
                prefix += incl[tcur]
                break
            prefix += w & 0x3FFFFF
            if tcur == 0:
                break
            tcur -= 1
            if (tcur & 3) == 3:
                yield
                g = tcur >> 2
                cur = part[4 * g:4 * g + 4]
        result[t] = prefix
        # inclusive publish: flag possibly BEFORE the value becomes visible (reordered stores) -> reader must wait
        if rnd.random() < 0.5:
            incl[t] = prefix + counts[t]
            yield
            part[t] = word(3, counts[t])
        else:
            part[t] = word(3, counts[t])
            yield
            yield
            incl[t] = prefix + counts[t]

    # tiles start in ticket order but with at most `conc` in flight (persistent CTAs)
    conc = rnd.randint(1, 12)
    nxt = 0
    live = []
    steps = 0
    while nxt < T or live:
        while len(live) < conc and nxt < T:
            live.append(tile(nxt)); nxt += 1
        i = rnd.randrange(len(live))
        try:
            next(live[i])
        except StopIteration:
            live.pop(i)
        steps += 1
        assert steps < 10_000_000, "livelock"
    exact = 0
    for t in range(T):
        assert result[t] == exact, (seed, t, result[t], exact)
        exact += counts[t]



@pytest.mark.parametrize("seed", range(0, 200, 10))
def test_compact_descriptor_lookback_protocol(seed):
    for s in range(seed, seed + 10):
        run(s, T=random.Random(s).randint(1, 120))
