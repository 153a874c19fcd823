"""CPU checks of the record builders the kernels run (mhb_kernels.cuh is __host__ __device__; the
mhb_selftest_* hooks execute it on the host): single records against a per-base Python restatement of
the reference, and a whole simulated `count` (selftest records -> numpy sort on the advertised sort
bytes -> run-length) against the oracle."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from megahit_b200 import formats as F
from megahit_b200 import lib
from oracle_pipeline import load_reads, oracle_count


def pack(bases):
    return F.pack_reads_fixed(np.asarray(bases, np.uint8)[None, :])[0][1:]


def expect_count_record(orig, k, q):
    """kmer_counter.cpp:158-252 on the REVERSED read; q is the file-orientation position."""
    L, K1 = len(orig), k + 1
    rv = orig[::-1]
    p = L - K1 - q
    fwd = list(rv[p:p + K1])
    rc = [3 - b for b in fwd[::-1]]
    strand = 1 if rc < fwd else 0
    key = rc if strand else fwd
    prev = rv[p - 1] if p > 0 else 4
    nxt = rv[p + K1] if p + K1 < L else 4
    if strand:
        prev, nxt = (4 if nxt == 4 else 3 - nxt), (4 if prev == 4 else 3 - prev)
    wr = (2 * K1 + 6 + 31) // 32
    rec = np.zeros(wr, np.uint64)
    for i, b in enumerate(key):
        rec[i >> 4] |= np.uint64(int(b) << (30 - 2 * (i & 15)))
    rec[wr - 1] |= np.uint64((int(prev) << 3) | int(nxt))
    return rec.astype(np.uint32), strand


def expect_s2s_record(seq, k, strand, offset, mult):
    """seq_to_sdbg.cpp:630-700."""
    L = len(seq)
    nc = k - (1 if offset + k > L else 0)
    counting = mult if (offset > 0 and offset + k <= L) else 0
    if strand == 0:
        prev = 4 if offset == 0 else seq[offset - 1]
        chars = list(seq[offset:offset + nc])
    else:
        prev = 4 if offset == 0 else 3 - seq[L - 1 - offset + 1]
        off2 = max(0, L - 1 - offset - (k - 1))
        chars = [3 - b for b in seq[off2:off2 + nc][::-1]]
    w = (2 * k + 20 + 31) // 32
    rec = np.zeros(w, np.uint64)
    for i, b in enumerate(chars):
        rec[i >> 4] |= np.uint64(int(b) << (30 - 2 * (i & 15)))
    rec[w - 1] |= np.uint64((int(nc == k) << 19) | (int(prev) << 16) | max(0, 65535 - counting))
    return rec.astype(np.uint32)


@pytest.mark.parametrize("k", [9, 14, 15, 16, 21, 27, 28, 29, 31, 32, 47, 63, 64, 99, 141, 255])
def test_count_record_builder(k):
    rng = np.random.default_rng(k)
    for _ in range(40):
        L = int(rng.integers(k + 1, k + 40))
        orig = rng.integers(0, 4, L).astype(np.uint8)
        if rng.random() < 0.3:  # palindromic / low-complexity stress
            orig[:] = rng.integers(0, 4)
        words = pack(orig)
        for q in {0, L - k - 1, int(rng.integers(0, L - k))}:
            got, strand = lib.selftest_count_record(words, L, k, q)
            exp, estrand = expect_count_record(orig, k, q)
            assert strand == estrand and (got == exp).all(), (k, L, q)


@pytest.mark.parametrize("k", [9, 15, 16, 21, 22, 27, 29, 31, 32, 39, 59, 79, 99, 119, 141, 255])
def test_s2s_record_builder(k):
    rng = np.random.default_rng(1000 + k)
    for _ in range(30):
        L = int(rng.integers(k + 1, k + 30))
        seq = rng.integers(0, 4, L).astype(np.uint8)
        words = pack(seq)
        mult = int(rng.integers(0, 65536))
        for strand in (0, 1):
            for offset in {0, 1, L - k, L - k + 1, int(rng.integers(0, L - k + 2))}:
                got = lib.selftest_s2s_record(words, L, k, strand, offset, mult)
                exp = expect_s2s_record(seq, k, strand, offset, mult)
                assert (got == exp).all(), (k, L, strand, offset)


@pytest.mark.parametrize("name,k,m", [("toy_k21", 21, 2), ("tandem_k27", 28, 2), ("synvar_k31_m1", 31, 1)])
def test_simulated_count_matches_oracle(name, k, m):
    """records from the device code (run on the host) + a numpy LSD sort over mhb_count_sort_bytes +
    run-length counting reproduce the oracle's solid edges: pins record layout and sort-byte selection."""
    case = os.path.join(GOLDEN, name)
    bin_words = np.fromfile(os.path.join(case, "reads.lib.bin"), np.uint32)
    recs = []
    pos = 0
    while pos < len(bin_words):
        L = int(bin_words[pos])
        nw = (L + 15) // 16
        w = bin_words[pos + 1:pos + 1 + nw]
        for q in range(0, max(0, L - k)):
            recs.append(lib.selftest_count_record(w, L, k, q)[0])
        pos += 1 + nw
    wr = lib.count_record_words(k)
    recs = np.array(recs, np.uint32).reshape(-1, wr)
    order = np.arange(len(recs))
    for b in lib.count_sort_bytes(k):  # LSD: stable sort per byte, least significant first
        digit = (recs[order, wr - 1 - (b >> 2)] >> (8 * (b & 3))) & 255
        order = order[np.argsort(digit, kind="stable")]
    srt = recs[order].copy()
    srt[:, wr - 1] &= np.uint32(0xFFFFFFC0)
    head = np.ones(len(srt), bool)
    head[1:] = (srt[1:] != srt[:-1]).any(axis=1)
    starts = np.nonzero(head)[0]
    counts = np.diff(np.append(starts, len(srt)))
    solid = counts >= m
    wpe = lib.words_per_edge(k)
    edges = np.zeros((int(solid.sum()), wpe), np.uint32)
    kw = (2 * (k + 1) + 31) // 32
    edges[:, :min(kw, wpe)] = srt[starts[solid]][:, :min(kw, wpe)]
    edges[:, wpe - 1] |= np.minimum(counts[solid], 65535).astype(np.uint32)
    c = oracle_count(load_reads(case), k, m)
    assert (edges == c["edges"]).all() and len(edges) == c["n_solid"]


# ------------------------------------------------------------------------------------------------
# A13: the round planner of the out-of-core count stage (host logic, no GPU)
# ------------------------------------------------------------------------------------------------
def test_plan_rounds_covers_all_bytes_within_cap():
    from megahit_b200 import lib
    rng = np.random.default_rng(5)
    for trial in range(50):
        hist = rng.integers(0, 1000, size=256).astype(np.uint64)
        if trial % 5 == 0:
            hist[rng.integers(0, 256, size=200)] = 0  # sparse: most leading bytes absent
        cap = int(max(int(hist.max()), int(hist.sum()) // int(rng.integers(1, 40))))
        ranges = lib.plan_rounds(hist, cap)
        # contiguous, ascending, complete
        assert ranges[0][0] == 0 and ranges[-1][1] == 255
        for (a, b), (c, d) in zip(ranges[:-1], ranges[1:]):
            assert a <= b and c == b + 1
        sums = [int(hist[a:b + 1].sum()) for a, b in ranges]
        assert all(s <= cap for s in sums) and sum(sums) == int(hist.sum())
        # greedy: a range could not have taken the next byte as well
        for (a, b), s in zip(ranges[:-1], sums[:-1]):
            assert s + int(hist[b + 1]) > cap


def test_plan_rounds_single_round_and_oversized_byte():
    from megahit_b200 import lib
    hist = np.full(256, 10, np.uint64)
    assert lib.plan_rounds(hist, 2560) == [(0, 255)]
    assert lib.plan_rounds(hist, 10) == [(i, i) for i in range(256)]
    hist[7] = 11
    with pytest.raises(lib.MhbError, match="more than one round can take"):
        lib.plan_rounds(hist, 10)
    assert lib.plan_rounds(np.zeros(256, np.uint64), 1) == [(0, 255)]


# ------------------------------------------------------------------------------------------------
# the rolling record builder (4 consecutive positions per call) against the position-by-position builder
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [16, 20, 21, 26, 27, 28, 29, 31])
def test_rolling_count_record_builder_matches(k):
    rng = np.random.default_rng(77 + k)
    wr = lib.count_record_words(k)
    for trial in range(25):
        L = int(rng.integers(k + 1, k + 140))
        orig = rng.integers(0, 4, L).astype(np.uint8)
        if trial % 6 == 0:
            orig[:] = rng.integers(0, 4)  # homopolymer: forward == reverse complement ties, strand flips
        if trial % 9 == 0:
            h = orig[: L // 2].copy()
            orig[L - len(h):] = 3 - h[::-1]  # reverse-complement palindrome across the read
        words = pack(orig)
        for q in range(0, L - k):
            rec, strand = lib.selftest_count_records_roll(words, L, k, q)
            for j in range(4):
                if q + j + k + 1 <= L:
                    e, es = lib.selftest_count_record(words, L, k, q + j)
                    ev = (int(e[0]) << 32) | int(e[1])
                    if wr == 3:
                        ev |= int(e[2])  # 3-word record: key in words 0-1, prev/next in word 2
                    assert ev == int(rec[j]) and es == int(strand[j]), (k, L, q, j)


def test_plan_rounds16_cuts_an_oversized_leading_byte_on_bucket_ids():
    """A13 planner used by the host rounds: a leading byte above the cap (A-prefix skew, poly-A) is cut on its second
    byte = on the reference's 16-bit bucket ids (base_engine.cpp:254-281), everything else on whole leading bytes"""
    rng = np.random.default_rng(5)
    hist = rng.integers(0, 1000, 256).astype(np.uint64)
    sub = np.zeros((256, 256), np.uint64)
    for b in (0, 17):  # two hot bytes
        sub[b] = rng.integers(0, 400, 256).astype(np.uint64)
        hist[b] = sub[b].sum()
    cap = 3000
    assert hist[0] > cap and hist[17] > cap
    ranges = lib.plan_rounds16(hist, sub, cap)
    # tiling of the bucket ids, ascending, every range within the cap
    assert ranges[0][0] == 0 and ranges[-1][1] == 65535
    per_bucket = np.zeros(65536, np.uint64)
    for b in range(256):
        if hist[b] > cap:
            per_bucket[b * 256:(b + 1) * 256] = sub[b]
        else:
            per_bucket[b * 256] = hist[b]  # unit = the whole byte: a cut may only fall on a byte boundary
    for (lo, hi), nxt in zip(ranges, ranges[1:] + [None]):
        assert lo <= hi and int(per_bucket[lo:hi + 1].sum()) <= cap
        if nxt is not None:
            assert nxt[0] == hi + 1
            assert (nxt[0] & 255) == 0 or hist[nxt[0] >> 8] > cap
    # without oversized bytes it degenerates to the byte planner
    h2 = np.full(256, 10, np.uint64)
    assert lib.plan_rounds16(h2, None, 25) == [(lo << 8, (hi << 8) | 255) for lo, hi in lib.plan_rounds(h2, 25)]
    assert lib.plan_rounds16(np.zeros(256, np.uint64), None, 1) == [(0, 65535)]
    # a single bucket above the cap is reported, and an oversized byte without its second-level histogram too
    sub2 = np.zeros((256, 256), np.uint64)
    sub2[3, 9] = 50
    h3 = np.zeros(256, np.uint64)
    h3[3] = 50
    with pytest.raises(lib.MhbError, match="bucket 0x0309"):
        lib.plan_rounds16(h3, sub2, 10)
    with pytest.raises(lib.MhbError, match="leading byte 0x03"):
        lib.plan_rounds16(h3, None, 10)
