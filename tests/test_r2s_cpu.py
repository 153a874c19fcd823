"""read2sdbg (A12): the device code's __host__ __device__ building blocks, run on the CPU through the self-test hooks of
libmhb and compared with the oracle (which tests/test_oracle_r2s.py pins against the reference binary): stage-1 records
in bucket input order, stage-2 items, kmlib::kmsort's permutation, and the whole of stage 1 + the mercy step on a
fixed-length library (records -> stable bucket partition -> kmsort -> Lv2Postprocess -> mercy), ending in the same
solid-edge bits and multiplicity histogram as the oracle.  No GPU, no kernels: the kernels wrap exactly these functions."""
import ctypes as C

import numpy as np
import pytest

from megahit_b200 import formats as F
from megahit_b200 import lib, synth
from oracle import oracle as O
from test_oracle_r2s import R2S, r2s_reads


def oracle_s1_records(reads, r, k, base_off):
    L_ = O.lib()
    nw = lib.r2s_s1_key_words(k)
    Ln = int(reads.len[r])
    out = np.zeros((max(Ln - k + 4, 1), nw + 2), np.uint32)
    w = reads.words[int(reads.word_off[r]):int(reads.word_off[r + 1])]
    L_.mhbo_s1_read_records.restype = C.c_uint
    n = L_.mhbo_s1_read_records(C.c_void_p(w.ctypes.data), C.c_uint(Ln), C.c_uint(k), C.c_uint64(base_off),
                                C.c_void_p(out.ctypes.data))
    return out[:n]


@pytest.mark.parametrize("lib_name,k", [("golden/syn150_k27", 27), ("golden/syn150_k27", 21), ("golden/syn150_k27", 31),
                                        ("golden/syn150_k27", 59), ("golden/synvar_k21_m3", 25), ("golden/tandem_k27", 28),
                                        ("golden/polya_k27", 27), ("golden_kmax/syn300_k255", 199)])
def test_stage1_records_match_oracle(lib_name, k):
    reads = O.unpack_bin(r2s_reads(lib_name), reverse=True)
    rng = np.random.default_rng(k)
    base = 0
    bases = np.concatenate([[0], np.cumsum(reads.len.astype(np.uint64))])
    for r in rng.choice(reads.n, size=min(reads.n, 40), replace=False):
        Ln = int(reads.len[r])
        if Ln < k + 1:
            continue
        base = int(bases[r])
        want = oracle_s1_records(reads, int(r), k, base)
        w = reads.words[int(reads.word_off[r]):int(reads.word_off[r + 1])]
        assert len(want) == Ln - k + 4
        for e in range(len(want)):
            got = lib.selftest_r2s_s1_record(w, Ln, k, e, base)
            assert (got == want[e]).all(), (lib_name, k, int(r), e)


@pytest.mark.parametrize("lib_name,k", [("golden/syn150_k27", 27), ("golden/syn150_k27", 21), ("golden/syn150_k27", 39),
                                        ("golden/tandem_k27", 27), ("golden/polya_k27", 27), ("golden_kmax/syn300_k255", 255)])
def test_stage2_items_match_oracle(lib_name, k):
    reads = O.unpack_bin(r2s_reads(lib_name), reverse=True)
    L_ = O.lib()
    W1, W = (2 * k + 4 + 31) // 32, lib.s2s_record_words(k)
    rng = np.random.default_rng(k + 1)
    for r in rng.choice(reads.n, size=min(reads.n, 12), replace=False):
        Ln = int(reads.len[r])
        if Ln < k + 1:
            continue
        w = reads.words[int(reads.word_off[r]):int(reads.word_off[r + 1])]
        for i in list(range(0, Ln - k, 7)) + [Ln - k - 1]:
            for strand in (0, 1):
                for t in (0, 1, 2):
                    ref = np.zeros(W1, np.uint32)
                    pal = C.c_int()
                    L_.mhbo_s2_record(C.c_void_p(w.ctypes.data), C.c_uint(k), C.c_uint(i), C.c_uint(strand), C.c_uint(t),
                                      C.c_void_p(ref.ctypes.data), C.byref(pal))
                    got, gpal = lib.selftest_r2s_item(w, Ln, k, i, strand, t)
                    assert gpal == pal.value
                    # same characters; flags nondollar<<3|prev (low 4 bits of word W1-1) moved to bits 19..16 of word W-1
                    flags = int(ref[W1 - 1]) & 15
                    chars = ref.copy()
                    chars[W1 - 1] &= np.uint32(0xFFFFFFF0)
                    exp = np.zeros(W, np.uint32)
                    exp[:W1] = chars
                    exp[W - 1] |= np.uint32((flags << 16) | 0xFFFF)
                    assert (got == exp).all(), (lib_name, k, int(r), i, strand, t)


def oracle_kmsort(recs, nw):
    recs = np.ascontiguousarray(recs, np.uint32).copy()
    O.lib().mhbo_kmsort(C.c_void_p(recs.ctypes.data), C.c_int64(len(recs)), C.c_uint(nw), C.c_uint(recs.shape[1]))
    return recs


@pytest.mark.parametrize("n,nw,distinct", [(1, 2, 1), (2, 2, 1), (64, 2, 5), (65, 2, 3), (300, 2, 4), (5000, 2, 40), (5000, 2, 5000),
                                           (20000, 1, 300), (3000, 3, 7), (4000, 5, 100), (70000, 2, 900), (1000, 15, 12)])
def test_kmsort_emulation_matches_oracle(n, nw, distinct):
    """records with FEW distinct keys (long runs of ties) and a payload that tells the ties apart"""
    rng = np.random.default_rng(n * 31 + nw)
    keys = rng.integers(0, 2 ** 32, size=(distinct, nw), dtype=np.uint64).astype(np.uint32)
    keys[:, 0] = (keys[:, 0] & 0xFFFF) | 0x12340000  # one bucket: the two leading bytes are constant
    if nw > 1 and distinct > 3:
        keys[: distinct // 2, 1:] = keys[0, 1:]  # many keys sharing all but a few bytes -> deep radix levels
        keys[: distinct // 2, nw - 1] = (keys[0, nw - 1] & 0xFFFFFF00) | rng.integers(0, 256, distinct // 2).astype(np.uint32)
    recs = np.zeros((n, nw + 2), np.uint32)
    recs[:, :nw] = keys[rng.integers(0, distinct, n)]
    recs[:, nw + 1] = np.arange(n, dtype=np.uint32)  # payload = input position
    want = oracle_kmsort(recs, nw)
    got = lib.selftest_kmsort(recs, nw)
    assert (got == want).all()
    # the shared-memory form: default staging, tiny staged ranges (deep ranges take the in-place walk), tiny tag capacity
    for cap, wcap in ((65535, 0), (65535, 70), (100, 0)):
        assert (lib.selftest_kmsort(recs, nw, smem=True, cap=cap, wcap=wcap) == want).all(), (cap, wcap)
    assert not (got[:, nw + 1] == np.sort(got[:, nw + 1])).all() or n <= 64 or distinct == 1  # it is not the stable order


def run_stage1_on_host(bin_words, n_reads, Lr, k, m, need_mercy, expect_big_buckets=False):
    """stage 1 + mercy step with the device code's host-callable pieces (fixed-length library)"""
    reads = O.unpack_bin(bin_words.tobytes(), reverse=True)
    nw = lib.r2s_s1_key_words(k)
    per = Lr - k + 4
    recs = np.zeros((n_reads * per, nw + 2), np.uint32)
    for r in range(n_reads):
        w = reads.words[int(reads.word_off[r]):int(reads.word_off[r + 1])]
        for e in range(per):
            recs[r * per + e] = lib.selftest_r2s_s1_record(w, Lr, k, e, r * Lr)
    order = np.argsort(recs[:, 0] >> 16, kind="stable")  # = two stable radix passes on the leading bytes
    recs = recs[order]
    bucket = recs[:, 0] >> 16
    bounds = np.searchsorted(bucket, np.arange(65537))
    if expect_big_buckets:
        assert np.diff(bounds).max() > 64 * 4
    n_bits = n_reads * Lr
    planes = np.zeros((5, n_bits // 32 + 2), np.uint32)  # is_solid, no_in, no_out, any, mercy
    counting = np.zeros(65536, np.int64)
    L_ = lib.load()
    for b in np.nonzero(np.diff(bounds))[0]:
        seg = lib.selftest_kmsort(recs[bounds[b]:bounds[b + 1]], nw)
        lib._check(L_.mhb_selftest_r2s_s1_group(seg.ctypes.data, len(seg), k, m, Lr, n_reads, int(need_mercy),
                                                planes[0].ctypes.data, planes[1].ctypes.data, planes[2].ctypes.data,
                                                planes[3].ctypes.data, counting.ctypes.data))
    n_mercy = 0
    if need_mercy:
        added = C.c_uint32()
        for r in range(n_reads):
            lib._check(L_.mhb_selftest_r2s_mercy_read(Lr, n_reads, r, k, planes[0].ctypes.data, planes[1].ctypes.data,
                                                      planes[2].ctypes.data, planes[3].ctypes.data, planes[4].ctypes.data,
                                                      C.byref(added)))
            n_mercy += added.value
        planes[0] |= planes[4]
    return planes[0], counting, n_mercy


@pytest.mark.parametrize("case", ["syn150", "lowcov", "deep"])
def test_stage1_and_mercy_on_host_match_oracle(case):
    if case == "syn150":
        b, n, Lr, k, m = np.frombuffer(r2s_reads("golden/syn150_k27"), np.uint32), 3000, 150, 27, 2
        b, n = b.reshape(n, -1)[:700].reshape(-1), 700
    elif case == "lowcov":
        b, n, Lr, k, m = np.frombuffer(r2s_reads("golden/lowcov_k21"), np.uint32), 400, 150, 21, 2
    else:  # tie classes far above the insertion-sort threshold: kmsort's permutation decides has_in / has_out
        n, Lr, k, m = 4000, 100, 27, 2
        b = synth.synth_reads(n, Lr, 1500, 0.01, seed=5).reshape(-1)
    reads = O.unpack_bin(np.ascontiguousarray(b).tobytes(), reverse=True)
    want = O.read2sdbg(reads, k, m, True, want_solid=True)
    solid, counting, n_mercy = run_stage1_on_host(np.ascontiguousarray(b), n, Lr, k, m, True, case == "deep")
    assert n_mercy == want["n_mercy"]
    assert (counting == want["counting"]).all()
    bits = np.unpackbits(want["is_solid"], bitorder="little")[: n * Lr]
    got = np.unpackbits(solid.view(np.uint8), bitorder="little")[: n * Lr]
    assert (got == bits).all()


@pytest.mark.parametrize("k", [9, 11, 15, 17, 23, 33, 45, 47, 49, 63, 65, 79, 95, 97, 111, 127, 129, 159, 191, 223, 237])
def test_stage1_records_every_key_width(k):
    """every instantiated key width (1 .. 15 words) of the stage-1 record builder against the oracle on random reads"""
    rng = np.random.default_rng(k)
    for Ln in (k + 1, k + 2, k + 17, 2 * k + 5):
        b = rng.integers(0, 4, Ln, dtype=np.uint8)
        if Ln > k + 3:
            b[3:3 + (k - 1)] = np.concatenate([b[3:3 + (k - 1) // 2], (3 - b[3:3 + (k - 1) // 2][::-1])])[: k - 1]  # near-palindrome
        reads = O.unpack_bin(F.pack_read(b).tobytes(), reverse=True)
        want = oracle_s1_records(reads, 0, k, 12345)
        w = reads.words[: int(reads.word_off[1])]
        assert len(want) == Ln - k + 4
        for e in range(len(want)):
            assert (lib.selftest_r2s_s1_record(w, Ln, k, e, 12345) == want[e]).all(), (k, Ln, e)
