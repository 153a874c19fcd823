"""`iterate` (N2): the device code's __host__ __device__ building blocks (flank records, flank search behind the prefix
table, read marking, edge emission) driven serially on the host (mhb_selftest_iterate) against the reference-minted
fixtures and the oracle.  No GPU: the kernels wrap exactly these functions."""
import numpy as np
import pytest

from megahit_b200 import formats as F
from megahit_b200 import lib
from oracle import oracle as O
from test_oracle_iter import contig_seqs, iter_cases, iter_inputs


@pytest.mark.parametrize("step", iter_cases())
def test_host_mirror_matches_reference(step):
    files, data = iter_inputs(step)
    cs = contig_seqs(files)
    b = np.frombuffer(data, np.uint32)
    n_reads = O.unpack_bin(data, reverse=False).n
    g = lib.iterate_host(cs.words, cs.word_off, cs.len, b, n_reads, step["k"], step["step"], selftest=True)
    assert g["n_edges"] == step["n_edges"] and g["edges"].shape[1] == step["words_per_edge"]
    assert F.sha256(g["edges"].tobytes()) == step["edges_sha256"]
    want, aligned = O.iterate(cs, O.unpack_bin(data, reverse=False), step["k"], step["step"])
    assert (g["edges"] == want).all() and g["n_aligned_reads"] == aligned


def test_host_mirror_duplicate_flanks_and_short_contigs():
    """contigs sharing a flank (the longer extension wins), a contig of exactly k+1 bases (one strand only), a
    palindromic flank, contigs shorter than k+1, reads shorter than k+step+1"""
    rng = np.random.default_rng(3)
    k, step = 21, 8
    g = rng.integers(0, 4, 400, dtype=np.uint8)
    pal = np.concatenate([g[:11], 3 - g[:11][::-1]])  # 22 = k+1 bases, its own reverse complement
    # A = g[0:100] and B = g[79:200] overlap by k bases, as unitigs do: a read across the junction is covered by A's end
    # flank + extension and B's start flank + extension -> a run of 16 marked positions >= step + 1
    contigs = [g[0:100], g[79:200], g[79:130], g[79:101], g[100:122], pal, g[200:215], np.concatenate([pal, g[300:330]]),
               g[250:340], 3 - g[319:400][::-1]]
    rows = [F.pack_reads_fixed(c[None, :])[0][1:] for c in contigs]
    off = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.uint64)
    cs = O.Seqs(np.concatenate(rows), off, np.array([len(c) for c in contigs], np.uint32))
    reads = [g[20:170], 3 - g[20:170][::-1], g[60:140], g[5:30], g[230:380], 3 - g[240:390][::-1],
             np.concatenate([pal, g[300:380]])]
    data = b"".join(F.pack_read(r).tobytes() for r in reads)
    want, aligned = O.iterate(cs, O.unpack_bin(data, reverse=False), k, step)
    got = lib.iterate_host(cs.words, cs.word_off, cs.len, np.frombuffer(data, np.uint32), len(reads), k, step, selftest=True)
    assert len(want) > 0 and (got["edges"] == want).all() and got["n_aligned_reads"] == aligned


@pytest.mark.parametrize("k,step,seed", [(15, 2, 1), (31, 28, 2), (47, 16, 3), (63, 2, 4), (99, 28, 5), (141, 28, 6), (211, 28, 7), (239, 4, 8)])
def test_host_mirror_matches_oracle_on_random_contig_sets(k, step, seed):
    """every register-width class of the kernels (2 / 4 / 8 / 17 words) up to the widest flank record (k + 1 = 240):
    contigs = a repeat-rich genome cut at arbitrary places with k-base overlaps, reads of mixed lengths and strands"""
    rng = np.random.default_rng(seed)
    G = 6000
    g = rng.integers(0, 4, G, dtype=np.uint8)
    for rl in (k + 3, k + step // 2, 2 * k):
        rep = rng.integers(0, 4, rl, dtype=np.uint8)
        for p in rng.choice(G - rl, 3, replace=False):
            g[p:p + rl] = rep
    cuts = np.sort(rng.choice(np.arange(k + 2, G - k - 2), 14, replace=False))
    bounds = [0] + list(cuts) + [G]
    contigs = [g[max(0, a - k):b] for a, b in zip(bounds[:-1], bounds[1:])]
    contigs += [3 - c[::-1] for c in contigs[::3]]                      # some given on the other strand
    contigs += [g[100:100 + k + 1], g[300:300 + k]]                     # exactly k+1 bases; too short
    rows = [F.pack_reads_fixed(c[None, :])[0][1:] for c in contigs]
    off = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.uint64)
    cs = O.Seqs(np.concatenate(rows), off, np.array([len(c) for c in contigs], np.uint32))
    reads = []
    for _ in range(300):
        Lr = int(rng.integers(k, 3 * k + 2 * step + 40))
        p = int(rng.integers(0, G - Lr))
        r = g[p:p + Lr].copy()
        e = rng.random(Lr) < 0.004
        r[e] = (r[e] + 1) & 3
        reads.append(3 - r[::-1] if rng.integers(0, 2) else r)
    data = b"".join(F.pack_read(r).tobytes() for r in reads)
    want, aligned = O.iterate(cs, O.unpack_bin(data, reverse=False), k, step)
    got = lib.iterate_host(cs.words, cs.word_off, cs.len, np.frombuffer(data, np.uint32), len(reads), k, step, selftest=True)
    assert len(want) > 0, "vacuous case"
    assert got["n_edges"] == len(want) and (got["edges"] == want).all() and got["n_aligned_reads"] == aligned
