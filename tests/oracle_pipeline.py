"""Test helper: the full count -> (mercy) -> seq2sdbg path on the ORACLE (tests only)."""
import os

import numpy as np

from megahit_b200 import formats as F
from oracle import oracle as O


def load_reads(case_dir):
    data = open(os.path.join(case_dir, "reads.lib.bin"), "rb").read()
    return O.unpack_bin(data, reverse=True)


def oracle_count(reads, k, m):
    c = O.count(reads, k, m)
    ids = O.cand_ids(c["first_0_out"], c["last_0_in"])
    c["cand_ids"] = ids
    c["cand_bytes"] = reads.bin_bytes(ids)
    c["counting_text"] = O.counting_text(c["counting"])
    return c


def oracle_sdbg_from_count(c, k, mercy=True):
    seqs, mult = O.edges_as_seqs(c["edges"], k)
    if mercy:
        cand = O.unpack_bin(c["cand_bytes"], reverse=False)
        me = O.gen_mercy(c["edges"], cand, k)
        if len(me):
            seqs = O.Seqs.concat([seqs, O.Seqs.from_fixed(me, k + 1)])
            mult = np.concatenate([mult, np.ones(len(me), np.uint16)])
        c["n_mercy"] = len(me)
    s = O.seq2sdbg(seqs, mult, k)
    s["stream"] = F.canonical_sdbg_from_arrays(s["bucket_items"], s["bucket_byte_off"], s["bytes"])
    return s


# ---- k > k_min: what SeqToSdbg::Initialize loads (seq_to_sdbg.cpp:359-528), restated in Python for the oracle ----
def _read_fasta(path):
    recs, name, seq = [], None, []
    if not os.path.exists(path):
        return recs
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith(">"):
            if name is not None:
                recs.append((name, "".join(seq)))
            name, seq = line[1:], []
        elif name is not None:
            seq.append(line)
    if name is not None:
        recs.append((name, "".join(seq)))
    return recs


def _contig_seqs(path, min_len, k_from, k_to, rows, mults):
    """contig_reader.h:52-119: min length, flag/multi parsed positionally from the comment, loop extension,
    sequences stored REVERSED (seq_to_sdbg.cpp:455)."""
    code = {c: i for i, c in enumerate("ACGT")}
    code.update({c.lower(): i for c, i in list(code.items())})
    code.update({"N": 2, "n": 2})
    for header, seq in _read_fasta(path):
        if len(seq) < min_len:
            continue
        comment = header.split(None, 1)[1] if len(header.split(None, 1)) > 1 else ""
        flag = ord(comment[5]) - ord("0") if len(comment) > 5 else 0
        mult = int(float(comment[13:].split()[0]) + 0.5) & 0xFFFF if len(comment) > 13 else 0
        if k_from < k_to and (flag & 2):
            if len(seq) < k_to + 1:
                continue
            seq = seq + seq[k_from:k_to]
        rows.append(np.array([code.get(c, 0) for c in reversed(seq)], np.uint8))
        mults.append(mult)


def load_chain_seqs(case_dir, k, k_from):
    info = F.parse_edges_info(os.path.join(case_dir, str(k)))
    edges = F.canonical_edges(os.path.join(case_dir, str(k)))
    seqs, mult = O.edges_as_seqs(edges, k) if len(edges) else (O.Seqs(np.zeros(0, np.uint32), np.zeros(1, np.uint64), np.zeros(0, np.uint32)), np.zeros(0, np.uint16))
    rows, mults = [], []
    _contig_seqs(os.path.join(case_dir, f"k{k_from}.contigs.fa"), k + 1, k_from, k, rows, mults)
    _contig_seqs(os.path.join(case_dir, f"k{k_from}.bubble_seq.fa"), k + 1, 0, 0, rows, mults)
    _contig_seqs(os.path.join(case_dir, f"k{k_from}.addi.fa"), k + 1, 0, 0, rows, mults)
    _contig_seqs(os.path.join(case_dir, f"k{k_from}.local.fa"), k + 1, 0, 0, rows, mults)
    words, off, lens = [], [0], []
    for r in rows:
        w = F.pack_reads_fixed(r[None, :])[0][1:]
        words.append(w)
        off.append(off[-1] + len(w))
        lens.append(len(r))
    cs = O.Seqs(np.concatenate(words) if words else np.zeros(0, np.uint32), np.array(off, np.uint64), np.array(lens, np.uint32))
    allseqs = O.Seqs.concat([seqs, cs])
    allmult = np.concatenate([mult, np.array(mults, np.uint16)])
    assert info.kmer_size == k
    return allseqs, allmult
