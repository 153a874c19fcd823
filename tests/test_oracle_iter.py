"""Pins the oracle's `iterate` restatement (oracle/mhb_oracle_iter.c) against sets of iterative edges written by the
UNMODIFIED reference (oracle/gen_golden_iter.py -> tests/golden_iter/, and the k = 21 -> 29 edges of the two chain
fixtures under tests/golden/).  CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from megahit_b200 import formats as F
from oracle import oracle as O

ITER_DIR = os.path.join(ROOT, "tests", "golden_iter")
ITER = json.load(open(os.path.join(ITER_DIR, "iter.json")))


def read_fasta(path):
    recs, name, seq = [], None, []
    if os.path.exists(path):
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


def contig_seqs(paths):
    """what AsyncContigReader hands to the flank index (async_sequence_reader.h:82-101): file orientation, contigs
    flagged kStandalone (1) or kLoop (2) discarded (contig_reader.h:64-69)"""
    code = {c: i for i, c in enumerate("ACGT")}
    code.update({c.lower(): i for c, i in list(code.items())})
    code.update({"N": 2, "n": 2})
    words, off, lens = [], [0], []
    for p in paths:
        for header, seq in read_fasta(p):
            comment = header.split(None, 1)[1] if len(header.split(None, 1)) > 1 else ""
            flag = ord(comment[5]) - ord("0") if len(comment) > 5 else 0
            if flag & 3 or not seq:
                continue
            w = F.pack_reads_fixed(np.array([code.get(c, 0) for c in seq], np.uint8)[None, :])[0][1:]
            words.append(w)
            off.append(off[-1] + len(w))
            lens.append(len(seq))
    return O.Seqs(np.concatenate(words) if words else np.zeros(0, np.uint32), np.array(off, np.uint64), np.array(lens, np.uint32))


def iter_inputs(step):
    """(contig files, `.bin` image) of a fixture step"""
    k = step["k"]
    if "chain" in step:
        d = os.path.join(GOLDEN, step["chain"])
        data = open(os.path.join(GOLDEN, step["lib"], "reads.lib.bin"), "rb").read()
    else:
        d = ITER_DIR
        data = open(os.path.join(ITER_DIR, "reads.lib.bin"), "rb").read()
    return [os.path.join(d, f"k{k}.contigs.fa"), os.path.join(d, f"k{k}.bubble_seq.fa")], data


def iter_cases():
    return [pytest.param(s, id=f"{s.get('chain', 'repeats')}-k{s['k']}+{s['step']}") for s in ITER["steps"]]


@pytest.mark.parametrize("step", iter_cases())
def test_oracle_iterate_matches_reference(step):
    files, data = iter_inputs(step)
    reads = O.unpack_bin(data, reverse=False)
    edges, _ = O.iterate(contig_seqs(files), reads, step["k"], step["step"])
    assert len(edges) == step["n_edges"] and edges.shape[1] == step["words_per_edge"]
    assert F.sha256(edges.tobytes()) == step["edges_sha256"]
