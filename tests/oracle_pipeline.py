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
