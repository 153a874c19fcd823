"""Pins the oracle (oracle/mhb_oracle.c) against fixtures minted by the UNMODIFIED reference binary
(oracle/gen_golden.py).  CPU only."""
import os

import pytest

from conftest import GOLDEN, golden_cases
from megahit_b200 import formats as F
from oracle_pipeline import load_reads, oracle_count, oracle_sdbg_from_count


@pytest.mark.parametrize("name,k,m,gold", golden_cases())
def test_oracle_matches_reference(name, k, m, gold):
    reads = load_reads(os.path.join(GOLDEN, name))
    c = oracle_count(reads, k, m)
    assert c["n_solid"] == gold["n_solid"]
    assert c["words_per_edge"] == gold["words_per_edge"] or gold["n_solid"] == 0
    assert F.sha256(c["edges"].tobytes()) == gold["edges_sha256"]
    assert F.sha256(c["cand_bytes"]) == gold["cand_sha256"]
    assert F.sha256(c["counting_text"]) == gold["counting_sha256"]
    s = oracle_sdbg_from_count(c, k, mercy=True)
    assert int(s["n_items"]) == gold["sdbg_items"]
    assert int(s["bucket_tips"].sum()) == gold["sdbg_tips"]
    assert int(s["bucket_large_mul"].sum()) == gold["sdbg_large_mul"]
    assert s["words_per_tip_label"] == gold["sdbg_words_per_tip_label"]
    assert F.sha256(s["stream"]) == gold["sdbg_sha256"]
