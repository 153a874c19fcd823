"""The oracle at the reference's largest k (255: 17-word count records, 16-word tip labels) and at k=199, against
fixtures minted by the unmodified reference (oracle/gen_golden.py kmax -> tests/golden_kmax/).  CPU only."""
import json
import os

import pytest

from conftest import ROOT
from megahit_b200 import formats as F
from oracle_pipeline import load_reads, oracle_count, oracle_sdbg_from_count

KMAX = os.path.join(ROOT, "tests", "golden_kmax")


def kmax_cases():
    out = []
    for name in sorted(os.listdir(KMAX)):
        g = json.load(open(os.path.join(KMAX, name, "golden.json")))
        for k, v in sorted(g["by_k"].items(), key=lambda kv: int(kv[0])):
            out.append(pytest.param(name, int(k), g["m"], v, id=f"{name}-k{k}"))
    return out


@pytest.mark.parametrize("name,k,m,gold", kmax_cases())
def test_oracle_matches_reference_at_large_k(name, k, m, gold):
    reads = load_reads(os.path.join(KMAX, name))
    c = oracle_count(reads, k, m)
    assert c["n_solid"] == gold["n_solid"] and c["words_per_edge"] == gold["words_per_edge"]
    assert F.sha256(c["edges"].tobytes()) == gold["edges_sha256"]
    assert F.sha256(c["cand_bytes"]) == gold["cand_sha256"]
    assert F.sha256(c["counting_text"]) == gold["counting_sha256"]
    s = oracle_sdbg_from_count(c, k, mercy=True)
    assert int(s["n_items"]) == gold["sdbg_items"] and int(s["bucket_tips"].sum()) == gold["sdbg_tips"]
    assert s["words_per_tip_label"] == gold["sdbg_words_per_tip_label"]
    assert F.sha256(s["stream"]) == gold["sdbg_sha256"]
