"""Pins the oracle's read2sdbg restatement (oracle/mhb_oracle_r2s.c, kmsort tie order included) against fixtures minted
by the UNMODIFIED reference binary (oracle/gen_golden_r2s.py -> tests/golden_r2s/r2s.json).  CPU only."""
import json
import os

import pytest

from conftest import ROOT
from megahit_b200 import formats as F
from megahit_b200 import synth
from oracle import oracle as O

R2S = json.load(open(os.path.join(ROOT, "tests", "golden_r2s", "r2s.json")))
_cache = {}


def r2s_reads(lib):
    """The read library of a fixture run: committed `.bin` image, or a seeded synthetic one (digest-only fixture)."""
    if lib not in _cache:
        if lib.startswith("synth:"):
            a = R2S["synth"][lib[6:]]
            data = synth.synth_reads(a["n_reads"], a["read_len"], a["genome_len"], a["err"], seed=a["seed"]).tobytes()
        else:
            data = open(os.path.join(ROOT, "tests", lib, "reads.lib.bin"), "rb").read()
        _cache[lib] = data
    return _cache[lib]


def r2s_cases(max_reads=None):
    out = []
    for r in R2S["runs"]:
        out.append(pytest.param(r, id=f"{r['lib'].split('/')[-1]}-k{r['k']}-m{r['m']}-mercy{r['mercy']}"))
    return out


def check_against_gold(s, gold):
    assert int(s["n_items"]) == gold["sdbg_items"]
    assert int(s["bucket_tips"].sum()) == gold["sdbg_tips"]
    assert int(s["bucket_large_mul"].sum()) == gold["sdbg_large_mul"]
    assert s["words_per_tip_label"] == gold["sdbg_words_per_tip_label"]
    stream = F.canonical_sdbg_from_arrays(s["bucket_items"], s["bucket_byte_off"], s["bytes"])
    assert F.sha256(stream) == gold["sdbg_sha256"]


@pytest.mark.parametrize("gold", r2s_cases())
def test_oracle_read2sdbg_matches_reference(gold):
    reads = O.unpack_bin(r2s_reads(gold["lib"]), reverse=True)
    s = O.read2sdbg(reads, gold["k"], gold["m"], bool(gold["mercy"]))
    assert s["n_mercy"] == gold["n_mercy"]
    if gold["m"] > 1:
        assert F.sha256(O.counting_text(s["counting"])) == gold["counting_sha256"]
    check_against_gold(s, gold)
