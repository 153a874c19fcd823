"""k > k_min: seq2sdbg over contigs (flags, multiplicities, loop extension), bubble/addi/local contigs and
UNSORTED iterative edges, against fixtures captured from the reference's own multi-k run (oracle/gen_golden.py chain)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from megahit_b200 import formats as F

CASES = ["chain_syn150", "chain_toy"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_on_contig_inputs(name):
    import oracle_pipeline as OP
    from oracle import oracle as O
    case = os.path.join(GOLDEN, name)
    g = json.load(open(os.path.join(case, "chain.json")))
    seqs, mult = OP.load_chain_seqs(case, g["k"], g["k_from"])
    s = O.seq2sdbg(seqs, mult, g["k"])
    assert int(s["n_items"]) == g["sdbg_items"] and int(s["bucket_tips"].sum()) == g["sdbg_tips"]
    stream = F.canonical_sdbg_from_arrays(s["bucket_items"], s["bucket_byte_off"], s["bytes"])
    assert F.sha256(stream) == g["sdbg_sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_seq2sdbg_subcommand_on_contig_inputs(name, tmp_path):
    from megahit_b200 import lib
    case = os.path.join(GOLDEN, name)
    g = json.load(open(os.path.join(case, "chain.json")))
    k, kf = g["k"], g["k_from"]
    p = str(tmp_path / str(k))
    lib.seq2sdbg_run(p, k=k, k_from=kf, input_prefix=os.path.join(case, str(k)),
                     contig=os.path.join(case, f"k{kf}.contigs.fa"), bubble=os.path.join(case, f"k{kf}.bubble_seq.fa"),
                     addi_contig=os.path.join(case, f"k{kf}.addi.fa"), local_contig=os.path.join(case, f"k{kf}.local.fa"),
                     need_mercy=False, host_mem=1e9, num_cpu_threads=2)
    info, stream, table = F.canonical_sdbg(p)
    assert info.k == g["sdbg_k"] and info.words_per_tip_label == g["sdbg_words_per_tip_label"]
    assert int(table[:, 0].sum()) == g["sdbg_items"] and int(table[:, 1].sum()) == g["sdbg_tips"]
    assert int(table[:, 2].sum()) == g["sdbg_large_mul"]
    assert F.sha256(stream) == g["sdbg_sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_s2s_host_variable_length_matches_oracle(name):
    import oracle_pipeline as OP
    from megahit_b200 import lib
    from oracle import oracle as O
    case = os.path.join(GOLDEN, name)
    g = json.load(open(os.path.join(case, "chain.json")))
    seqs, mult = OP.load_chain_seqs(case, g["k"], g["k_from"])
    os_ = O.seq2sdbg(seqs, mult, g["k"])
    gs = lib.s2s_host(seqs.words, seqs.word_off, seqs.len, mult, g["k"])
    assert gs["bytes"] == os_["bytes"] and gs["n_items"] == os_["n_items"]
    assert (gs["w_count"] == os_["w_count"]).all()
