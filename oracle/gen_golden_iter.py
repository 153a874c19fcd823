#!/usr/bin/env python
"""Mint fixtures for `megahit_core iterate` with the UNMODIFIED reference: its Python driver is run on a
seeded synthetic library with --k-list 21,29,49,77 --keep-tmp-files (steps 8, 20, 28 = the largest allowed); for every
step the contigs / bubble files iterate reads are committed together with the digest of the set of iterative edges it
wrote (P.edges.0 is in hash-table order: the canonical form is the ascending set of records).  The reference binary is
then re-run directly with 1 and 4 threads to confirm that the set does not depend on scheduling.

    python oracle/gen_golden_iter.py      ->  tests/golden_iter/
"""
from __future__ import annotations

import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from megahit_b200 import formats as F  # noqa: E402
from megahit_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "megahit_core_ref")
DRIVER = os.path.join(ROOT, "oracle", "_ref", "megahit")
OUT = os.path.join(ROOT, "tests", "golden_iter")
SYNTH = dict(n_reads=6000, read_len=150, genome_len=30000, err=0.01, seed=31)  # + repeats, see repeat_reads()
KLIST = [21, 29, 49, 77]


def edge_set_digest(prefix):
    info = open(prefix + ".edges.info").read().split()
    W, n = int(info[3]), int(info[9])
    e = np.fromfile(prefix + ".edges.0", np.uint32).reshape(-1, W)
    assert len(e) == n and int(info[11]) == 0
    u = np.unique(e, axis=0)  # lexicographic by words = ascending records
    assert len(u) == n, "the reference wrote a duplicate"
    return {"n_edges": int(n), "words_per_edge": W, "kmer_size": int(info[1]), "edges_sha256": F.sha256(u.tobytes()),
            "all_mult_zero": bool(((u[:, -1] & 0xFFFF) == 0).all())}


def repeat_reads():
    """a genome with repeats of 35, 60 and 90 bases (longer than k = 21 / 29, some shorter than 49 / 77): contigs break at
    the repeats for the small k, and reads spanning them give iterative edges at every step"""
    rng = np.random.default_rng(SYNTH["seed"])
    g = rng.integers(0, 4, SYNTH["genome_len"], dtype=np.uint8)
    for rl, copies in ((35, 12), (60, 12), (90, 8)):
        rep = rng.integers(0, 4, rl, dtype=np.uint8)
        for p in rng.choice(len(g) - rl, copies, replace=False):
            g[p:p + rl] = rep
    n, L = SYNTH["n_reads"], SYNTH["read_len"]
    pos = rng.integers(0, len(g) - L + 1, size=n)
    b = g[pos[:, None] + np.arange(L)[None, :]]
    rc = rng.integers(0, 2, size=n).astype(bool)
    b[rc] = 3 - b[rc][:, ::-1]
    e = rng.random(b.shape) < SYNTH["err"]
    b[e] = (b[e] + rng.integers(1, 4, size=int(e.sum()), dtype=np.uint8)) & 3
    return F.pack_reads_fixed(b)


def main():
    os.makedirs(OUT, exist_ok=True)
    res = {"synth": SYNTH, "steps": []}
    with tempfile.TemporaryDirectory() as tmp:
        b = repeat_reads()
        b.tofile(os.path.join(OUT, "reads.lib.bin"))
        with open(f"{tmp}/reads.fa", "w") as f:
            for i, row in enumerate(b):
                L, w = int(row[0]), row[1:]
                f.write(f">r{i}\n" + "".join("ACGT"[(int(w[j >> 4]) >> (30 - 2 * (j & 15))) & 3] for j in range(L)) + "\n")
        os.makedirs(f"{tmp}/bin")
        shutil.copy(DRIVER, f"{tmp}/bin/megahit")
        os.symlink(REF, f"{tmp}/bin/megahit_core")
        subprocess.run([sys.executable, f"{tmp}/bin/megahit", "-r", f"{tmp}/reads.fa", "--k-list", ",".join(map(str, KLIST)),
                        "-o", f"{tmp}/out", "--keep-tmp-files", "-t", "4"], check=True, capture_output=True)
        # the driver's read library must be the synthetic `.bin` image itself (the tests regenerate it from the seed)
        lib_bin = np.fromfile(f"{tmp}/out/tmp/reads.lib.bin", np.uint32)
        assert (lib_bin == b.reshape(-1)).all(), "buildlib changed the reads"
        ic = f"{tmp}/out/intermediate_contigs"
        for k, kn in zip(KLIST[:-1], KLIST[1:]):
            for fn in (f"k{k}.contigs.fa", f"k{k}.bubble_seq.fa"):
                shutil.copy(f"{ic}/{fn}", OUT)
            d = edge_set_digest(f"{tmp}/out/tmp/k{kn}/{kn}")
            assert d["all_mult_zero"] and d["kmer_size"] == kn
            for t in (1, 4):
                p = f"{tmp}/it{k}_{t}"
                subprocess.run([REF, "iterate", "-c", f"{ic}/k{k}.contigs.fa", "-b", f"{ic}/k{k}.bubble_seq.fa", "-t", str(t),
                                "-k", str(k), "-s", str(kn - k), "-o", p, "-r", f"{tmp}/out/tmp/reads.lib.bin"],
                               check=True, capture_output=True)
                assert edge_set_digest(p) == d, "iterate depends on the thread count"
            d.update({"k": k, "step": kn - k})
            res["steps"].append(d)
            print(k, kn, d["n_edges"])
    # the two chain fixtures of tests/golden (k = 21 -> 29), whose iterate output is already committed
    for name, lib in (("chain_syn150", "syn150_k27"), ("chain_toy", "toy_k21")):
        d = edge_set_digest(os.path.join(ROOT, "tests", "golden", name, "29"))
        d.update({"k": 21, "step": 8, "chain": name, "lib": lib})
        res["steps"].append(d)
        print(name, d["n_edges"])
    json.dump(res, open(os.path.join(OUT, "iter.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
