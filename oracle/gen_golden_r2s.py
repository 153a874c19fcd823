#!/usr/bin/env python
"""Mint read2sdbg (1-pass route) fixtures by running the UNMODIFIED reference binary
(oracle/_ref/megahit_core_ref read2sdbg) on the read libraries already committed under tests/golden*/ and on
seeded synthetic libraries that are regenerated at test time (only their digests are committed).

    python oracle/gen_golden_r2s.py

Output: tests/golden_r2s/r2s.json = list of runs {lib, k, m, mercy, digests}.  For every run the reference is
executed with 1 and 4 threads and with --mem_flag 0 and 1 (different Lv1 pass boundaries) and must give identical
canonical streams - that is the determinism the CUDA path has to reproduce, tie order of kmsort included
(read_to_sdbg_s1.cpp:393-401 reads prev/next of the FIRST item of a (k-1)-mer group for all its members).
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from megahit_b200 import formats as F  # noqa: E402
from megahit_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "megahit_core_ref")
OUT = os.path.join(ROOT, "tests", "golden_r2s")

# (library, k, m, need_mercy).  Libraries: a directory under tests/ holding reads.lib.*, or "synth:<name>"
RUNS = [
    ("golden/toy_k21", 21, 2, 1), ("golden/toy_k21", 21, 1, 0), ("golden/toy_k21", 29, 2, 0),
    ("golden/syn150_k27", 27, 2, 1), ("golden/syn150_k27", 27, 2, 0), ("golden/syn150_k27", 27, 1, 1),
    ("golden/syn150_k27", 21, 3, 1), ("golden/syn150_k27", 31, 2, 1), ("golden/syn150_k27", 59, 2, 1),
    ("golden/syn150_k27", 99, 2, 1), ("golden/syn150_k27", 141, 1, 0),
    ("golden/synvar_k21_m3", 21, 3, 1), ("golden/synvar_k21_m3", 31, 1, 0), ("golden/synvar_k21_m3", 25, 2, 1),
    ("golden/polya_k27", 27, 2, 1), ("golden/polya_k27", 27, 1, 0), ("golden/tandem_k27", 27, 2, 1),
    ("golden/tandem_k27", 28, 2, 1), ("golden/lowcov_k21", 21, 2, 1), ("golden/empty_k21", 21, 2, 1),
    ("golden_kmax/syn300_k255", 255, 1, 0), ("golden_kmax/syn300_k255", 255, 2, 1), ("golden_kmax/syn300_k255", 199, 2, 1),
    ("golden_kmax/syn300_k255", 237, 2, 1),  # the widest stage-1 record the device sort takes (17 words)
    # buckets far above kmsort's insertion-sort threshold (64): the American-flag permutation decides the tie order
    ("synth:deep", 27, 2, 1), ("synth:deep", 21, 3, 1), ("synth:wide", 27, 2, 1), ("synth:wide", 23, 1, 0),
    ("synth:mid", 27, 2, 1),
]

# seeded libraries (megahit_b200.synth.synth_reads arguments): regenerated identically by the tests
SYNTH = {
    "deep": dict(n_reads=20000, read_len=100, genome_len=30000, err=0.01, seed=101),   # ~65x: tie classes > 64
    "wide": dict(n_reads=60000, read_len=150, genome_len=1000000, err=0.01, seed=102),  # ~116 diverse records per bucket
    "mid": dict(n_reads=40000, read_len=120, genome_len=200000, err=0.02, seed=103),    # 24x, two radix levels in hot buckets
}


def write_synth(name, prefix):
    a = SYNTH[name]
    b = synth.synth_reads(a["n_reads"], a["read_len"], a["genome_len"], a["err"], seed=a["seed"])
    F.write_lib(prefix, b, a["n_reads"], a["n_reads"] * a["read_len"], a["read_len"])


def run_ref(lib, prefix, k, m, mercy, threads, mem_flag):
    cmd = [REF, "read2sdbg", "-k", str(k), "-m", str(m), "--host_mem", "4e9", "--mem_flag", str(mem_flag),
           "--output_prefix", prefix, "--num_cpu_threads", str(threads), "--read_lib_file", lib]
    if mercy:
        cmd.append("--need_mercy")
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if r.returncode != 0:
        sys.stderr.write(r.stderr.decode()[-3000:])
        raise SystemExit("reference read2sdbg failed: " + " ".join(cmd))
    return r.stderr.decode()


def digest(prefix, m):
    info, stream, table = F.canonical_sdbg(prefix)
    d = {"sdbg_k": info.k, "sdbg_words_per_tip_label": info.words_per_tip_label, "sdbg_items": int(table[:, 0].sum()),
         "sdbg_tips": int(table[:, 1].sum()), "sdbg_large_mul": int(table[:, 2].sum()),
         "sdbg_sha256": F.sha256(stream), "sdbg_bytes": len(stream)}
    if m > 1:  # stage 1 only runs for m > 1 (main_sdbg_build.cpp:141-147)
        d["counting_sha256"] = F.sha256(open(prefix + ".counting", "rb").read())
    return d


def main():
    os.makedirs(OUT, exist_ok=True)
    res = []
    with tempfile.TemporaryDirectory() as tmp:
        libs = {}
        for lib, k, m, mercy in RUNS:
            if lib not in libs:
                if lib.startswith("synth:"):
                    libs[lib] = os.path.join(tmp, lib[6:])
                    write_synth(lib[6:], libs[lib])
                else:
                    libs[lib] = os.path.join(ROOT, "tests", lib, "reads.lib")
            p = os.path.join(tmp, "o")
            log = run_ref(libs[lib], p, k, m, mercy, 4, 1)
            d = digest(p, m)
            n_mercy = [l.split()[-1] for l in log.splitlines() if "Number mercy" in l]
            for threads, mem_flag in ((1, 1), (3, 0)):
                run_ref(libs[lib], p + "x", k, m, mercy, threads, mem_flag)
                assert digest(p + "x", m) == d, f"{lib} k={k} m={m}: depends on threads / pass boundaries!"
            d.update({"lib": lib, "k": k, "m": m, "mercy": mercy, "n_mercy": int(n_mercy[0]) if n_mercy else 0})
            res.append(d)
            print(lib, k, m, mercy, d["sdbg_items"], d["sdbg_tips"], d["n_mercy"])
    json.dump({"synth": SYNTH, "runs": res}, open(os.path.join(OUT, "r2s.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    if not os.path.exists(REF):
        raise SystemExit("build oracle/_ref first: make -C oracle ref")
    main()
