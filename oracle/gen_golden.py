#!/usr/bin/env python
"""Mint golden fixtures under tests/golden/ by running the UNMODIFIED reference binary
(oracle/_ref/megahit_core_ref, built by oracle/Makefile from /root/reference) on small inputs.

Run in the build container only (needs /root/reference for the toy reads and the Python driver):

    python oracle/gen_golden.py            # all cases
    python oracle/gen_golden.py syn150_k27 # one case

For every case the input read library is committed next to a `golden.json` holding sha256 digests of
the canonical streams: bucket-ordered edges, `.cand`, `.counting`, bucket-ordered SdBG
items + header fields.  Small cases also keep the raw streams for debugging.
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
GOLD = os.path.join(ROOT, "tests", "golden")
REFROOT = "/root/reference"


def run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, **kw)
    if r.returncode != 0:
        sys.stderr.write(r.stderr.decode()[-4000:])
        raise SystemExit(f"command failed: {' '.join(cmd)}")
    return r.stderr.decode()


def ref_count(lib, prefix, k, m, threads=4):
    return run([REF, "count", "-k", str(k), "-m", str(m), "--host_mem", "4e9", "--mem_flag", "1",
                "--output_prefix", prefix, "--num_cpu_threads", str(threads), "--read_lib_file", lib])


def ref_seq2sdbg(prefix, k, out_prefix, mercy=True, threads=4, extra=()):
    cmd = [REF, "seq2sdbg", "--host_mem", "4e9", "--mem_flag", "1", "--output_prefix", out_prefix,
           "--num_cpu_threads", str(threads), "-k", str(k), "--kmer_from", "0"]
    if prefix:
        cmd += ["--input_prefix", prefix]
    if mercy:
        cmd.append("--need_mercy")
    return run(cmd + list(extra))


def digest_count_sdbg(prefix, keep_raw_dir=None):
    edges = F.canonical_edges(prefix)
    cand = open(prefix + ".cand", "rb").read()
    counting = open(prefix + ".counting", "rb").read()
    info, stream, table = F.canonical_sdbg(prefix)
    d = {
        "n_solid": int(len(edges)), "words_per_edge": int(edges.shape[1]),
        "edges_sha256": F.sha256(edges.tobytes()), "cand_sha256": F.sha256(cand), "cand_bytes": len(cand),
        "counting_sha256": F.sha256(counting),
        "sdbg_k": info.k, "sdbg_words_per_tip_label": info.words_per_tip_label,
        "sdbg_items": int(table[:, 0].sum()), "sdbg_tips": int(table[:, 1].sum()),
        "sdbg_large_mul": int(table[:, 2].sum()), "sdbg_sha256": F.sha256(stream), "sdbg_bytes": len(stream),
    }
    if keep_raw_dir:
        edges.tofile(os.path.join(keep_raw_dir, "edges.canon.bin"))
        open(os.path.join(keep_raw_dir, "cand.bin"), "wb").write(cand)
        open(os.path.join(keep_raw_dir, "sdbg.canon.bin"), "wb").write(stream)
    return d


def case_dir(name):
    d = os.path.join(GOLD, name)
    os.makedirs(d, exist_ok=True)
    return d


def make_case(name, lib_prefix_src, ks, m, keep_raw=False, threads=4):
    """lib_prefix_src: path prefix of an existing .bin/.lib_info pair; copied into the case dir."""
    d = case_dir(name)
    for suf in (".bin", ".lib_info"):
        shutil.copy(lib_prefix_src + suf, os.path.join(d, "reads.lib" + suf))
    lib = os.path.join(d, "reads.lib")
    res = {"m": m, "by_k": {}}
    with tempfile.TemporaryDirectory() as tmp:
        for k in ks:
            p = os.path.join(tmp, f"k{k}")
            ref_count(lib, p, k, m, threads)
            ref_seq2sdbg(p, k, p, True, threads)
            raw = d if (keep_raw and k == ks[0]) else None
            res["by_k"][str(k)] = digest_count_sdbg(p, raw)
            # determinism across thread counts (the canonical-stream invariant)
            p1 = os.path.join(tmp, f"k{k}_t1")
            ref_count(lib, p1, k, m, 1)
            ref_seq2sdbg(p1, k, p1, True, 1)
            assert digest_count_sdbg(p1) == res["by_k"][str(k)], f"{name} k={k}: thread-count dependent!"
    json.dump(res, open(os.path.join(d, "golden.json"), "w"), indent=1, sort_keys=True)
    print(name, {k: (v["n_solid"], v["sdbg_items"], v["sdbg_tips"]) for k, v in res["by_k"].items()})


def gen_toy():
    """BASELINE.json configs[0]: the reference's own test_data toy PE+SE set, k=21."""
    with tempfile.TemporaryDirectory() as tmp:
        R = os.path.join(REFROOT, "test_data")
        subprocess.check_call(f"gzip -dc {R}/r1.il.fa.gz > {tmp}/r1.fa && bzip2 -dc {R}/r2.il.fa.bz2 > {tmp}/r2.fa",
                              shell=True)
        with open(f"{tmp}/reads.lib", "w") as f:
            f.write(f"il1\ninterleaved {tmp}/r1.fa\nil2\ninterleaved {tmp}/r2.fa\n"
                    f"pe\npe {R}/r3_1.fa {R}/r3_2.fa\nse4\nse {R}/r4.fa\nloop\nse {R}/loop.fa\n")
        run([REF, "buildlib", f"{tmp}/reads.lib", f"{tmp}/reads.lib"])
        make_case("toy_k21", f"{tmp}/reads.lib", [21], 2, keep_raw=True)


def gen_syn150():
    with tempfile.TemporaryDirectory() as tmp:
        b = synth.synth_reads(3000, 150, 15000, 0.01, seed=7)
        F.write_lib(f"{tmp}/r", b, 3000, 3000 * 150, 150)
        make_case("syn150_k27", f"{tmp}/r", [27], 2, keep_raw=True)
        make_case("syn150_klist", f"{tmp}/r", [21, 29, 39, 59, 79, 99, 119, 141], 2)


def gen_synvar():
    with tempfile.TemporaryDirectory() as tmp:
        w = synth.synth_reads_varlen(1500, 0, 260, 12000, 0.01, seed=11)
        n = 1500
        # count bases as the reference does (zero-length reads become 1 fake base)
        pos, nb, mx = 0, 0, 0
        while pos < len(w):
            L = int(w[pos])
            nb += max(L, 1)
            mx = max(mx, L)
            pos += 1 + (L + 15) // 16
        F.write_lib(f"{tmp}/r", w, n, nb, mx)
        make_case("synvar_k21_m3", f"{tmp}/r", [21], 3)
        make_case("synvar_k31_m1", f"{tmp}/r", [31], 1)


def gen_degenerate():
    """Edge cases the reference's simple_test exercises: empty input and all-identical reads."""
    with tempfile.TemporaryDirectory() as tmp:
        F.write_lib(f"{tmp}/e", np.zeros(0, np.uint32), 0, 0, 0)
        make_case("empty_k21", f"{tmp}/e", [21], 2)
        polya = np.zeros((400, 120), np.uint8)
        polya[200:, :] = 3  # poly-T == rc of poly-A: one giant run, palindromic neighbourhoods
        F.write_lib(f"{tmp}/p", F.pack_reads_fixed(polya), 400, 400 * 120, 120)
        make_case("polya_k27", f"{tmp}/p", [27], 2)
        rng = np.random.default_rng(5)
        unit = rng.integers(0, 4, 7, dtype=np.uint8)
        rep = np.tile(unit, 40)[:150]
        reads = np.stack([np.roll(rep, int(s))[:150] for s in rng.integers(0, 7, 600)])
        F.write_lib(f"{tmp}/t", F.pack_reads_fixed(reads), 600, 600 * 150, 150)
        make_case("tandem_k27", f"{tmp}/t", [27, 28], 2)


def gen_chain():
    """k > k_min: seq2sdbg fed by the reference's own assemble/local/iterate outputs (contigs with multiplicity and
    flags, bubble sequences, additional + local contigs, UNSORTED iterative edges).  Captured by running the
    reference's Python driver (src/megahit, used from /root/reference at generation time only) with --keep-tmp-files."""
    R = os.path.join(REFROOT, "test_data")
    jobs = {
        "chain_syn150": None,  # reads written below
        "chain_toy": ["-1", f"{R}/r3_1.fa", "-2", f"{R}/r3_2.fa", "--12", f"{R}/r1.il.fa.gz,{R}/r2.il.fa.bz2", "-r",
                      f"{R}/r4.fa,{R}/loop.fa"],
    }
    for name, reads_args in jobs.items():
        with tempfile.TemporaryDirectory() as tmp:
            os.makedirs(f"{tmp}/bin")
            shutil.copy(os.path.join(REFROOT, "src", "megahit"), f"{tmp}/bin/megahit")
            os.symlink(REF, f"{tmp}/bin/megahit_core")
            if reads_args is None:
                b = synth.synth_reads(3000, 150, 15000, 0.01, seed=7)
                with open(f"{tmp}/reads.fa", "w") as f:
                    for i, row in enumerate(b):
                        L, w = int(row[0]), row[1:]
                        f.write(f">r{i}\n" + "".join("ACGT"[(int(w[j >> 4]) >> (30 - 2 * (j & 15))) & 3] for j in range(L)) + "\n")
                reads_args = ["-r", f"{tmp}/reads.fa"]
            run([sys.executable, f"{tmp}/bin/megahit", *reads_args, "--k-list", "21,29,39", "-o", f"{tmp}/out",
                 "--keep-tmp-files", "-t", "4"])
            d = case_dir(name)
            ic = f"{tmp}/out/intermediate_contigs"
            for fn in ("k21.contigs.fa", "k21.bubble_seq.fa", "k21.addi.fa", "k21.local.fa"):
                shutil.copy(f"{ic}/{fn}", d)
                shutil.copy(f"{ic}/{fn}.info", d)
            for fn in ("29.edges.0", "29.edges.info"):
                shutil.copy(f"{tmp}/out/tmp/k29/{fn}", d)
            info, stream, table = F.canonical_sdbg(f"{tmp}/out/tmp/k29/29")
            res = {"k": 29, "k_from": 21, "sdbg_k": info.k, "sdbg_words_per_tip_label": info.words_per_tip_label,
                   "sdbg_items": int(table[:, 0].sum()), "sdbg_tips": int(table[:, 1].sum()),
                   "sdbg_large_mul": int(table[:, 2].sum()), "sdbg_sha256": F.sha256(stream), "sdbg_bytes": len(stream)}
            # thread-count independence of the reference on this input
            p1 = f"{tmp}/t1"
            run([REF, "seq2sdbg", "--host_mem", "4e9", "--mem_flag", "1", "--output_prefix", p1, "--num_cpu_threads", "1",
                 "-k", "29", "--kmer_from", "21", "--input_prefix", f"{tmp}/out/tmp/k29/29", "--addi_contig", f"{ic}/k21.addi.fa",
                 "--local_contig", f"{ic}/k21.local.fa", "--contig", f"{ic}/k21.contigs.fa", "--bubble", f"{ic}/k21.bubble_seq.fa"])
            assert F.sha256(F.canonical_sdbg(p1)[1]) == res["sdbg_sha256"]
            json.dump(res, open(os.path.join(d, "chain.json"), "w"), indent=1, sort_keys=True)
            print(name, res["sdbg_items"], res["sdbg_tips"], "loop contigs:",
                  sum(1 for l in open(f"{d}/k21.contigs.fa") if l.startswith(">") and "flag=2" in l or "flag=3" in l))


def gen_kmax():
    """The largest k the reference supports (kmax = 255, main.cpp:100-103): 17-word count records, 16-word tip labels.
    Kept apart from tests/golden/ (tests/golden_kmax/): the oracle is pinned against it on the CPU, the GPU comparison
    at this width has not been run yet."""
    global GOLD
    keep = GOLD
    GOLD = os.path.join(ROOT, "tests", "golden_kmax")
    try:
        with tempfile.TemporaryDirectory() as tmp:
            b = synth.synth_reads(700, 300, 9000, 0.004, seed=21)
            F.write_lib(f"{tmp}/r", b, 700, 700 * 300, 300)
            make_case("syn300_k255", f"{tmp}/r", [255, 199], 2)
    finally:
        GOLD = keep


def gen_lowcov():
    """Reads that overlap only at their ends (coverage 1, 2 inside the 30-base overlaps): with m = 2 only the overlap
    (k+1)-mers are solid, every read has a "no out" tip at its left overlap and a "no in" tip at its right one, so almost
    every (k+1)-mer of every read comes back as a mercy edge - far more mercy edges than solid ones, and more than
    n/m + 1 - n_solid (the capacity the fused build used to give them: ADVICE r1)."""
    with tempfile.TemporaryDirectory() as tmp:
        rng = np.random.default_rng(21)
        L, ov, n = 150, 30, 400
        g = rng.integers(0, 4, (L - ov) * n + ov, dtype=np.uint8)
        reads = np.stack([g[i * (L - ov): i * (L - ov) + L] for i in range(n)])
        flip = rng.random(n) < 0.5  # random strand
        reads[flip] = (3 - reads[flip])[:, ::-1]
        F.write_lib(f"{tmp}/l", F.pack_reads_fixed(reads), n, n * L, L)
        make_case("lowcov_k21", f"{tmp}/l", [21], 2)


CASES = {"lowcov": gen_lowcov, "kmax": gen_kmax, "chain": gen_chain, "toy": gen_toy, "syn150": gen_syn150, "synvar": gen_synvar, "degenerate": gen_degenerate}

if __name__ == "__main__":
    if not os.path.exists(REF):
        raise SystemExit("build oracle/_ref first: make -C oracle ref")
    sel = sys.argv[1:] or list(CASES)
    for c in sel:
        CASES[c]()
