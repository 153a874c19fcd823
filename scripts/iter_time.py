#!/usr/bin/env python
"""Times `iterate` (21 -> 29 and 21 -> 41) on a repeat-rich synthetic library: contigs assembled by the reference binary on
the box's host cores, then both `megahit_core iterate` binaries on the same files.  Prints one JSON line per step."""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from megahit_b200 import formats as F  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "megahit_core_ref")
OURS = os.path.join(ROOT, "megahit_b200", "bin", "megahit_core")
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
L, G = 150, 5 * n_reads
rng = np.random.default_rng(12)
g = rng.integers(0, 4, G, dtype=np.uint8)
for rl, per_mb in ((30, 260), (45, 200), (70, 130)):
    rep = rng.integers(0, 4, rl, dtype=np.uint8)
    for p in rng.choice(G - rl, max(1, per_mb * G // 1_000_000), replace=False):
        g[p:p + rl] = rep
with tempfile.TemporaryDirectory() as tmp:
    out = []
    chunk = 1 << 18
    for s in range(0, n_reads, chunk):
        n = min(chunk, n_reads - s)
        pos = rng.integers(0, G - L + 1, size=n)
        b = g[pos[:, None] + np.arange(L)[None, :]]
        rc = rng.integers(0, 2, size=n).astype(bool)
        b[rc] = 3 - b[rc][:, ::-1]
        e = rng.random(b.shape) < 0.01
        b[e] = (b[e] + rng.integers(1, 4, size=int(e.sum()), dtype=np.uint8)) & 3
        out.append(F.pack_reads_fixed(b))
    F.write_lib(f"{tmp}/reads.lib", np.concatenate(out), n_reads, n_reads * L, L)
    t = str(os.cpu_count())

    def run(cmd):
        t0 = time.time()
        r = subprocess.run(cmd, capture_output=True, text=True, cwd=tmp)
        assert r.returncode == 0, (cmd, r.stderr[-1500:])
        return time.time() - t0, r.stderr

    p21 = f"{tmp}/k21"
    t_build = run([OURS, "count", "-k", "21", "-m", "2", "--host_mem", "6e10", "--output_prefix", p21, "--num_cpu_threads", t,
                   "--read_lib_file", f"{tmp}/reads.lib"])[0]
    t_build += run([OURS, "seq2sdbg", "--host_mem", "6e10", "--output_prefix", p21, "--num_cpu_threads", t, "-k", "21", "--kmer_from",
                    "0", "--input_prefix", p21, "--need_mercy"])[0]
    t_asm = run([REF, "assemble", "-s", p21, "-o", p21, "-t", t, "--min_standalone", "200", "--prune_level", "2", "--merge_len", "20",
                 "--merge_similar", "0.95", "--cleaning_rounds", "5", "--disconnect_ratio", "0.1", "--low_local_ratio", "0.2",
                 "--min_depth", "2", "--bubble_level", "2", "--max_tip_len", "-1", "--careful_bubble"])[0]
    for step in (8, 20):
        line = {"what": "iterate", "n_reads": n_reads, "k": 21, "step": step, "sdbg_build_wall_s": round(t_build, 2),
                "reference_assemble_wall_s": round(t_asm, 2), "contig_bytes": os.path.getsize(p21 + ".contigs.fa")}
        for name, core in (("reference", REF), ("ours", OURS)):
            w, err = run([core, "iterate", "-c", p21 + ".contigs.fa", "-b", p21 + ".bubble_seq.fa", "-t", t, "-k", "21", "-s", str(step),
                          "-o", f"{tmp}/{name}{step}", "-r", f"{tmp}/reads.lib.bin"])
            info = open(f"{tmp}/{name}{step}.edges.info").read().split()
            line[name] = {"wall_s": round(w, 3), "n_edges": int(info[9])}
            gpu = [x for x in err.splitlines() if "iterate done" in x]
            if gpu:
                line[name]["log"] = gpu[-1].split("- ")[-1]
        W = int(info[3])
        a = np.unique(np.fromfile(f"{tmp}/reference{step}.edges.0", np.uint32).reshape(-1, W), axis=0)
        c = np.fromfile(f"{tmp}/ours{step}.edges.0", np.uint32).reshape(-1, W)
        line["identical_sets"] = bool(len(a) == len(c) and (a == c).all())
        line["reference_cores"] = os.cpu_count()
        print(json.dumps(line), flush=True)
