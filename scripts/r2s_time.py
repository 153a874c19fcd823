#!/usr/bin/env python
"""Times the 1-pass build (mhb_read2sdbg_host) on a synthetic library and, optionally, the reference binary's
`read2sdbg` on the same library (host cores).  Prints one JSON line per configuration."""
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
from megahit_b200 import lib, synth  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
with_ref = len(sys.argv) > 2 and sys.argv[2] == "ref"
L, k = 150, 27
b = synth.synth_reads(n_reads, L, 5 * n_reads, 0.01, seed=99)
for m, mercy in ((2, True), (1, False)):
    lib.read2sdbg_host(b.reshape(-1), n_reads, k, m, mercy)  # warm-up (allocations, module load)
    os.environ["MHB_R2S_TRACE"] = "1"  # per-phase times on stderr
    t0 = time.time()
    g = lib.read2sdbg_host(b.reshape(-1), n_reads, k, m, mercy)
    del os.environ["MHB_R2S_TRACE"]
    wall = time.time() - t0
    line = {"what": "read2sdbg", "n_reads": n_reads, "k": k, "m": m, "mercy": mercy, "edge_positions": g["n_edge_records"],
            "sort_items": g["n_sort_items"], "distinct_items": g["n_distinct_items"], "sdbg_items": g["n_items"],
            "n_mercy": g["n_mercy"], "ms": g["ms"], "wall_s": round(wall, 3),
            "edges_per_s": g["n_edge_records"] / (g["ms"]["total"] / 1e3)}
    if with_ref:
        ref = os.path.join(ROOT, "oracle", "_ref", "megahit_core_ref")
        with tempfile.TemporaryDirectory() as tmp:
            F.write_lib(f"{tmp}/r", b, n_reads, n_reads * L, L)
            t0 = time.time()
            subprocess.run([ref, "read2sdbg", "-k", str(k), "-m", str(m), "--host_mem", "6e10", "--mem_flag", "1",
                            "--output_prefix", f"{tmp}/o", "--num_cpu_threads", str(os.cpu_count()), "--read_lib_file", f"{tmp}/r"]
                           + (["--need_mercy"] if mercy else []), check=True, capture_output=True)
            line["reference_s"] = round(time.time() - t0, 3)
            line["reference_cores"] = os.cpu_count()
    print(json.dumps(line), flush=True)
