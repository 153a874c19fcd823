"""GPU tests for code paths written after this round's GPU budget was spent: they are opt-in paths (environment
variables) that do not change the default behaviour, marked xfail(strict=False) until a GPU run has confirmed them -
a pass shows up as XPASS, a failure cannot break the suite.  Remove the marker once verified."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="written without GPU access; verify, then drop the marker")]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CHILD = r"""
import json, os, sys
import numpy as np
sys.path.insert(0, %r)
from megahit_b200 import formats as F, lib
case, k, m = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
bin_words = np.fromfile(os.path.join(case, "reads.lib.bin"), np.uint32)
_, n_reads = F.read_lib_info(os.path.join(case, "reads.lib"))
g = lib.build_host(bin_words, n_reads, k, m, need_mercy=True, want_edges=True)
print("RESULT " + json.dumps({"sdbg": F.sha256(lib.sdbg_stream_from_table(g["bucket_table"], g["bytes"])),
                              "edges": F.sha256(g["edges"].tobytes()), "n_items": int(g["n_items"])}))
""" % ROOT


@pytest.mark.parametrize("name,k,m,gold", [c for c in golden_cases() if c.values[0] in ("syn150_k27", "tandem_k27", "polya_k27")])
def test_fused_build_with_chunked_upload_matches_reference(name, k, m, gold):
    """MHB_H2D_CHUNKS: the library uploaded in pieces, extraction overlapping the copies -> same SdBG as the reference"""
    import json
    env = dict(os.environ, MHB_H2D_CHUNKS="3")
    p = subprocess.run([sys.executable, "-c", _CHILD, os.path.join(GOLDEN, name), str(k), str(m)], env=env,
                       capture_output=True, text=True, timeout=300)
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
    assert line, p.stderr[-800:]
    r = json.loads(line[-1][7:])
    assert r["sdbg"] == gold["sdbg_sha256"] and r["edges"] == gold["edges_sha256"]


def test_new_radix_pass_variants_sort_correctly():
    """compact look-back descriptors (bit 15) and two-stream ranking (bit 16): every new variant, each in its own
    process (scripts/sort_sweep.py: a hang or crash only loses that variant), must reproduce torch's stable sort"""
    import json
    cfgs = [256 + b for b in (0x8080, 0x10080, 0x18080, 0x9080, 0x8082, 0x10082)]
    p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "sort_sweep.py"), ",".join(map(str, cfgs)), "3000000",
                        "1000000"], capture_output=True, text=True, timeout=900)
    res = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(res) == len(cfgs), p.stderr[-800:]
    bad = [r for r in res if not r.get("ok")]
    assert not bad, bad


_CHILD_ROLL = r"""
import json, os, sys
import numpy as np
sys.path.insert(0, %r)
from megahit_b200 import formats as F, lib
case, k, m = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
bin_words = np.fromfile(os.path.join(case, "reads.lib.bin"), np.uint32)
_, n_reads = F.read_lib_info(os.path.join(case, "reads.lib"))
g = lib.count_host(bin_words, n_reads, k, m, want_mercy=True)
print("RESULT " + json.dumps({"edges": F.sha256(g["edges"].tobytes()), "n_solid": int(g["n_solid"]),
                              "cand": [int(x) for x in g["cand_ids"][:50]], "n_cand": int(len(g["cand_ids"])),
                              "n_has_tips": int(g["n_has_tips"])}))
""" % ROOT


@pytest.mark.parametrize("name,k,m,gold", [c for c in golden_cases() if c.values[0] in ("syn150_k27", "toy_k21", "tandem_k27",
                                                                                      "polya_k27", "synvar_k21_m3")])
def test_rolling_extract_and_mark_match_default(name, k, m, gold):
    """MHB_EXTRACT_ROLL=1 (rolling record builder in the extract and mercy-mark kernels): same edges, same candidate
    reads as the default kernels and as the reference"""
    import json
    out = []
    for roll in (False, True):
        env = dict(os.environ)
        env.pop("MHB_EXTRACT_ROLL", None)
        if roll:
            env["MHB_EXTRACT_ROLL"] = "1"
        p = subprocess.run([sys.executable, "-c", _CHILD_ROLL, os.path.join(GOLDEN, name), str(k), str(m)], env=env,
                           capture_output=True, text=True, timeout=300)
        line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
        assert line, p.stderr[-800:]
        out.append(json.loads(line[-1][7:]))
    assert out[0] == out[1]
    assert out[1]["edges"] == gold["edges_sha256"] and out[1]["n_solid"] == gold["n_solid"]
