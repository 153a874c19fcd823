"""The oracle is test infrastructure: nothing under megahit_b200/ may import, link or call it."""
import os
import re

from conftest import ROOT


def test_product_never_touches_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "megahit_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"import\s+oracle|from\s+oracle|oracle[./]|mhbo_|libmhb_oracle|mhb_oracle", src):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, f"product files reference the oracle: {bad}"


def test_required_layout():
    for p in ("include/mhb.h", "oracle/mhb_oracle.c", "oracle/Makefile", "bench.py", "__graft_entry__.py",
              "DESIGN.md", "INTEGRATION.md", "tests/golden"):
        assert os.path.exists(os.path.join(ROOT, p)), p
