"""The reference-side binding: integration/scasa_r.c must follow include/scasa_cuda.h.
R is not installed, so the shim is type-checked against stand-in R declarations (tests/mock_r/)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_r_shim_compiles_against_the_header():
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only",
                        "-I" + os.path.join(ROOT, "tests", "mock_r"), "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "integration", "scasa_r.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_r_shim_calls_only_exported_entry_points():
    from scasa_b200 import lib
    L = lib.load()
    src = open(os.path.join(ROOT, "integration", "scasa_r.c")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)                       # calls, not mentions in comments
    called = set(re.findall(r"\b(scasa_[a-z0-9_]+)\s*\(", src)) - {"scasa_r_create", "scasa_r_quant", "scasa_r_genemat"}
    assert {"scasa_ctx_create", "scasa_quant", "scasa_eqclass_map", "scasa_chunk_split", "scasa_gene_sum"} <= called
    for name in called:
        assert hasattr(L, name), name


def test_integration_md_shows_the_shipped_files():
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for f in ("scasa_r.c", "scasa_quant_step2_patch.R"):
        body = open(os.path.join(ROOT, "integration", f)).read().strip()
        assert body in md, f
