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
    called = {n for n in re.findall(r"\b(scasa_[a-z0-9_]+)\s*\(", src) if not n.startswith("scasa_r_")}
    assert {"scasa_ctx_create", "scasa_quant_multi", "scasa_quant_files", "scasa_eqclass_map", "scasa_chunk_split", "scasa_gene_sum",
            "scasa_dm0_is_fixed_point", "scasa_set_gene_map", "scasa_device_count"} <= called
    for name in called:
        assert hasattr(L, name), name


def test_integration_md_shows_the_shipped_files():
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for f in ("scasa_r.c", "scasa_quant_step2_patch.R", "scasa_quant_step3_patch.R"):
        body = open(os.path.join(ROOT, "integration", f)).read().strip()
        assert body in md, f


def test_r_patches_call_only_entry_points_the_shim_defines():
    shim = open(os.path.join(ROOT, "integration", "scasa_r.c")).read()
    defined = set(re.findall(r"^SEXP (scasa_r_[a-z0-9_]+)\(", shim, flags=re.M))
    for f in ("scasa_quant_step2_patch.R", "scasa_quant_step3_patch.R"):
        src = open(os.path.join(ROOT, "integration", f)).read()
        for name in re.findall(r'\.Call\("([a-z0-9_]+)"', src):
            assert name in defined, (f, name)
    # .Call arity = parameters of the C function
    for name, args in re.findall(r"^SEXP (scasa_r_[a-z0-9_]+)\(([^)]*)\)", shim, flags=re.M | re.S):
        n_params = len([a for a in args.split(",") if a.strip()])
        for f in ("scasa_quant_step2_patch.R",):
            src = open(os.path.join(ROOT, "integration", f)).read()
            for m in re.finditer(r'\.Call\("%s"' % name, src):
                depth, i, n_args, cur = 0, m.end(), 0, ""
                while True:                                         # count top-level commas up to the closing parenthesis
                    ch = src[i]
                    if ch in "([":
                        depth += 1
                    elif ch in ")]":
                        if depth == 0:
                            break
                        depth -= 1
                    elif ch == "," and depth == 0:
                        n_args += 1
                    i += 1
                assert n_args == n_params, (name, n_args, n_params)
