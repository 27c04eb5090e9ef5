"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports
every symbol include/frecsys_b200.h declares; without a GPU compute calls fail loudly."""
import os
import re

import pytest

import helpers


def header_symbols():
    src = open(os.path.join(helpers.ROOT, "include", "frecsys_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(frx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    pkg = helpers.load_pkg()
    L = pkg.lib()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/frecsys_b200.h but not exported"
    assert sorted(pkg.EXPORTS) == syms


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    pkg = helpers.load_pkg()
    with pytest.raises(pkg.FrxError, match="no CPU fallback"):
        pkg.Context(0)


def test_config_struct_layout_matches_header():
    pkg = helpers.load_pkg()
    src = open(os.path.join(helpers.ROOT, "include", "frecsys_b200.h")).read()
    body = src[src.index("typedef struct frx_config {"):src.index("} frx_config;")]
    fields = re.findall(r"^\s*(?:int|float|unsigned)\s+([a-z_]+);", body, flags=re.M)
    assert fields == [f[0] for f in pkg.FrxConfig._fields_]


def test_product_does_not_touch_the_oracle():
    """The product path must never import, link or call anything under oracle/."""
    bad = []
    for base in ("safer2-recommender_b200", "include", "tools"):
        for dp, _, fns in os.walk(os.path.join(helpers.ROOT, base)):
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp", "Makefile")):
                    txt = open(os.path.join(dp, fn), errors="ignore").read()
                    if re.search(r"oracle[/\.]|frecsys_oracle|orc_", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad
