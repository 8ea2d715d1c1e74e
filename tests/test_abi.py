# SPDX-License-Identifier: Apache-2.0
"""The C-ABI library loads without a GPU and exports exactly what include/xntt.h declares."""
import os
import re

import pytest

from conftest import ROOT


def declared_functions():
    text = open(os.path.join(ROOT, "include", "xntt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xntt_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(pkg):
    assert declared_functions() == sorted(pkg.SYMBOLS)


def test_library_exports_every_declared_symbol(pkg):
    assert os.path.exists(pkg.LIB_PATH), "build with __graft_entry__.build()"
    lib = pkg.Library(pkg.LIB_PATH)  # resolves every symbol, raises AttributeError otherwise
    for name in declared_functions():
        assert hasattr(lib.lib, name)
    assert lib.version().endswith("sm_100a")
    assert lib.lib.xntt_strerror(0) == b"ok"
    assert lib.lib.xntt_strerror(-1) == b"invalid argument"


def test_library_is_sm100a_only(pkg):
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", pkg.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_no_fallback_when_library_missing(pkg, tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.Library(str(tmp_path / "libxntt.so"))


def test_product_never_touches_the_oracle():
    """Nothing under sve-ntt_b200/ or include/ may reference oracle/ or the host emulator."""
    bad = []
    for base in ("sve-ntt_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            if any(part in dp for part in ("build", "lib", "_build", "__pycache__")):
                continue
            for fn in fns:
                if fn.endswith((".so", ".o", ".pyc")):
                    continue
                text = open(os.path.join(dp, fn), errors="ignore").read()
                if re.search(r"oracle/|libntt_oracle|libnttref|libxntt_emu|oracle_lib", text):
                    bad.append(os.path.join(dp, fn))
    assert not bad, bad
