"""The C-ABI library loads and exports every symbol include/lastz_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from lastz_b200 import capi


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lastz_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lzb_[a-z_]+)\s*\(", text)))


def test_binding_lists_every_declared_symbol():
    assert sorted(capi.SYMBOLS) == declared_symbols()


@pytest.mark.parametrize("path", [capi.PRODUCT_LIB, capi.ORACLE_LIB])
def test_library_exports_the_abi(path):
    if not os.path.exists(path):
        pytest.skip(f"{path} not built")
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_struct_layouts_match_reference():
    # segment.h:46-64, edit_script.h:30-61 (offsets recorded in SURVEY.md 8b)
    S, A = capi.Segment, capi.Alignel
    assert ctypes.sizeof(S) == 48 and S.pos1.offset == 8 and S.s.offset == 20 and S.scoreCov.offset == 32 and S.filter.offset == 40
    assert ctypes.sizeof(A) == 64 and A.beg1.offset == 12 and A.s.offset == 28 and A.script.offset == 48 and A.hspId.offset == 56
    assert capi.EditScript.op.offset == 12


def test_product_has_no_cpu_fallback():
    """Without a GPU lzb_open must fail loudly (there is no CPU path in the product)."""
    if not os.path.exists(capi.PRODUCT_LIB):
        pytest.skip("product library not built")
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); from lastz_b200 import capi; lib = capi.load_product(); "
            "ctx = lib.lzb_open(0); print('CTX', bool(ctx), lib.lzb_last_error().decode())" % ROOT)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env).stdout
    assert "CTX False" in out and "no CPU fallback" in out
