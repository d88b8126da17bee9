"""CPU tests of the drop-in boundary: the C-ABI library builds/loads here (no GPU), exports every symbol that
include/dm_abi.h declares, the ctypes table covers them all, and the product path refuses to run without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "dm_abi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from diff_mining_b200 import _abi

    if not os.path.exists(_abi.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    lib = ctypes.CDLL(_abi.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in dm_abi.h but not exported"


def test_ctypes_table_matches_header():
    from diff_mining_b200 import _abi

    assert sorted(_abi.SIGNATURES) == _declared_symbols()
    lib = _abi.load()
    assert lib.dm_abi_version() == 1


def test_variant_switches_documented_in_the_header_are_accepted():
    """every kernel-variant name dm_abi.h documents is known to the library (a host-side switch: no GPU needed) and an
    unknown one is an error with a message, not a silent no-op"""
    from diff_mining_b200 import _abi

    hdr = open(os.path.join(ROOT, "include", "dm_abi.h")).read()
    doc = hdr[hdr.index("kernel-variant switches"):hdr.index("int dm_op_set_variant")]
    names = re.findall(r"^ \*   ([a-z0-9_]+) ", doc, flags=re.M)
    assert {"igemm_pair", "igemm_ng4", "igemm_ws", "gn_fused", "gn_epilogue", "xattn", "attn3", "prefix_share"} <= set(names)
    lib = _abi.load()
    for n in names:
        assert lib.dm_op_set_variant(n.encode(), -1) == 0, n
    assert lib.dm_op_set_variant(b"no_such_variant", 1) != 0
    assert b"no_such_variant" in lib.dm_last_error()


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from diff_mining_b200.engine import Engine

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(0)


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "diff-mining_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "from oracle" not in txt and "import oracle" not in txt, f
