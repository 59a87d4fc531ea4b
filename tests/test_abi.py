"""The C-ABI library loads on a machine without a GPU, exports every symbol include/ovo_b200.h declares, the
ctypes table covers exactly those symbols, and calls fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ovo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ovo_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib_built):
    lib = ctypes.CDLL(lib_built)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ovo_b200.h but not exported"


def test_ctypes_table_matches_header(lib_built):
    from ovo_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()
    assert _lib.lib().ovo_version() >= 100


def test_struct_sizes():
    from ovo_b200 import _lib
    assert ctypes.sizeof(_lib.VoteRow) == 32
    assert ctypes.sizeof(_lib.BlockWeights) == 12 * 8
    assert ctypes.sizeof(_lib.VitCfg) == 15 * 4


def test_fails_loudly_without_gpu(lib_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ovo_b200 import _lib
    h = ctypes.c_void_p()
    rc = _lib.lib().ovo_map_create(ctypes.byref(h))
    assert rc < 0 and len(_lib.lib().ovo_last_error()) > 0
    with pytest.raises(RuntimeError):
        from ovo_b200.map import SemanticMap
        SemanticMap()
    with pytest.raises(RuntimeError):
        from ovo_b200.encoder import RegionEncoder, EncoderConfig
        RegionEncoder(EncoderConfig(), {})


def test_no_oracle_import_in_product():
    """The product package never imports the oracle (the oracle is test infrastructure)."""
    pkg = os.path.join(ROOT, "ovo_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "/root/reference" not in txt or f == "tokenizer.py", f
