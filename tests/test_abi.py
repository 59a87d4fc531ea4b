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
    assert ctypes.sizeof(_lib.PoolHeadWeights) == 2 * 4 + 12 * 8         # ovo_pool_head_weights
    assert ctypes.sizeof(_lib.CropParams) == 6 * 4                        # ovo_crop_params
    assert ctypes.sizeof(_lib.MergerLayer) == 12 * 8                      # ovo_merger_layer
    assert ctypes.sizeof(_lib.MergerWeights) == 4 * 4 + 8 + 8 + 3 * 8 + 8  # ovo_merger_weights (int padded to the pointer)


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


def test_new_rows_fail_loudly_without_gpu(lib_built):
    """Crop-based descriptors, label transfer and the learned merger have no CPU path either."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ovo_b200 import _lib, eval_utils
    with pytest.raises(RuntimeError):
        eval_utils.knn(torch.rand(100, 3), torch.rand(10, 3), k=5)
    with pytest.raises(RuntimeError):
        eval_utils.match_labels_to_vtx(torch.zeros(100, dtype=torch.long), torch.rand(100, 3), torch.rand(10, 3))
    lib = _lib.lib()
    buf = (ctypes.c_float * 64)()
    out = (ctypes.c_int32 * 64)()
    rc = lib.ovo_knn(buf, 8, buf, 2, 1, 0.0, out, None, None)            # host pointers, no device: must not return success
    assert rc < 0 and len(lib.ovo_last_error()) > 0
    assert lib.ovo_knn(buf, 2, buf, 2, 5, 0.0, out, None, None) < 0      # fewer points than neighbours: argument check first
    assert lib.ovo_fuse_clips(buf, buf, buf, 1, 8, 0, 0.4, 0.1, buf, None) < 0     # `vanilla` has no fusion rule
    assert lib.ovo_encode_crops(None, None, 0, 0, None, 0, None, None, None, None) < 0


def test_no_oracle_import_in_product():
    """The product package never imports the oracle (the oracle is test infrastructure)."""
    pkg = os.path.join(ROOT, "ovo_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "/root/reference" not in txt or f == "tokenizer.py", f


def test_only_tests_smoke_and_bench_import_the_oracle():
    """oracle/ is test infrastructure: outside tests/ and oracle/ itself only bench.py (CPU legs) and __graft_entry__.py (smoke)
    may import it."""
    allowed = {os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")}
    for root, dirs, files in os.walk(ROOT):
        dirs[:] = [d for d in dirs if d not in (".git", "tests", "oracle", "gpurun_out", "baseline", "__pycache__", ".pytest_cache")]
        for f in files:
            path = os.path.join(root, f)
            if f.endswith(".py") and path not in allowed:
                assert not re.search(r"^\s*(from|import)\s+oracle\b", open(path).read(), flags=re.M), path


def test_header_is_plain_c99(tmp_path):
    """include/ovo_b200.h is the C ABI: it must compile as C99 (no C++ / torch types), and a C program must link against the library."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "abi.c"
    src.write_text('#include "ovo_b200.h"\n#include <stdio.h>\n'
                   'int main(void) { ovo_crop_params p; p.embed_type = OVO_EMBED_HOVSG; (void)p;\n'
                   '  ovo_map_t* m = 0; int rc = ovo_map_create(&m);\n'
                   '  printf("%d %d %s\\n", ovo_version(), rc, rc < 0 ? ovo_last_error() : "ok"); if (m) ovo_map_destroy(m); return 0; }\n')
    inc = os.path.join(ROOT, "include")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", "-I", inc, str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    from ovo_b200 import build
    lib = build.build()
    exe = tmp_path / "abi"
    r = subprocess.run([gcc, "-std=c99", "-I", inc, str(src), "-o", str(exe), lib, f"-Wl,-rpath,{os.path.dirname(lib)}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.split()[0] == "100", r.stdout + r.stderr
