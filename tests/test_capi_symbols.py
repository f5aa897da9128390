"""CPU: the C-ABI library loads without a GPU and exports every function include/ged_b200.h declares; compute entry
points refuse to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from ppo_bihyb_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "ged_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"typedef struct \w+ \{.*?\} \w+;", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ged_\w+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared_functions() == sorted(_lib.EXPORTED_SYMBOLS)


def test_library_loads_and_exports_every_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        from ppo_bihyb_b200 import build
        build.build()
    lib = _lib.load()
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(raw, name), name
    assert lib.ged_abi_version() == _lib.ABI_VERSION
    assert b"sm_100a" in lib.ged_version()
    n = ctypes.c_int32(0)
    m = ctypes.c_int32(0)
    lib.ged_limits(ctypes.byref(n), ctypes.byref(m))
    assert n.value >= 200 and m.value >= 101 * 101
    assert 0 < lib.ged_solve_smem_bytes(60, 60) <= 227 * 1024


def test_struct_layout_matches_header():
    """ctypes mirrors of ged_pairs / ged_solve_desc / ged_solve_out must have the C compiler's size (checked by
    compiling a two-line probe against the real header)."""
    import subprocess
    import tempfile
    probe = '#include "ged_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n", sizeof(ged_pairs), sizeof(ged_solve_desc), sizeof(ged_solve_out));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(probe)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "p"), os.path.join(d, "p.c")])
        sizes = [int(v) for v in subprocess.check_output([os.path.join(d, "p")]).split()]
    assert sizes == [ctypes.sizeof(_lib.GedPairs), ctypes.sizeof(_lib.GedSolveDesc), ctypes.sizeof(_lib.GedSolveOut)]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from ppo_bihyb_b200 import ged_env
    with pytest.raises(_lib.GedLibraryError):
        _lib.require_cuda()
    with pytest.raises(_lib.GedLibraryError):
        ged_env.hungarian_lap(torch.zeros(3, 3), 2, 2)


def test_bad_arguments_are_reported_not_crashed():
    lib = _lib.load()
    assert lib.ged_solve_batch(None, None, None) == _lib.GED_E_BADARG
    assert b"null" in lib.ged_last_error()
    assert lib.ged_lsape_batch(-1, None, None, 1, 1, None, 4, None, 1, None, 1, None, None, None) == _lib.GED_E_BADARG
    pairs = _lib.GedPairs()
    pairs.num_pairs, pairs.max_n1, pairs.max_n2 = 1, 300, 300
    assert lib.ged_kv_batch(ctypes.byref(pairs), ctypes.cast(8, _lib.c_f32p), ctypes.cast(8, _lib.c_f32p), ctypes.cast(8, _lib.c_f32p), 1 << 20, None) != _lib.GED_OK
