"""The C-ABI library loads and exports every symbol include/dsrl_b200.h declares; without a GPU every compute
entry point fails loudly (no CPU fallback).  CPU only."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dsrl_b200.h")


@pytest.fixture(scope="module")
def lib():
    from dualsuperreslearningforsemseg_b200 import build, _lib
    build.build()
    return _lib.lib()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dsrl_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    from dualsuperreslearningforsemseg_b200 import _lib
    syms = declared_symbols()
    assert set(syms) == set(_lib.EXPORTS)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in dsrl_b200.h but not exported"


def test_version_and_size_queries(lib):
    assert lib.dsrl_version() == 200
    # size queries are pure host arithmetic: valid geometry > 0, invalid geometry 0
    assert lib.dsrl_fa_saved_bytes(0, 0, 6, 1, 1, 64, 128, 8) > 0
    assert lib.dsrl_fa_workspace_bytes(0, 0, 6, 1, 1, 64, 128, 8) > 0
    assert lib.dsrl_fa_saved_bytes(0, 0, 6, 1, 2, 64, 128, 8) == 0      # reference mode needs equal shapes
    assert lib.dsrl_fa_saved_bytes(0, 0, 1, 1, 1, 4, 4, 8) == 0         # smaller than the pooling window


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_entry_points_fail_loudly_without_gpu(lib):
    from dualsuperreslearningforsemseg_b200 import _lib
    buf = (ctypes.c_char * 4096)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    rc = lib.dsrl_seg_counts(p, _lib.I64, p, _lib.U8, None, 1, 16, 19, 255, p, None)
    assert rc == _lib.ERR_CUDA
    assert b"no CPU fallback" in lib.dsrl_last_error()
    rc = lib.dsrl_fa_forward(0, 0, p, p, 1, 1, 1, 8, 8, 8, 1, 1, p, p, 4096, p, 4096, None)
    assert rc == _lib.ERR_CUDA
    with pytest.raises(_lib.DsrlError):
        _lib.check(rc)


def test_argument_validation_happens_before_device_probe(lib):
    from dualsuperreslearningforsemseg_b200 import _lib
    buf = (ctypes.c_char * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.dsrl_seg_counts(p, _lib.I64, p, _lib.U8, None, 1, 16, 0, 255, p, None) == _lib.ERR_UNSUPPORTED
    assert lib.dsrl_seg_counts(None, _lib.I64, p, _lib.U8, None, 1, 16, 19, 255, p, None) == _lib.ERR_BAD_ARG
    assert lib.dsrl_fa_forward(0, 0, p, p, 1, 1, 2, 8, 8, 8, 1, 1, p, p, 64, p, 64, None) == _lib.ERR_BAD_SHAPE
    assert lib.dsrl_fa_forward(7, 0, p, p, 1, 1, 1, 8, 8, 8, 1, 1, p, p, 64, p, 64, None) == _lib.ERR_BAD_ARG


def test_cross_entropy_entry_points_validate_and_fail_loudly(lib):
    from dualsuperreslearningforsemseg_b200 import _lib
    buf = (ctypes.c_char * 65536)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    need = lib.dsrl_ce_saved_bytes(2, 64)
    assert need >= 2 * 64 * 8 and lib.dsrl_ce_saved_bytes(-1, 64) == 0
    # argument validation happens before the device probe
    assert lib.dsrl_ce_forward(None, p, _lib.U8, 2, 19, 64, 255, _lib.REDUCE_MEAN, p, p, need, None) == _lib.ERR_BAD_ARG
    assert lib.dsrl_ce_forward(p, p, _lib.U8, 2, 19, 64, 255, _lib.REDUCE_NONE, p, p, need, None) == _lib.ERR_UNSUPPORTED
    assert lib.dsrl_ce_forward(p, p, _lib.U8, 2, 0, 64, 255, _lib.REDUCE_MEAN, p, p, need, None) == _lib.ERR_BAD_SHAPE
    assert lib.dsrl_ce_forward(p, p, _lib.U8, 2, 19, 64, 255, _lib.REDUCE_MEAN, p, p, need - 1, None) == _lib.ERR_BAD_ARG
    assert lib.dsrl_ce_backward(p, p, _lib.U8, 2, 19, 64, 255, _lib.REDUCE_SUM, p, need, None, p, None) == _lib.ERR_BAD_ARG
    if not torch.cuda.is_available():
        assert lib.dsrl_ce_forward(p, p, _lib.U8, 2, 19, 64, 255, _lib.REDUCE_MEAN, p, p, need, None) == _lib.ERR_CUDA
        assert b"no CPU fallback" in lib.dsrl_last_error()


def test_missing_library_fails_loudly(tmp_path):
    """No silent fallback: with the shared library absent the first use raises and says how to build it."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from dualsuperreslearningforsemseg_b200 import _lib\n"
            "try:\n    _lib.lib()\nexcept RuntimeError as e:\n    print('RAISED', 'no CPU or PyTorch fallback' in str(e))\n" % ROOT)
    env = dict(os.environ, DSRL_B200_LIB=str(tmp_path / "nope.so"))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert "RAISED True" in out.stdout, out.stdout + out.stderr
