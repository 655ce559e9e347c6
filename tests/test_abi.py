"""CPU: the C-ABI library loads without a GPU and exports exactly what include/bihome_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'bihome_b200.h')


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(bh_[a-z0-9_]+)\s*\(', src)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    if not os.path.isfile(os.path.join(ROOT, 'bihome_b200', 'libbihome_b200.so')):
        g.build()
    from bihome_b200 import cabi
    return cabi.lib()


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for n in ('bh_dlt4_fwd', 'bh_dlt4_bwd', 'bh_warp_fwd', 'bh_warp_bwd', 'bh_bihome_fwd_bwd', 'bh_dltn_fwd', 'bh_dltn_bwd',
              'bh_pairgen_draw', 'bh_pairgen_apply', 'bh_mace', 'bh_strerror', 'bh_version'):
        assert n in names


def test_library_exports_every_declared_symbol(lib):
    from bihome_b200 import cabi
    names = declared_functions()
    assert sorted(cabi.SIGNATURES) == names, 'ctypes table and header disagree'
    raw = ctypes.CDLL(cabi.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), n


def test_version_and_error_text(lib):
    assert lib.bh_version() == 100
    assert lib.bh_strerror(0) == b'ok'
    assert b'NULL' in lib.bh_strerror(-1)
    assert b'unknown' in lib.bh_strerror(-99)


def test_argument_checks_do_not_need_a_gpu(lib):
    """bad arguments are rejected before any launch (codes < 0), so this runs on the CPU box"""
    assert lib.bh_dlt4_fwd(None, None, None, 4, 128.0, 128.0, None) == -1
    assert lib.bh_dlt4_fwd(None, ctypes.c_void_p(16), ctypes.c_void_p(16), 0, 128.0, 128.0, None) == -2
    assert lib.bh_warp_fwd(None, None, None, None, 1, 1, 8, 8, 8, 8, 4, 0, None) == -1
    assert lib.bh_warp_fwd(ctypes.c_void_p(8), ctypes.c_void_p(16), ctypes.c_void_p(16), None, 1, 1, 8, 8, 8, 8, 0, 0, None) == -3
    assert lib.bh_mace(None, None, None, 1, None) == -1
    assert lib.bh_pairgen_draw(ctypes.c_void_p(16), ctypes.c_void_p(16), 4, 1, 100, 100, 32, 128, 32.0, 1, 0, None) == -2
    assert lib.bh_warp_bwd_workspace_bytes(4, 1, 240, 320, 240, 320, 0) > 0
    # K6: compiled geometry, grid sizing (a host function) and argument checks
    assert lib.bh_fieldhead_supported(16, 128) == 1 and lib.bh_fieldhead_supported(64, 512) == 0
    assert lib.bh_fieldhead_grid(0, 100) == 1 and lib.bh_fieldhead_grid(1, 100) == 4 and lib.bh_fieldhead_grid(1, 0) == 0
    assert lib.bh_fieldhead_grid(1, 256 * 128 * 128) == 148 * 5
    p = ctypes.c_void_p(64)
    assert lib.bh_fieldhead_fwd(None, p, p, p, p, p, 1, 16, 16, 128, 0, None) == -1
    assert lib.bh_fieldhead_fwd(p, p, p, p, p, p, 1, 16, 64, 512, 0, None) == -5
    assert lib.bh_fieldhead_bwd(p, p, p, p, p, p, p, 0, 16, 16, 128, 0, None) == -2
    assert lib.bh_fieldhead_moments(ctypes.c_void_p(8), p, 10, 16, None) == -3
    assert lib.bh_fieldhead_affine(p, p, p, None, 10, 16, 1, None) == -1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from bihome_b200 import cabi
    monkeypatch.setattr(cabi, '_lib', None)
    monkeypatch.setattr(cabi, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(ImportError, match='no CPU fallback'):
        cabi.lib()
