"""ctypes binding of libbihome_b200.so (include/bihome_b200.h).  Fails loudly when the library is missing."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libbihome_b200.so')

_vp, _i, _f, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t
_d = ctypes.c_double
_u64 = ctypes.c_uint64

# name -> (restype, argtypes); mirrors include/bihome_b200.h one to one
SIGNATURES = {
    'bh_version': (_i, []),
    'bh_strerror': (ctypes.c_char_p, [_i]),
    'bh_launch_count': (ctypes.c_ulonglong, []),
    'bh_dlt4_fwd': (_i, [_vp, _vp, _vp, _i, _f, _f, _vp]),
    'bh_dlt4_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _f, _f, _vp]),
    'bh_warp_fwd': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    'bh_warp_bwd_workspace_bytes': (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    'bh_warp_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    'bh_bihome_fwd_bwd': (_i, [_vp] * 10 + [_f] + [_vp] * 10 + [_i, _i, _i, _i, _i, _vp]),
    'bh_bihome_rescale': (_i, [_vp] * 9 + [_i, _i, _i, _i, _vp]),
    'bh_triplet_fwd_bwd': (_i, [_vp] * 10 + [_i, _i, _i, _i] + [_f] * 5 + [_vp] * 12 + [_i, _i, _i, _i, _i, _vp]),
    'bh_triplet_rescale': (_i, [_vp] * 11 + [_i, _i, _i, _i, _vp]),
    'bh_dltn_fwd': (_i, [_vp] * 7 + [_i, _i, _i, _i, _vp]),
    'bh_dltn_bwd': (_i, [_vp] * 9 + [_i, _i, _i, _i, _vp]),
    'bh_pairgen_draw': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _f, _u64, _u64, _vp]),
    'bh_pairgen_apply': (_i, [_vp] * 6 + [_i, _i, _i, _i, _i, _d, _d, _vp]),
    'bh_pairgen_image': (_i, [_vp] * 4 + [_i, _i, _i, _i, _d, _d, _vp]),
    'bh_mace': (_i, [_vp, _vp, _vp, _i, _vp]),
    'bh_tune_set': (_i, [ctypes.c_char_p, _i]),
    'bh_fieldhead_supported': (_i, [_i, _i]),
    'bh_fieldhead_grid': (_i, [_i, ctypes.c_longlong]),
    'bh_fieldhead_moments': (_i, [_vp, _vp, ctypes.c_longlong, _i, _vp]),
    'bh_fieldhead_fwd': (_i, [_vp] * 6 + [_i, _i, _i, _i, _i, _vp]),
    'bh_fieldhead_bwd': (_i, [_vp] * 7 + [_i, _i, _i, _i, _i, _vp]),
    'bh_fieldhead_affine': (_i, [_vp] * 4 + [ctypes.c_longlong, _i, _i, _vp]),
    'bh_stem_supported': (_i, [_i]),
    'bh_stem_workspace_bytes': (_sz, [_i]),
    'bh_stem_fwd': (_i, [_vp] * 5 + [_f, _f] + [_vp] * 4 + [_sz, _i, _i, _i, _i, _vp]),
    'bh_stem_bwd': (_i, [_vp] * 8 + [_sz, _i, _i, _i, _i, _vp]),
    'bh_bnact_fwd': (_i, [_vp] * 6 + [_f, _f] + [_vp] * 3 + [_sz, ctypes.c_longlong, _i, _vp]),
    'bh_bnact_bwd': (_i, [_vp] * 9 + [_sz, ctypes.c_longlong, _i, _vp]),
    'bh_bnact2_fwd': (_i, [_vp] * 6 + [_f, _f] + [_vp] * 4 + [_f, _f] + [_vp] * 4 + [_sz, ctypes.c_longlong, _i, _vp]),
    'bh_bnact2_bwd': (_i, [_vp] * 13 + [_sz, ctypes.c_longlong, _i, _vp]),
    'bh_bias_add': (_i, [_vp, _vp, ctypes.c_longlong, _i, _vp]),
    'bh_bias_grad': (_i, [_vp, _vp, _vp, _sz, ctypes.c_longlong, _i, _vp]),
}

_lib = None


def lib():
    """The loaded library (cached).  Raises ImportError with the build recipe if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                'bihome_b200: %s is missing -- build it with `make -C bihome_b200/csrc` or '
                '`python -c "import __graft_entry__ as g; g.build()"`; there is no CPU fallback.' % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)       # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


class BihomeError(RuntimeError):
    pass


def check(code, what):
    if code != 0:
        raise BihomeError('%s failed (%d): %s' % (what, code, lib().bh_strerror(code).decode()))


def launch_count():
    return int(lib().bh_launch_count())
