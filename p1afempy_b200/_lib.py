"""ctypes binding of libp1b200.so (include/p1b200.h).

There is NO CPU fallback: if the shared library is missing or cannot be loaded
every entry point of the package raises.  ``python __graft_entry__.py`` (or
``make -C p1afempy_b200/csrc``) builds it in-tree.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libp1b200.so')

P1_OK = 0
P1_EINVAL = -1
P1_EINDEX = -2
P1_ECUDA = -3
P1_ETOPOLOGY = -4
P1_ERANGE = -5
P1_ECOMM = -6
P1_ERETRY = -7

STATUS_INDEX = 1
STATUS_RETRY = 16
STATUS_SLOWPATH = 64

PLAN_DEFAULT = 0
PLAN_EXACT = 1
PLAN_KEEP_ELEMENTS = 2
PLAN_TOPOLOGY = 4
PLAN_SLOT_INDEX = 8
PLAN_FORCE_RECORDS = 16
PLAN_FORCE_SLIM = 32
PLAN_TINY_SPILL = 64
PLAN_TILED = 128

OP_STORED = 0
OP_STIFFNESS = 1
OP_MASS = 2
OP_LOCAL = 3
OP_GENERAL = 4
OP_WEIGHTED = 5

_vp = C.c_void_p
_i64 = C.c_int64
_int = C.c_int

class QuadArgs(C.Structure):
    """p1_quad_args (include/p1b200.h)."""
    _fields_ = [('a11q', _vp), ('a12q', _vp), ('a21q', _vp), ('a22q', _vp), ('phiq', _vp),
                ('weights_host', _vp), ('points_host', _vp), ('n_qp', _int)]


# name -> argtypes; every function returns int unless listed in _RESTYPES
PROTOTYPES = {
    'p1_version': [],
    'p1_last_error': [],
    'p1_ctx_create': [_int, C.POINTER(_vp)],
    'p1_ctx_destroy': [_vp],
    'p1_ctx_trim': [_vp],
    'p1_ctx_set_cache_limit': [_vp, _i64],
    'p1_profile_enable': [_vp, _int],
    'p1_profile_reset': [_vp],
    'p1_profile_report': [_vp, C.c_char_p, _i64],
    'p1_narrow_indices': [_vp, _vp, _int, _i64, _vp, _vp],
    'p1_check_indices': [_vp, _vp, _i64, _i64, _vp],
    'p1_widen_indices': [_vp, _vp, _i64, _vp, _vp],
    'p1_plan_create': [_vp, _vp, _i64, _i64, _i64, _i64, _int, _vp, C.POINTER(_vp)],
    'p1_plan_create_for': [_vp, _vp, _i64, _i64, _i64, _i64, _int, _int, _vp, _vp, _vp,
                           C.POINTER(_vp)],
    'p1_plan_create_quad': [_vp, _vp, _i64, _i64, _i64, _i64, _int, _int, _vp,
                            C.POINTER(QuadArgs), _vp, C.POINTER(_vp)],
    'p1_plan_create_interleaved': [_vp, _vp, _i64, _i64, _int, _int, _int, _int, _int, _vp,
                                   _vp, _vp, C.POINTER(_vp)],
    'p1_plan_rows': [_vp],
    'p1_plan_destroy': [_vp],
    'p1_partition_rows': [_vp, _vp, _i64, _i64, _int, _vp, _vp],
    'p1_plan_info': [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64),
                     C.POINTER(_i64)],
    'p1_assemble_csr': [_vp, _int, _vp, _vp, _vp, _vp, _vp, _vp],
    'p1_element_neighbours': [_vp, _vp, _vp],
    'p1_chunk_nnz': [_vp, _vp, _i64, _int, _i64, _vp, _vp],
    'p1_chunk_offsets': [_vp, _vp, _int, _i64, _int, _vp, _vp, _vp],
    'p1_block_submesh': [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp],
    'p1_block_submesh_fill': [_vp, _vp, _i64, _vp, _vp, _vp, _vp],
    'p1_eta_r_partial': [_vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp],
    'p1_nvb_sort_longest_edge': [_vp, _vp, _vp, _i64, _vp],
    'p1_nvb_plan': [_vp, _i64, _i64, _vp, _vp, _i64, _vp, _vp, _int, _vp, _vp, _vp, _vp, _vp,
                    _vp],
    'p1_nvb_apply': [_vp, _vp, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _int, _vp, _vp,
                     _vp, _vp, _vp, _vp, _int, _vp, _vp, _vp, _vp, _vp],
    'p1_apply_neumann': [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp],
    'p1_csr_spmv': [_vp, _i64, _vp, _vp, _vp, _vp, C.c_double, _vp, _vp, _vp],
    'p1_dot': [_vp, _i64, _vp, _vp, C.POINTER(C.c_double), _vp],
    'p1_pcg': [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, C.c_double, _int,
               C.POINTER(_int), C.POINTER(C.c_double), _vp],
    'p1_assemble_cold': [_vp, _vp, _i64, _i64, _int, _int, _int, _int, _vp, _vp,
                         C.POINTER(QuadArgs), _i64, _vp, _vp, _vp, _vp, _vp],
    'p1_mass_triplets': [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp],
    'p1_stiffness_triplets': [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp],
    'p1_geometry': [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp],
    'p1_quadrature_points': [_vp, _vp, _vp, _i64, _vp, _int, _vp, _vp],
    'p1_nodal_at_quadrature': [_vp, _vp, _vp, _i64, _vp, _int, _vp, _vp],
    'p1_centroids': [_vp, _vp, _vp, _i64, _vp, _vp],
    'p1_local_general_stiffness': [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp,
                                   _vp, _int, _vp, _vp],
    'p1_local_weighted_mass': [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _int, _vp,
                               _vp],
    'p1_integrate_quadrature': [_vp, _vp, _vp, _i64, _vp, _vp, _int, C.POINTER(C.c_double),
                                _vp],
    'p1_load_vector': [_vp, _vp, _vp, _vp, _vp, _int, _vp, _vp],
    'p1_load_vector_midpoint': [_vp, _vp, _vp, _vp, _vp],
    'p1_load_vector_direct': [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _int, _vp, _vp],
    'p1_load_vector_midpoint_direct': [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp],
    'p1_geometric_data': [_vp, _vp, _vp, _int, _vp, _vp, _vp, C.POINTER(_i64),
                          _vp],
    'p1_eta_r': [_vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp],
    'p1_boundary_edges': [_vp, _vp, _vp, _i64, _vp, _vp, _vp],
}
_RESTYPES = {'p1_last_error': C.c_char_p, 'p1_plan_rows': _i64}

_lib = None
_lock = threading.Lock()


class P1Error(RuntimeError):
    pass


def load():
    """Loads the shared library once; raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise P1Error(
                'libp1b200.so is not built (%s missing). Build it with '
                '`make -C p1afempy_b200/csrc` or `python -c "import '
                '__graft_entry__ as g; g.build()"`. There is no CPU fallback.'
                % LIB_PATH)
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, argtypes in PROTOTYPES.items():
            fn = getattr(lib, name)      # AttributeError if a symbol is missing
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, _int)
        _lib = lib
    return _lib


def last_error() -> str:
    msg = load().p1_last_error()
    return msg.decode('utf-8', 'replace') if msg else ''


def check(rc: int):
    """Maps a status code to the exception the reference would raise
    (SURVEY.md section 8(b), row 'Errors')."""
    if rc == P1_OK:
        return
    msg = last_error()
    if rc == P1_EINDEX:
        raise IndexError(msg)
    if rc in (P1_EINVAL, P1_ETOPOLOGY, P1_ERANGE):
        raise ValueError(msg)
    raise P1Error('libp1b200 error %d: %s' % (rc, msg))
