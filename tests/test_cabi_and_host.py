"""CPU-side checks: the C-ABI library loads and exports every symbol that
include/p1b200.h declares (no compute calls), host logic (cubature tables,
scipy's index-dtype rule, error mapping)."""
import ctypes
import os
import re

import numpy as np
import pytest

from p1afempy_b200 import _lib, cubature
from p1afempy_b200.device import scipy_index_dtype

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(REPO, 'include', 'p1b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(p1_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), name
    # and the Python binding knows all of them
    assert set(names) == set(_lib.PROTOTYPES), set(names) ^ set(_lib.PROTOTYPES)


def test_ctypes_prototypes_match_the_header():
    """Same number of arguments, and pointer / 64-bit / 32-bit in the same places:
    a mismatch here would not fail loudly at call time."""
    text = open(os.path.join(REPO, 'include', 'p1b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    text = re.sub(r'^\s*#.*$', '', text, flags=re.M)
    decls = re.findall(r'\b(?:int|int64_t|const char \*)\s*(p1_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;', text, flags=re.S)
    seen = {}
    for name, args in decls:
        args = ' '.join(args.split())
        params = [] if args in ('', 'void') else [a.strip() for a in args.split(',')]
        seen[name] = params
    assert set(seen) == set(_lib.PROTOTYPES), set(seen) ^ set(_lib.PROTOTYPES)
    for name, params in seen.items():
        proto = _lib.PROTOTYPES[name]
        assert len(proto) == len(params), (name, params, proto)
        for c_decl, ct in zip(params, proto):
            is_ptr = '*' in c_decl
            if is_ptr:
                assert ct is ctypes.c_void_p or isinstance(ct, type(ctypes.POINTER(ctypes.c_int))) \
                    or ct is ctypes.c_char_p, (name, c_decl, ct)
            elif c_decl.startswith('int64_t'):
                assert ct is ctypes.c_int64, (name, c_decl, ct)
            elif c_decl.startswith('int '):
                assert ct is ctypes.c_int, (name, c_decl, ct)


def test_no_compute_without_gpu_is_loud():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from p1afempy_b200 import solvers
    with pytest.raises(_lib.P1Error):
        solvers.get_stiffness_matrix(np.zeros((3, 2)), np.array([[0, 1, 2]]))
    lib = _lib.load()
    handle = ctypes.c_void_p()
    assert lib.p1_ctx_create(0, ctypes.byref(handle)) != 0
    assert _lib.last_error()


def test_error_mapping():
    assert _lib.check(0) is None
    for code, exc in ((_lib.P1_EINDEX, IndexError), (_lib.P1_EINVAL, ValueError),
                      (_lib.P1_ETOPOLOGY, ValueError), (_lib.P1_ECUDA, _lib.P1Error)):
        with pytest.raises(exc):
            _lib.check(code)


def test_cubature_rules_integrate_monomials():
    """weights sum to 1/2; MIDPOINT is (1/3,1/3); the DAYTAYLOR stand-in is
    exact to total degree 6: int x^a y^b = a! b! / (a+b+2)!."""
    from math import factorial
    w, p = cubature.resolve_rule(cubature.CubatureRuleEnum.MIDPOINT)
    assert np.array_equal(w, [0.5]) and np.allclose(p, [[1 / 3, 1 / 3]])
    for rule, degree in ((cubature.CubatureRuleEnum.LAUFFER_LINEAR, 1),
                         (cubature.CubatureRuleEnum.SMPLX1, 2),
                         (cubature.CubatureRuleEnum.DAYTAYLOR, 6)):
        w, p = cubature.resolve_rule(rule)
        assert abs(w.sum() - 0.5) < 1e-15
        for a in range(degree + 1):
            for b in range(degree + 1 - a):
                exact = factorial(a) * factorial(b) / factorial(a + b + 2)
                assert abs(np.sum(w * p[:, 0] ** a * p[:, 1] ** b) - exact) < 1e-14


def test_rule_is_data():
    w, p = cubature.resolve_rule(([0.25, 0.25], [[0.2, 0.2], [0.4, 0.3]]))
    assert w.dtype == np.float64 and p.shape == (2, 2)

    class Holder:
        weights = np.array([0.5])
        integration_points = np.array([[1 / 3, 1 / 3]])
    assert cubature.resolve_rule(Holder())[0].shape == (1,)
    with pytest.raises(ValueError):
        cubature.resolve_rule(([0.5, 0.5], [[0.1, 0.1]]))


def test_scipy_index_dtype_rule():
    """scipy: int32 unless max(nnz_coo, N) exceeds int32 (scipy/sparse/_coo.py:419);
    checked against scipy itself at small sizes, by formula at the 256M point."""
    from scipy.sparse import coo_matrix
    e = np.array([[0, 1, 2]])
    a = coo_matrix((np.ones(9), (np.repeat(e, 3), np.tile(e, 3).ravel()))).tocsr()
    assert a.indices.dtype == scipy_index_dtype(1, 3, 3) == np.int32
    assert scipy_index_dtype(67108864, 33562625, 33562625) == np.int32     # 64M mesh
    assert scipy_index_dtype(268435456, 134234113, 134234113) == np.int64  # 256M mesh


def test_cubature_stand_ins_are_named_and_announced():
    """DAYTAYLOR / SMPLX1 are tables of the third-party triangle_cubature package:
    offline they resolve to the rules this package ships under their own names,
    and say so (once per rule)."""
    import warnings
    from p1afempy_b200 import cubature as cub
    cub._warned.clear()
    with warnings.catch_warnings(record=True) as seen:
        warnings.simplefilter('always')
        w, p = cub.resolve_rule(cub.CubatureRuleEnum.DAYTAYLOR)
        cub.resolve_rule(cub.CubatureRuleEnum.DAYTAYLOR)
        w2, p2 = cub.resolve_rule(cub.CubatureRuleEnum.CONICAL_GAUSS_4X4)
    assert [x.category for x in seen] == [cub.StandInRuleWarning]
    assert np.array_equal(w, w2) and np.array_equal(p, p2)
    assert abs(w.sum() - 0.5) < 1e-15 and w.shape == (16,)
    # degree 6: x^a y^b, a + b <= 6, integrates to a! b! / (a + b + 2)!
    from math import factorial as f
    for a in range(7):
        for b in range(7 - a):
            exact = f(a) * f(b) / f(a + b + 2)
            assert abs((w * p[:, 0] ** a * p[:, 1] ** b).sum() - exact) < 1e-15
    w1, p1 = cub.resolve_rule(cub.CubatureRuleEnum.MIDPOINT)
    assert np.array_equal(w1, [0.5]) and np.array_equal(p1, [[1. / 3., 1. / 3.]])


def test_io_helpers(tmp_path):
    """The reference's text readers (io_helpers.py:5-132: uint32 indices, optional
    1-based shift, a single boundary edge keeps two dimensions) and the binary
    save_mesh / load_mesh round trip."""
    from p1afempy_b200 import io_helpers as io
    c = np.array([[0., 0.], [1., 0.], [0., 1.], [1., 1.]])
    e = np.array([[1, 2, 3], [2, 4, 3]])
    np.savetxt(tmp_path / 'c.dat', c)
    np.savetxt(tmp_path / 'e.dat', e, fmt='%d')
    np.savetxt(tmp_path / 'b.dat', np.array([[1, 2]]), fmt='%d')
    cc, ee = io.read_mesh(path_to_coordinates=tmp_path / 'c.dat',
                          path_to_elements=tmp_path / 'e.dat', shift_indices=True)
    assert np.array_equal(cc, c) and ee.dtype == np.uint32 and np.array_equal(ee, e - 1)
    b = io.read_boundary_condition(tmp_path / 'b.dat', shift_indices=True)
    assert b.shape == (1, 2) and b.dtype == np.uint32 and np.array_equal(b, [[0, 1]])
    io.save_mesh(tmp_path / 'mesh', c, e - 1, [b, np.zeros((0, 2), dtype=int)])
    c2, e2, b2 = io.load_mesh(tmp_path / 'mesh.npz')
    assert np.array_equal(c2, c) and np.array_equal(e2, e - 1) and e2.dtype == np.int32
    assert len(b2) == 2 and np.array_equal(b2[0], b) and b2[1].shape == (0, 2)
