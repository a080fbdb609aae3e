"""ctypes loader for oracle/_ref/libqhull_ref.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The library is the reference's own vendored Qhull 2019.1 (spatial/qhull_src/src + spatial/qhull_misc.c) compiled from
/root/reference by oracle/Makefile, with oracle/qhull_ref_driver.c restating the reference's Cython wrapper
(spatial/qhull.pyx:1867-1885, 338-363, 563-568, 573-723).  It pins the one thing the stock-SciPy stand-in of
oracle/reference_port.py could not: how the reference's Qhull splits co-circular lattice points.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libqhull_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(LIB_PATH)
        lib.qhull_ref_version.restype = C.c_char_p
        lib.qhull_ref_delaunay2d.restype = C.c_int
        lib.qhull_ref_delaunay2d.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        lib.qhull_ref_delaunay2d_probe.restype = C.c_int
        lib.qhull_ref_delaunay2d_probe.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib = lib
    return _lib


def version() -> str:
    return _load().qhull_ref_version().decode()


def delaunay(points_rc: np.ndarray):
    """`Delaunay(points)` of interp2d.py:55 through the reference's vendored Qhull.
    Returns (simplices [T,3] int32, neighbors [T,3] int32) exactly as the wrapper's get_simplex_facet_array."""
    p = np.ascontiguousarray(points_rc, dtype=np.float64)
    n = p.shape[0]
    cap = 2 * n + 16
    simp = np.zeros((cap, 3), np.int32)
    nb = np.zeros((cap, 3), np.int32)
    T = _load().qhull_ref_delaunay2d(p.ctypes.data, n, simp.ctypes.data, nb.ctypes.data, cap)
    if T < 0:
        raise RuntimeError(f"reference Qhull failed (rc={T}) on {n} points")
    return simp[:T].copy(), nb[:T].copy()


def delaunay_probe(points_rc: np.ndarray):
    """The same Qhull run with what decides the split of co-circular cells: (simplices [T,3], owner [T], vertex_id [n]);
    owner = -1 for an ordinary facet, else the merged facet `Qt` cut the triangle from; vertex_id = Qhull's insertion
    order of every point (qhull_ref_driver.c)."""
    p = np.ascontiguousarray(points_rc, dtype=np.float64)
    n = p.shape[0]
    cap = 2 * n + 16
    simp = np.zeros((cap, 3), np.int32)
    owner = np.zeros(cap, np.int32)
    vid = np.zeros(n, np.int32)
    T = _load().qhull_ref_delaunay2d_probe(p.ctypes.data, n, simp.ctypes.data, owner.ctypes.data, vid.ctypes.data, cap)
    if T < 0:
        raise RuntimeError(f"reference Qhull failed (rc={T}) on {n} points")
    return simp[:T].copy(), owner[:T].copy(), vid
