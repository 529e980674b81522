"""Python driver of the FAKE MATLAB C Matrix API (tests/mex_stub): builds mxArrays from numpy
objects, calls the UNMODIFIED MEX shim p-dmpc_b200/matlab/pdmpc_b200_mex.cpp and converts the
results back.  Stands in for MATLAB, which this image does not have."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tests", "mex_stub")
LIB = os.path.join(STUB, "libpdmpc_mex_fake.so")
SRC = [os.path.join(ROOT, "p-dmpc_b200", "matlab", "pdmpc_b200_mex.cpp"), os.path.join(STUB, "fake_matlab.cpp")]
CSRC = os.path.join(ROOT, "p-dmpc_b200", "csrc")

(CREATE, DESTROY, UPLOAD_MPA, PLAN, STATS, PLAN_SAMPLED, PLAN_TIMESTEP, PLAN_JOINT, UPLOAD_ROAD, SAMPLE_INPUTS,
 UPLOAD_REACHABLE_SETS, ASSEMBLE_OBSTACLES) = range(12)
_CELL, _STRUCT, _LOGICAL, _DOUBLE, _UINT64 = 1, 2, 3, 6, 13


def build() -> str:
    deps = SRC + [os.path.join(STUB, "mex.h"), os.path.join(ROOT, "include", "pdmpc_b200.h"),
                  os.path.join(CSRC, "libpdmpc_b200.so")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-fPIC", "-shared", "-I", STUB, "-I",
                        os.path.join(ROOT, "include"), "-o", LIB] + SRC +
                       ["-L", CSRC, "-lpdmpc_b200", "-Wl,-rpath," + CSRC], check=True)
    return LIB


class MexError(RuntimeError):
    pass


class FakeMatlab:
    def __init__(self):
        L = C.CDLL(build())
        P = C.c_void_p
        L.fm_double.argtypes = [C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_double)]
        L.fm_double.restype = P
        L.fm_cell.argtypes = [C.c_size_t, C.c_size_t]
        L.fm_cell.restype = P
        L.fm_set_cell.argtypes = [P, C.c_size_t, P]
        L.fm_struct.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
        L.fm_struct.restype = P
        L.fm_set_field.argtypes = [P, C.c_char_p, P]
        L.fm_call.argtypes = [C.c_int, C.POINTER(P), C.c_int, C.POINTER(P)]
        L.fm_call.restype = C.c_int
        L.fm_last_error.restype = C.c_char_p
        L.fm_ndim.argtypes = [P]
        L.fm_ndim.restype = C.c_size_t
        L.fm_dim.argtypes = [P, C.c_size_t]
        L.fm_dim.restype = C.c_size_t
        L.fm_class.argtypes = [P]
        L.fm_class.restype = C.c_int
        L.fm_doubles.argtypes = [P]
        L.fm_doubles.restype = C.POINTER(C.c_double)
        L.fm_u64.argtypes = [P]
        L.fm_u64.restype = C.c_uint64
        L.fm_get_cell.argtypes = [P, C.c_size_t]
        L.fm_get_cell.restype = P
        L.fm_get_field.argtypes = [P, C.c_char_p]
        L.fm_get_field.restype = P
        L.fm_free.argtypes = [P]
        L.fm_clear_mex.argtypes = []
        self.L = L

    # ---- numpy/python -> mxArray (column-major, like MATLAB) ----
    def to_mx(self, v):
        L = self.L
        if isinstance(v, dict):
            names = (C.c_char_p * len(v))(*[k.encode() for k in v])
            s = L.fm_struct(len(v), names)
            for k, x in v.items():
                L.fm_set_field(s, k.encode(), self.to_mx(x))
            return s
        if isinstance(v, list):                      # list (1 x n cell) or list of lists (m x n cell)
            if v and isinstance(v[0], list):
                m, n = len(v), len(v[0])
                c = L.fm_cell(m, n)
                for i in range(m):
                    for j in range(n):
                        if v[i][j] is not None:
                            L.fm_set_cell(c, i + m * j, self.to_mx(v[i][j]))
                return c
            c = L.fm_cell(len(v), 1) if v else L.fm_cell(0, 0)
            for i, x in enumerate(v):
                if x is not None:
                    L.fm_set_cell(c, i, self.to_mx(x))
            return c
        a = np.asarray(v, dtype=np.float64)
        if a.ndim == 0:
            a = a.reshape(1, 1)
        elif a.ndim == 1:
            a = a.reshape(1, -1)
        dims = (C.c_size_t * a.ndim)(*a.shape)
        flat = np.ascontiguousarray(a.reshape(-1, order="F"))
        return L.fm_double(a.ndim, dims, flat.ctypes.data_as(C.POINTER(C.c_double)))

    def from_mx(self, p):
        L = self.L
        cls = L.fm_class(p)
        dims = [L.fm_dim(p, i) for i in range(L.fm_ndim(p))]
        n = int(np.prod(dims))
        if cls == _UINT64:
            return int(L.fm_u64(p))
        if cls == _CELL:
            return [self.from_mx(L.fm_get_cell(p, i)) for i in range(n)]
        if cls == _STRUCT:
            raise NotImplementedError("use field()")
        a = np.ctypeslib.as_array(L.fm_doubles(p), shape=(n,)).copy().reshape(dims, order="F") if n else np.zeros(dims)
        if cls == _LOGICAL:
            return bool(a.reshape(-1)[0])
        return a

    def field(self, p, name):
        return self.from_mx(self.L.fm_get_field(p, name.encode()))

    def call(self, nlhs, *args, raw=False):
        """[out1, ...] = pdmpc_b200_mex(args...)"""
        L = self.L
        prhs = (C.c_void_p * len(args))(*[self.to_mx(a) if not isinstance(a, _Raw) else a.p for a in args])
        plhs = (C.c_void_p * max(nlhs, 1))()
        rc = L.fm_call(nlhs, plhs, len(args), prhs)
        for a, p in zip(args, prhs):
            if not isinstance(a, _Raw):
                L.fm_free(p)
        if rc != 0:
            raise MexError(L.fm_last_error().decode())
        if raw:
            return [plhs[i] for i in range(nlhs)]
        out = [self.from_mx(plhs[i]) for i in range(nlhs)]
        for i in range(nlhs):
            L.fm_free(plhs[i])
        return out

    def clear_mex(self):
        self.L.fm_clear_mex()


class _Raw:
    def __init__(self, p):
        self.p = p


def matlab_mpa(mpa):
    """(transition_matrix_single [nT x nT x Hp], maneuvers {nT x nT} of structs) as the MATLAB
    MotionPrimitiveAutomaton object holds them (MotionPrimitiveAutomaton.m:5-17)."""
    nT = mpa.n_trims
    trans = np.transpose(mpa.transition.astype(np.float64), (1, 2, 0))     # (t1, t2, k)
    man = [[None] * nT for _ in range(nT)]
    names = ("area", "area_without_offset", "area_large_offset")
    for e in range(mpa.n_edges):
        d = {"dx": mpa.edge_dx[e], "dy": mpa.edge_dy[e], "dyaw": mpa.edge_dyaw[e]}
        for k, nm in enumerate(names):
            n = int(mpa.area_npts[e, k])
            d[nm] = np.vstack([mpa.area_x[e, k, :n], mpa.area_y[e, k, :n]])
        man[mpa.edge_from[e] - 1][mpa.edge_to[e] - 1] = d
    return trans, man
