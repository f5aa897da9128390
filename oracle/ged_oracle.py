"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the CPU oracle in ``oracle/ged_oracle.c``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module, and only as the checker.  The product path (``ppo-bihyb_b200/``) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libged_oracle.so")
_lib = None

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_f32p = ctypes.POINTER(ctypes.c_float)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_i64p = ctypes.POINTER(ctypes.c_int64)


class _Trace(ctypes.Structure):
    _fields_ = [("s_vv", c_f64p), ("s_vb", c_f64p), ("ub", c_f64p), ("alpha", c_f64p), ("beta", c_f64p),
                ("t0", c_f64p), ("b_rowmatch", c_i32p), ("b_colmatch", c_i32p), ("cost", c_f32p),
                ("x_after", c_f32p), ("init_rowmatch", c_i32p), ("init_colmatch", c_i32p),
                ("init_cost", c_f32p), ("lap_steps", c_i64p)]


def build(force=False):
    """Compile the oracle with the recipe in oracle/Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "ged_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.ged_oracle_version.restype = ctypes.c_char_p
        _lib.ged_oracle_score.restype = ctypes.c_double
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def lsap(cost):
    """scipy.optimize.linear_sum_assignment restatement.  Returns (col4row, steps)."""
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    nr, nc = cost.shape
    c4r = np.empty(nr, np.int32)
    r4c = np.empty(nc, np.int32)
    steps = ctypes.c_int64(0)
    st = lib().ged_oracle_lsap(nr, nc, _p(cost, c_f64p), _p(c4r, c_i32p), _p(r4c, c_i32p), ctypes.byref(steps))
    if st:
        raise ValueError(f"lsap status {st}")
    return c4r, steps.value


def lsape(C):
    """hungarian_lap restatement (ged_env.py:329-351).  C is (n1+1)x(n2+1) fp32.  -> rowmatch, colmatch, lb, steps"""
    C = np.ascontiguousarray(C, dtype=np.float32)
    n1, n2 = C.shape[0] - 1, C.shape[1] - 1
    rm = np.empty(max(n1, 1), np.int32)
    cm = np.empty(max(n2, 1), np.int32)
    lb = ctypes.c_double(0)
    steps = ctypes.c_int64(0)
    st = lib().ged_oracle_lsape(n1, n2, _p(C, c_f32p), _p(rm, c_i32p), _p(cm, c_i32p), ctypes.byref(lb), ctypes.byref(steps))
    if st:
        raise ValueError(f"lsape status {st}")
    return rm[:n1], cm[:n2], lb.value, steps.value


def match_to_dense(rowmatch, colmatch):
    n1, n2 = len(rowmatch), len(colmatch)
    x = np.zeros((n1 + 1, n2 + 1), np.float32)
    x[np.arange(n1), rowmatch] = 1
    ins = np.nonzero(np.asarray(colmatch) == n1)[0]
    x[n1, ins] = 1
    return x


def label_cost(lab1, lab2):
    lab1 = np.ascontiguousarray(lab1, np.int32)
    lab2 = np.ascontiguousarray(lab2, np.int32)
    n1, n2 = len(lab1), len(lab2)
    out = np.empty((n1 + 1, n2 + 1), np.float32)
    lib().ged_oracle_label_cost(n1, n2, _p(lab1, c_i32p), _p(lab2, c_i32p), _p(out, c_f32p))
    return out


def init_cost_closed_form(adj1, adj2):
    adj1, adj2 = _u8(adj1), _u8(adj2)
    n1, n2 = adj1.shape[0], adj2.shape[0]
    out = np.empty((n1 + 1, n2 + 1), np.float32)
    lib().ged_oracle_init_cost_closed_form(n1, n2, _p(adj1, c_u8p), _p(adj2, c_u8p), _p(out, c_f32p))
    return out


def init_cost_bruteforce(adj1, adj2, Cn):
    adj1, adj2 = _u8(adj1), _u8(adj2)
    Cn = np.ascontiguousarray(Cn, np.float32)
    n1, n2 = adj1.shape[0], adj2.shape[0]
    out = np.empty((n1 + 1, n2 + 1), np.float32)
    st = lib().ged_oracle_init_cost_bruteforce(n1, n2, _p(adj1, c_u8p), _p(adj2, c_u8p), _p(Cn, c_f32p), _p(out, c_f32p))
    if st:
        raise ValueError(f"status {st}")
    return out


def kv(adj1, adj2, Cn, X):
    adj1, adj2 = _u8(adj1), _u8(adj2)
    Cn = np.ascontiguousarray(Cn, np.float32)
    X = np.ascontiguousarray(X, np.float32)
    n1, n2 = adj1.shape[0], adj2.shape[0]
    out = np.empty((n1 + 1, n2 + 1), np.float32)
    lib().ged_oracle_kv(n1, n2, _p(adj1, c_u8p), _p(adj2, c_u8p), _p(Cn, c_f32p), _p(X, c_f32p), _p(out, c_f32p))
    return out


def score(adj1, adj2, Cn, rowmatch, colmatch):
    adj1, adj2 = _u8(adj1), _u8(adj2)
    Cn = np.ascontiguousarray(Cn, np.float32)
    rm = np.ascontiguousarray(rowmatch, np.int32)
    cm = np.ascontiguousarray(colmatch, np.int32)
    n1, n2 = adj1.shape[0], adj2.shape[0]
    return lib().ged_oracle_score(n1, n2, _p(adj1, c_u8p), _p(adj2, c_u8p), _p(Cn, c_f32p), _p(rm, c_i32p), _p(cm, c_i32p))


def solve(adj1, adj2, Cn, edit=None, max_iter=100, solver="ipfp", trace=False, x_init=None):
    """Canonical-order ipfp_ged / hungarian_ged + scoring on the unedited pair.  ``x_init`` ((n1+1) x (n2+1) fp32) replaces
    the Hungarian start (teacher forcing, like ``ged_solve_desc.x_init``).

    Returns dict(rowmatch, colmatch, ged, ged_edited, iters[, trace...]).
    """
    adj1, adj2 = _u8(adj1), _u8(adj2)
    Cn = np.ascontiguousarray(Cn, np.float32)
    n1, n2 = adj1.shape[0], adj2.shape[0]
    R, C = n1 + 1, n2 + 1
    assert Cn.shape == (R, C)
    eu, ev = (-1, -1) if edit is None else (int(edit[0]), int(edit[1]))
    rm = np.empty(n1, np.int32)
    cm = np.empty(n2, np.int32)
    ged = ctypes.c_float(0)
    ged_e = ctypes.c_float(0)
    iters = ctypes.c_int32(0)
    tr = None
    keep = {}
    if trace:
        T = max_iter
        keep = dict(s_vv=np.zeros(T), s_vb=np.zeros(T), ub=np.zeros(T), alpha=np.zeros(T), beta=np.zeros(T),
                    t0=np.zeros(T), b_rowmatch=np.zeros((T, n1), np.int32), b_colmatch=np.zeros((T, n2), np.int32),
                    cost=np.zeros((T, R, C), np.float32), x_after=np.zeros((T, R, C), np.float32),
                    init_rowmatch=np.zeros(n1, np.int32), init_colmatch=np.zeros(n2, np.int32),
                    init_cost=np.zeros((R, C), np.float32), lap_steps=np.zeros(1, np.int64))
        tr = _Trace(*[_p(keep[name], typ) for name, typ in _Trace._fields_])
    if x_init is not None:
        x_init = np.ascontiguousarray(x_init, np.float32)
        assert x_init.shape == (R, C)
    st = lib().ged_oracle_solve_from(n1, n2, _p(adj1, c_u8p), _p(adj2, c_u8p), _p(Cn, c_f32p), eu, ev, int(max_iter),
                                0 if solver == "ipfp" else 1, _p(x_init, c_f32p) if x_init is not None else None,
                                _p(rm, c_i32p), _p(cm, c_i32p),
                                ctypes.byref(ged), ctypes.byref(ged_e), ctypes.byref(iters),
                                ctypes.byref(tr) if tr is not None else None)
    if st:
        raise ValueError(f"oracle solve status {st}")
    out = dict(rowmatch=rm, colmatch=cm, ged=ged.value, ged_edited=ged_e.value, iters=iters.value)
    if trace:
        n = iters.value
        for k_, v_ in keep.items():
            out[k_] = v_[:n] if k_ in ("s_vv", "s_vb", "ub", "alpha", "beta", "t0", "b_rowmatch", "b_colmatch", "cost", "x_after") else v_
        out["lap_steps"] = int(keep["lap_steps"][0])
    return out
