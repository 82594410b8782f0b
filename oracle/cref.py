"""ctypes wrapper of oracle/ws_oracle.c (TEST INFRASTRUCTURE ONLY; see oracle/ref.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "libws_oracle.so")
_lib = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        vp, i64, dbl = C.c_void_p, C.c_int64, C.c_double
        L.orc_exp_norm.argtypes = [vp, i64, vp]
        L.orc_ess_perc.argtypes = [vp, i64]
        L.orc_ess_perc.restype = dbl
        L.orc_logsumexp.argtypes = [vp, i64]
        L.orc_logsumexp.restype = dbl
        L.orc_stratified_us.argtypes = [vp, i64, vp]
        L.orc_icdf.argtypes = [vp, vp, i64, vp]
        L.orc_icdf.restype = i64
        L.orc_ssm2d_run.argtypes = [i64, i64, vp, C.c_uint64, dbl, vp, vp]
        L.orc_ssm2d_run.restype = dbl
        L.orc_lgssm1d_run.argtypes = [i64, i64, vp, dbl, dbl, dbl, dbl, C.c_uint64, dbl, vp, vp]
        L.orc_lgssm1d_run.restype = dbl
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def exp_norm(logw):
    a = np.ascontiguousarray(logw, dtype=np.float64)
    out = np.empty_like(a)
    lib().orc_exp_norm(_p(a), a.size, _p(out))
    return out


def ess_perc(w):
    a = np.ascontiguousarray(w, dtype=np.float64)
    return lib().orc_ess_perc(_p(a), a.size)


def logsumexp(logw):
    a = np.ascontiguousarray(logw, dtype=np.float64)
    return lib().orc_logsumexp(_p(a), a.size)


def stratified_us(r):
    a = np.ascontiguousarray(r, dtype=np.float64)
    out = np.empty_like(a)
    lib().orc_stratified_us(_p(a), a.size, _p(out))
    return out


def icdf(w, us):
    w = np.ascontiguousarray(w, dtype=np.float64)
    us = np.ascontiguousarray(us, dtype=np.float64)
    idx = np.empty(w.size, dtype=np.int64)
    clamped = lib().orc_icdf(_p(w), _p(us), w.size, _p(idx))
    return idx, clamped


def ssm2d_run(n, obs, seed=1, ess_perc_min=1.0):
    obs = np.ascontiguousarray(obs, dtype=np.float64)
    mean = np.zeros(2)
    nres = C.c_int64()
    le = lib().orc_ssm2d_run(int(n), obs.shape[0], _p(obs), int(seed), float(ess_perc_min), _p(mean), C.byref(nres))
    return le, mean, nres.value


def lgssm1d_run(n, ys, a=0.9, q=1.0, r=0.5, x0_std=1.0, seed=1, ess_perc_min=1.0):
    ys = np.ascontiguousarray(ys, dtype=np.float64)
    mean = np.zeros(1)
    nres = C.c_int64()
    le = lib().orc_lgssm1d_run(int(n), ys.size, _p(ys), a, q, r, x0_std, int(seed), float(ess_perc_min), _p(mean),
                               C.byref(nres))
    return le, mean[0], nres.value
