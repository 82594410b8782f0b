"""ctypes wrapper of tests/host/harness.cpp: the product's statement lowering + micro-op interpreter
run on the CPU (test infrastructure; see the header of harness.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "host", "_harness.so")
SRC = os.path.join(HERE, "host", "harness.cpp")
_lib = None


def lib():
    global _lib
    if _lib is None:
        deps = [SRC] + [os.path.join(HERE, "..", "weightedsampling.jl_b200", "csrc", f)
                        for f in ("ws_lowering.h", "ws_vm.cuh", "ws_math.cuh", "ws_exchange.h", "ws_vm_sl.cuh")]
        if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-o", SO, SRC],
                           check=True)
        L = C.CDLL(SO)
        L.hh_new.restype = C.c_void_p
        L.hh_new.argtypes = [C.c_int64, C.c_uint64]
        L.hh_error.restype = C.c_char_p
        for name in ("hh_free", "hh_error", "hh_flush", "hh_n_flush", "hh_window_ops", "hh_window_regs",
                     "hh_window_loads", "hh_window_stores", "hh_tape_len", "hh_window_signature"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.hh_col.argtypes = [C.c_void_p, C.c_int]
        L.hh_set_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.hh_get_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.hh_get_logw.argtypes = [C.c_void_p, C.c_void_p]
        L.hh_set_replay.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        L.hh_score.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.hh_count_slots_le.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
        L.hh_randn2.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        L.hh_spacings.argtypes = [C.c_int64, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]
        L.hh_set_replay_variates.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.hh_rand_gamma.argtypes = [C.c_double, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int64]
        L.hh_rand_poisson.argtypes = [C.c_double, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int64]
        L.hh_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        L.hh_importance_normal.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
        L.hh_sample_expr.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class HostStore:
    """Duck-types the bits of DeviceColumnStore that the Python front-end needs (`_lookup`, `_ensure`,
    `_call`), routing the statement calls into the CPU harness."""

    def __init__(self, n, seed=0):
        self.L = lib()
        self.h = C.c_void_p(self.L.hh_new(n, seed))
        self.n = n
        self.names = {}
        self.widths = []
        self._keep = []

    def _lookup(self, name):
        if name in self.names:
            return self.names[name], self.widths[self.names[name]]
        return -1, 0

    def _ensure(self, name, width):
        if name in self.names:
            assert self.widths[self.names[name]] == width
            return self.names[name]
        cid = self.L.hh_col(self.h, width)
        self.names[name] = cid
        self.widths.append(width)
        return cid

    def colnames(self):
        return list(self.names)

    def _call(self, fn, *args):
        name = {"ws_assign": "hh_assign", "ws_assign_vec": "hh_assign_vec", "ws_sample_normal": "hh_sample_normal",
                "ws_sample_exponential": "hh_sample_exponential", "ws_sample_mvnormal": "hh_sample_mvnormal",
                "ws_observe_normal": "hh_observe_normal", "ws_observe_exponential": "hh_observe_exponential",
                "ws_observe_mvnormal": "hh_observe_mvnormal", "ws_weight_expr": "hh_weight_expr",
                "ws_sample_importance_normal": "hh_importance_normal", "ws_sample_expr": "hh_sample_expr"}[fn]
        rc = getattr(self.L, name)(self.h, *args)
        if rc != 0:
            raise RuntimeError(f"{name} failed ({rc}): {self.L.hh_error(self.h).decode()}")
        if getattr(self, "autoflush", False):   # long programs: the harness has no overflow-driven flush (the runtime does)
            self.flush()

    def setcol(self, name, v):
        v = np.asarray(v, dtype=np.float64)
        planes = v[None, :] if v.ndim == 1 else v.T
        cid = self._ensure(name, planes.shape[0])
        for k in range(planes.shape[0]):
            p = np.ascontiguousarray(planes[k])
            self.L.hh_set_plane(self.h, cid, k, p.ctypes.data_as(C.c_void_p))

    def getcol(self, name):
        cid, w = self._lookup(name)
        out = np.empty((w, self.n))
        for k in range(w):
            row = np.empty(self.n)
            self.L.hh_get_plane(self.h, cid, k, row.ctypes.data_as(C.c_void_p))
            out[k] = row
        return out[0] if w == 1 else out.T.copy()

    def logw(self):
        out = np.empty(self.n)
        self.L.hh_get_logw(self.h, out.ctypes.data_as(C.c_void_p))
        return out

    def set_replay(self, normals=(), uniforms=(), exponentials=(), variates=()):
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (normals, uniforms, exponentials)]
        v = np.ascontiguousarray(variates, dtype=np.float64)
        self.L.hh_set_replay_variates(self.h, v.ctypes.data_as(C.c_void_p), v.size)
        self._keep = arrs
        self.L.hh_set_replay(self.h, arrs[0].ctypes.data_as(C.c_void_p), arrs[0].size,
                             arrs[1].ctypes.data_as(C.c_void_p), arrs[1].size,
                             arrs[2].ctypes.data_as(C.c_void_p), arrs[2].size)

    def score(self, n_entries):
        out = np.empty(self.n)
        rc = self.L.hh_score(self.h, n_entries, out.ctypes.data_as(C.c_void_p))
        assert rc == 0
        return out

    def score_device_order(self, n_entries):
        """the fold as the move / score kernels run it: renumbered registers, decoded ops, 128-entry chunks, run
        super-ops.  Returns (scores, number of runs, register-file rows)."""
        out = np.empty(self.n)
        runs, rows = C.c_int(), C.c_int()
        self.L.hh_score_device_order.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        rc = self.L.hh_score_device_order(self.h, n_entries, out.ctypes.data_as(C.c_void_p), C.byref(runs), C.byref(rows))
        assert rc == 0
        return out, runs.value, rows.value

    def tape_len(self):
        return self.L.hh_tape_len(self.h)

    def flush(self):
        assert self.L.hh_flush(self.h) == 0

    def n_flush(self):
        return self.L.hh_n_flush(self.h)


class HostState:
    """Minimal SMCState stand-in so that core.Assign/Sample/Observe/Weight.apply run unchanged."""

    def __init__(self, n, seed=0):
        self.store = HostStore(n, seed)
        self.resampled = False
