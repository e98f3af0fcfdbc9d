"""ctypes binding of the CPU oracle (oracle/knn_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libknn_oracle.so"

SIM = {"cosine": 0, "msd": 1, "pearson": 2, "pearson_baseline": 3}
KNN_TYPE = {"basic": 0, "centered": 1, "zscore": 2, "baseline": 3}
TIE = {"go": 0, "canonical": 1}


def build(force: bool = False) -> Path:
    src = _HERE / "knn_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < max(
            src.stat().st_mtime, (_HERE / "knn_oracle.h").stat().st_mtime):
        subprocess.run(["make", "-C", str(_HERE), "-B", "libknn_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


class IdRating(C.Structure):
    _fields_ = [("id", C.c_int64), ("rating", C.c_double)]


class Params(C.Structure):
    _fields_ = [("sim", C.c_int), ("knn_type", C.c_int), ("user_based", C.c_int), ("k", C.c_int),
                ("min_k", C.c_int), ("n_jobs", C.c_int), ("tie_policy", C.c_int),
                ("reg", C.c_double), ("lr", C.c_double), ("n_epochs", C.c_int),
                ("shrinkage", C.c_double), ("baseline_als", C.c_int), ("als_epochs", C.c_int),
                ("reg_u", C.c_double), ("reg_i", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(str(_LIB_PATH))
    p64 = C.POINTER(C.c_int64)
    pd = C.POINTER(C.c_double)
    p32 = C.POINTER(C.c_int32)
    L.or_params_default.argtypes = [C.POINTER(Params)]
    L.or_sim_lists.restype = C.c_double
    L.or_sim_lists.argtypes = [C.c_int, C.POINTER(IdRating), C.c_int64, C.POINTER(IdRating), C.c_int64]
    L.or_sort_by_id.argtypes = [C.POINTER(IdRating), C.c_int64]
    L.or_trainset_new.restype = C.c_void_p
    L.or_trainset_new.argtypes = [p64, p64, pd, C.c_int64]
    L.or_trainset_free.argtypes = [C.c_void_p]
    for f in ("or_trainset_user_count", "or_trainset_item_count"):
        getattr(L, f).restype = C.c_int64
        getattr(L, f).argtypes = [C.c_void_p]
    L.or_trainset_global_mean.restype = C.c_double
    L.or_trainset_global_mean.argtypes = [C.c_void_p]
    for f in ("or_trainset_convert_user", "or_trainset_convert_item"):
        getattr(L, f).restype = C.c_int64
        getattr(L, f).argtypes = [C.c_void_p, C.c_int64]
    for f in ("or_trainset_inner_users", "or_trainset_inner_items"):
        getattr(L, f).restype = p32
        getattr(L, f).argtypes = [C.c_void_p]
    L.or_knn_new.restype = C.c_void_p
    L.or_knn_new.argtypes = [C.POINTER(Params)]
    L.or_knn_free.argtypes = [C.c_void_p]
    L.or_knn_fit.argtypes = [C.c_void_p, C.c_void_p]
    L.or_knn_fit_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
    L.or_knn_predict.restype = C.c_double
    L.or_knn_predict.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
    L.or_knn_predict_batch.argtypes = [C.c_void_p, p64, p64, C.c_int64, pd, C.c_int]
    L.or_knn_predict_neighbors.restype = C.c_int
    L.or_knn_predict_neighbors.argtypes = [C.c_void_p, C.c_int64, C.c_int64, p64, pd, C.c_int]
    L.or_knn_n.restype = C.c_int64
    L.or_knn_n.argtypes = [C.c_void_p]
    L.or_knn_sims_row.restype = pd
    L.or_knn_sims_row.argtypes = [C.c_void_p, C.c_int64]
    for f in ("or_knn_means", "or_knn_stddevs", "or_knn_bias"):
        getattr(L, f).restype = pd
        getattr(L, f).argtypes = [C.c_void_p]
    L.or_knn_global_mean.restype = C.c_double
    L.or_knn_global_mean.argtypes = [C.c_void_p]
    L.or_knn_topk.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, p32, pd]
    L.or_knn_pair_sums.argtypes = [C.c_void_p, C.c_int64, C.c_int64, p64]
    L.or_baseline_fit.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, pd, pd, pd]
    L.or_slope_fit.restype = C.c_void_p
    L.or_slope_fit.argtypes = [C.c_void_p, C.c_int]
    L.or_slope_fit_rows.restype = C.c_void_p
    L.or_slope_fit_rows.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64]
    L.or_slope_free.argtypes = [C.c_void_p]
    L.or_slope_n.restype = C.c_int64
    L.or_slope_n.argtypes = [C.c_void_p]
    L.or_slope_dev.restype = C.POINTER(C.c_double)
    L.or_slope_dev.argtypes = [C.c_void_p]
    L.or_slope_user_means.restype = C.POINTER(C.c_double)
    L.or_slope_user_means.argtypes = [C.c_void_p]
    L.or_slope_predict.restype = C.c_double
    L.or_slope_predict.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
    L.or_slope_predict_batch.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int64, pd]
    L.or_rows_sims.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_int64, pd]
    L.or_rows_sims.restype = None
    L.or_baseline_als.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, pd, pd]
    L.or_baseline_als.restype = None
    L.or_rmse.restype = C.c_double
    L.or_rmse.argtypes = [pd, pd, C.c_int64]
    L.or_mae.restype = C.c_double
    L.or_mae.argtypes = [pd, pd, C.c_int64]
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def sim_lists(sim: str, a, b, sort: bool = True) -> float:
    """a, b: lists of (id, rating).  sort=True mirrors NewSortedIdRatings (core/data.go:249)."""
    L = lib()
    A = (IdRating * max(1, len(a)))(*[IdRating(i, r) for i, r in a])
    B = (IdRating * max(1, len(b)))(*[IdRating(i, r) for i, r in b])
    if sort:
        L.or_sort_by_id(A, len(a))
        L.or_sort_by_id(B, len(b))
    return L.or_sim_lists(SIM[sim], A, len(a), B, len(b))


class TrainSet:
    def __init__(self, users, items, ratings):
        L = lib()
        self.users = np.ascontiguousarray(users, dtype=np.int64)
        self.items = np.ascontiguousarray(items, dtype=np.int64)
        self.ratings = np.ascontiguousarray(ratings, dtype=np.float64)
        self.h = L.or_trainset_new(_p(self.users, C.c_int64), _p(self.items, C.c_int64),
                                   _p(self.ratings, C.c_double), len(self.ratings))

    def __del__(self):
        if getattr(self, "h", None):
            lib().or_trainset_free(self.h)
            self.h = None

    @property
    def user_count(self):
        return lib().or_trainset_user_count(self.h)

    @property
    def item_count(self):
        return lib().or_trainset_item_count(self.h)

    @property
    def global_mean(self):
        return lib().or_trainset_global_mean(self.h)

    def inner_users(self):
        return np.ctypeslib.as_array(lib().or_trainset_inner_users(self.h), shape=(len(self.ratings),)).copy()

    def inner_items(self):
        return np.ctypeslib.as_array(lib().or_trainset_inner_items(self.h), shape=(len(self.ratings),)).copy()

    def baseline(self, reg=0.02, lr=0.005, n_epochs=20):
        ub = np.zeros(self.user_count, dtype=np.float64)
        ib = np.zeros(self.item_count, dtype=np.float64)
        gb = C.c_double(0.0)
        lib().or_baseline_fit(self.h, reg, lr, n_epochs, _p(ub, C.c_double), _p(ib, C.c_double), C.byref(gb))
        return ub, ib, gb.value


def _baseline_als(self, reg_u=15.0, reg_i=10.0, n_epochs=10):
    """EXTENSION (parity unpinned): ALS baselines, the checker of rs_baseline_als."""
    ub = np.zeros(self.user_count, dtype=np.float64)
    ib = np.zeros(self.item_count, dtype=np.float64)
    lib().or_baseline_als(self.h, reg_u, reg_i, n_epochs, _p(ub, C.c_double), _p(ib, C.c_double))
    return ub, ib, self.global_mean


TrainSet.baseline_als = _baseline_als


class KNN:
    def __init__(self, sim="msd", knn_type="basic", user_based=True, k=40, min_k=1, n_jobs=1,
                 tie_policy="canonical", reg=0.02, lr=0.005, n_epochs=20, shrinkage=0.0, baseline="sgd",
                 reg_u=15.0, reg_i=10.0, als_epochs=10):
        p = Params()
        lib().or_params_default(C.byref(p))
        p.sim, p.knn_type, p.user_based = SIM[sim], KNN_TYPE[knn_type], int(user_based)
        p.k, p.min_k, p.n_jobs, p.tie_policy = k, min_k, n_jobs, TIE[tie_policy]
        p.reg, p.lr, p.n_epochs, p.shrinkage = reg, lr, n_epochs, shrinkage
        p.baseline_als, p.als_epochs, p.reg_u, p.reg_i = int(baseline == "als"), als_epochs, reg_u, reg_i
        self.params = p
        self.h = lib().or_knn_new(C.byref(p))
        self.train = None

    def __del__(self):
        if getattr(self, "h", None):
            lib().or_knn_free(self.h)
            self.h = None

    def fit(self, train: TrainSet, rows=None):
        self.train = train  # keep alive
        if rows is None:
            lib().or_knn_fit(self.h, train.h)
        else:
            lib().or_knn_fit_rows(self.h, train.h, rows[0], rows[1])
        return self

    @property
    def n(self):
        return lib().or_knn_n(self.h)

    def sims(self, copy=True):
        """The N x N matrix.  copy=False returns a view that dies with this object."""
        n = self.n
        view = np.ctypeslib.as_array(lib().or_knn_sims_row(self.h, 0), shape=(n, n))
        return view.copy() if copy else view

    def _vec(self, fn):
        ptr = fn(self.h)
        return None if not ptr else np.ctypeslib.as_array(ptr, shape=(self.n,)).copy()

    def means(self):
        return self._vec(lib().or_knn_means)

    def stddevs(self):
        return self._vec(lib().or_knn_stddevs)

    def bias(self):
        return self._vec(lib().or_knn_bias)

    def global_mean(self):
        return lib().or_knn_global_mean(self.h)

    def predict(self, u, i):
        return lib().or_knn_predict(self.h, int(u), int(i))

    def predict_batch(self, users, items, n_threads=1):
        users = np.ascontiguousarray(users, dtype=np.int64)
        items = np.ascontiguousarray(items, dtype=np.int64)
        out = np.empty(len(users), dtype=np.float64)
        lib().or_knn_predict_batch(self.h, _p(users, C.c_int64), _p(items, C.c_int64), len(users),
                                   _p(out, C.c_double), n_threads)
        return out

    def predict_neighbors(self, u, i, cap=1024):
        ids = np.empty(cap, dtype=np.int64)
        sims = np.empty(cap, dtype=np.float64)
        c = lib().or_knn_predict_neighbors(self.h, int(u), int(i), _p(ids, C.c_int64), _p(sims, C.c_double), cap)
        return ids[:c].copy(), sims[:c].copy()

    def topk(self, kk, row0=0, row1=None):
        row1 = self.n if row1 is None else row1
        idx = np.empty((row1 - row0, kk), dtype=np.int32)
        sim = np.empty((row1 - row0, kk), dtype=np.float64)
        lib().or_knn_topk(self.h, kk, row0, row1, _p(idx, C.c_int32), _p(sim, C.c_double))
        return idx, sim

    def pair_sums(self, a, b):
        out = np.zeros(6, dtype=np.int64)
        lib().or_knn_pair_sums(self.h, int(a), int(b), _p(out, C.c_int64))
        return out


def rmse(pred, truth):
    pred = np.ascontiguousarray(pred, dtype=np.float64)
    truth = np.ascontiguousarray(truth, dtype=np.float64)
    return lib().or_rmse(_p(pred, C.c_double), _p(truth, C.c_double), len(pred))


def mae(pred, truth):
    pred = np.ascontiguousarray(pred, dtype=np.float64)
    truth = np.ascontiguousarray(truth, dtype=np.float64)
    return lib().or_mae(_p(pred, C.c_double), _p(truth, C.c_double), len(pred))


class SlopeOne:
    """core/slope_one.go restated."""

    def __init__(self):
        self.h = None
        self.train = None

    def fit(self, train: TrainSet, n_jobs=8, rows=None):
        self.train = train
        self.h = (lib().or_slope_fit(train.h, n_jobs) if rows is None
                  else lib().or_slope_fit_rows(train.h, n_jobs, rows[0], rows[1]))
        return self

    def __del__(self):
        if getattr(self, "h", None):
            lib().or_slope_free(self.h)
            self.h = None

    def dev(self):
        n = lib().or_slope_n(self.h)
        return np.ctypeslib.as_array(lib().or_slope_dev(self.h), shape=(n, n)).copy()

    def user_means(self):
        return np.ctypeslib.as_array(lib().or_slope_user_means(self.h), shape=(self.train.user_count,)).copy()

    def predict(self, u, i):
        return lib().or_slope_predict(self.h, int(u), int(i))

    def predict_batch(self, users, items):
        users = np.ascontiguousarray(users, dtype=np.int64)
        items = np.ascontiguousarray(items, dtype=np.int64)
        out = np.empty(len(users), dtype=np.float64)
        lib().or_slope_predict_batch(self.h, _p(users, C.c_int64), _p(items, C.c_int64), len(users),
                                     _p(out, C.c_double))
        return out


def rows_sims(train: TrainSet, sim: str, user_based: bool, rows):
    """Similarities of the given left rows against all N (n_rows x N), for shapes too large for the
    N x N matrix."""
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    n = train.user_count if user_based else train.item_count
    out = np.empty((len(rows), n), dtype=np.float64)
    lib().or_rows_sims(train.h, SIM[sim], int(user_based), _p(rows, C.c_int64), len(rows), _p(out, C.c_double))
    return out

