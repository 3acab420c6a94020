"""ctypes binding of ``include/sushie_b200.h`` (the C ABI of the CUDA engine).

The shared library is built in-tree by ``__graft_entry__.build()`` /
``python -m sushie_b200.build`` into ``sushie_b200/_lib/libsushie_b200.so``.  There is no
CPU fallback: if the library or a CUDA device is missing, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import List, Optional, Sequence

import numpy as np

from .utils import RowView

SB_MAX_ANCESTRIES = 6
SB_F64, SB_F32, SB_I8, SB_B2 = 0, 1, 2, 3
SB_CENTER, SB_SCALE = 1, 2
SB_FLAG_SINGLE_PHASE = 1
I8_MAX_SNPS = 8192  # widest locus the hard-call (one byte per entry) traversal handles
B2_MAX_SNPS, B2_MAX_ROWS, B2_MAX_COVAR_BASIS = 8192, 1024, 8  # limits of the bit-plane (0 / 1 / 2 hard calls) traversal

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", "libsushie_b200.so")
# development: A/B a differently built engine without touching the tree
_LIB_PATH = os.environ.get("SUSHIE_B200_LIB", _LIB_PATH)

EXPORTED_SYMBOLS = (
    "sb_version",
    "sb_device_count",
    "sb_init",
    "sb_destroy",
    "sb_last_error",
    "sb_set_stream",
    "sb_ibss_ind",
    "sb_ibss_ss",
    "sb_batch_create_ind",
    "sb_batch_create_ss",
    "sb_batch_run",
    "sb_batch_fetch",
    "sb_batch_destroy",
    "sb_batch_storage",
    "sb_batch_purity",
    "sb_batch_purity_many",
    "sb_batch_select",
)

_pd = C.POINTER(C.c_double)


class SbIndProblem(C.Structure):
    _fields_ = [
        ("k", C.c_int32),
        ("p", C.c_int32),
        ("L", C.c_int32),
        ("n", C.POINTER(C.c_int32)),
        ("X", C.POINTER(C.c_void_p)),
        ("y", C.POINTER(_pd)),
        ("x_dtype", C.c_int32),
        ("standardize", C.c_int32),
        ("pi", _pd),
        ("resid_var", _pd),
        ("effect_covar", _pd),
        ("adj_times", _pd),
        ("adj_plus", _pd),
        ("em_update", C.c_int32),
        ("covar_q", C.POINTER(_pd)),
        ("covar_q_cols", C.POINTER(C.c_int32)),
        ("x_rows", C.POINTER(C.POINTER(C.c_int32))),
        ("x_src_rows", C.POINTER(C.c_int32)),
    ]


class SbSsProblem(C.Structure):
    _fields_ = [
        ("k", C.c_int32),
        ("p", C.c_int32),
        ("L", C.c_int32),
        ("ns", _pd),
        ("ld", C.POINTER(C.c_void_p)),
        ("z", C.POINTER(_pd)),
        ("ld_dtype", C.c_int32),
        ("pi", _pd),
        ("resid_var", _pd),
        ("effect_covar", _pd),
        ("adj_times", _pd),
        ("adj_plus", _pd),
        ("em_update", C.c_int32),
    ]


class SbOpts(C.Structure):
    _fields_ = [
        ("max_iter", C.c_int32),
        ("min_tol", C.c_double),
        ("storage", C.c_int32),
        ("team_size", C.c_int32),
        ("flags", C.c_int32),
        ("reserve_sms", C.c_int32),
        ("reserved", C.c_int32 * 2),
    ]


class SbResult(C.Structure):
    _fields_ = [
        ("alpha", _pd),
        ("post_mean", _pd),
        ("post_mean_sq", _pd),
        ("weighted_sum_covar", _pd),
        ("kl", _pd),
        ("log_bf", _pd),
        ("effect_covar", _pd),
        ("resid_var", _pd),
        ("elbo", _pd),
        ("n_iter", C.c_int32),
        ("elbo_increase", C.c_int32),
        ("status", C.c_int32),
        ("reserved", C.c_int32),
        ("effect_order", C.POINTER(C.c_int32)),
    ]


class SbSelection(C.Structure):
    _fields_ = [
        ("order", C.POINTER(C.c_int32)),
        ("csum", _pd),
        ("sizes", C.POINTER(C.c_int32)),
    ]


class SbRunStats(C.Structure):
    _fields_ = [
        ("device_ms", C.c_double),
        ("prologue_ms", C.c_double),
        ("n_ser", C.c_int64),
        ("n_iter", C.c_int64),
        ("algorithmic_bytes", C.c_double),
        ("n_launches", C.c_int32),
        ("team_size", C.c_int32),
        ("n_teams", C.c_int32),
        ("rows_per_band", C.c_int32),
        ("n_stages", C.c_int32),
        ("smem_bytes", C.c_int32),
        ("n_run_launches", C.c_int32),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


_lib = None
_lib_lock = threading.Lock()


def library_path() -> str:
    return _LIB_PATH


def load_library():
    """Load ``libsushie_b200.so`` and declare prototypes.  Raises if it is not built."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(
                f"sushie_b200: CUDA engine not built ({_LIB_PATH} missing). "
                "Run `python -m sushie_b200.build` (needs nvcc); there is no CPU fallback."
            )
        lib = C.CDLL(_LIB_PATH)
        vp = C.c_void_p
        lib.sb_version.restype = C.c_int
        lib.sb_device_count.restype = C.c_int
        lib.sb_init.argtypes = [C.c_int, C.POINTER(vp)]
        lib.sb_destroy.argtypes = [vp]
        lib.sb_last_error.argtypes = [vp]
        lib.sb_last_error.restype = C.c_char_p
        lib.sb_set_stream.argtypes = [vp, vp]
        lib.sb_ibss_ind.argtypes = [vp, C.POINTER(SbIndProblem), C.c_int, C.POINTER(SbOpts), C.POINTER(SbResult), C.POINTER(SbRunStats)]
        lib.sb_ibss_ss.argtypes = [vp, C.POINTER(SbSsProblem), C.c_int, C.POINTER(SbOpts), C.POINTER(SbResult), C.POINTER(SbRunStats)]
        lib.sb_batch_create_ind.argtypes = [vp, C.POINTER(SbIndProblem), C.c_int, C.POINTER(SbOpts), C.POINTER(vp)]
        lib.sb_batch_create_ss.argtypes = [vp, C.POINTER(SbSsProblem), C.c_int, C.POINTER(SbOpts), C.POINTER(vp)]
        lib.sb_batch_run.argtypes = [vp, vp, C.POINTER(SbRunStats)]
        lib.sb_batch_fetch.argtypes = [vp, vp, C.POINTER(SbResult)]
        lib.sb_batch_destroy.argtypes = [vp, vp]
        lib.sb_batch_storage.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(C.c_int64)]
        lib.sb_batch_purity.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int, _pd]
        lib.sb_batch_purity_many.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                             C.POINTER(C.c_int32), _pd]
        lib.sb_batch_select.argtypes = [vp, vp, _pd, C.POINTER(SbSelection)]
        for name in EXPORTED_SYMBOLS:
            fn = getattr(lib, name)
            if name not in ("sb_last_error",):
                fn.restype = C.c_int
        _lib = lib
        return lib


def _dptr(a: Optional[np.ndarray]):
    return a.ctypes.data_as(_pd) if a is not None else None


def _f64(a, shape=None) -> np.ndarray:
    out = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        out = out.reshape(shape)
    return out


def _matrix_and_dtype(a, allow_i8: bool):
    """Return a C-contiguous array the engine can read directly and its SB_* dtype code."""
    a = np.asarray(a)
    if a.dtype == np.float32:
        return np.ascontiguousarray(a), SB_F32
    if allow_i8 and a.dtype == np.int8:
        return np.ascontiguousarray(a), SB_I8
    return np.ascontiguousarray(a, dtype=np.float64), SB_F64


class EngineError(RuntimeError):
    pass


class IndProblem:
    """Host-side description of one individual-level problem (keeps the arrays alive)."""

    def __init__(self, Xs, ys, L, pi, resid_var, effect_covar, adj_times, adj_plus, em_update, standardize=0, covar_q=None):
        self.k = len(Xs)
        # orthonormal bases of [1, covariates] per ancestry: genotypes are residualised on device (SB_B2 storage)
        self.covar_q = None if covar_q is None else [_f64(q) for q in covar_q]
        if self.covar_q is not None:
            self._qptr = (_pd * self.k)(*[_dptr(q) for q in self.covar_q])
            self._qcols = np.asarray([q.shape[1] for q in self.covar_q], dtype=np.int32)
        # row views (utils.RowView: cross-validation folds): the base matrix is what the engine reads, the row list
        # says which of its rows this problem is made of (gathered on the device)
        views = [x if isinstance(x, RowView) else None for x in Xs]
        mats = [_matrix_and_dtype(x.base if v is not None else x, allow_i8=True) for x, v in zip(Xs, views)]
        dts = {d for _, d in mats}
        if len(dts) > 1:  # mixed input types: promote everything to fp64
            mats = [(np.ascontiguousarray(m, dtype=np.float64), SB_F64) for m, _ in mats]
        self.Xs = [m for m, _ in mats]
        self.x_dtype = mats[0][1]
        self.ys = [_f64(y).reshape(-1) for y in ys]
        self.n = np.asarray([m.shape[0] if v is None else len(v) for m, v in zip(self.Xs, views)], dtype=np.int32)
        self.rows = None
        if any(v is not None for v in views):
            self.rows = [None if v is None else v.rows for v in views]
            i32p = C.POINTER(C.c_int32)
            self._rowsptr = (i32p * self.k)(*[None if r is None else r.ctypes.data_as(i32p) for r in self.rows])
            self._src_rows = np.asarray([m.shape[0] for m in self.Xs], dtype=np.int32)
        self.p = int(self.Xs[0].shape[1])
        self.L = int(L)
        self.pi = None if pi is None else _f64(pi).reshape(-1)
        self.resid_var = _f64(resid_var).reshape(-1)
        self.effect_covar = _f64(effect_covar, (self.L, self.k, self.k))
        self.adj_times = _f64(adj_times, (self.k, self.k))
        self.adj_plus = _f64(adj_plus, (self.k, self.k))
        self.em_update = int(bool(em_update))
        self.standardize = int(standardize)
        self._xptr = (C.c_void_p * self.k)(*[x.ctypes.data for x in self.Xs])
        self._yptr = (_pd * self.k)(*[_dptr(y) for y in self.ys])

    @property
    def cost(self) -> float:
        return float(np.sum(self.n)) * self.p * self.L

    def fill(self, s: SbIndProblem):
        s.k, s.p, s.L = self.k, self.p, self.L
        s.n = self.n.ctypes.data_as(C.POINTER(C.c_int32))
        s.X = C.cast(self._xptr, C.POINTER(C.c_void_p))
        s.y = C.cast(self._yptr, C.POINTER(_pd))
        s.x_dtype = self.x_dtype
        s.standardize = self.standardize
        s.pi = _dptr(self.pi)
        s.resid_var = _dptr(self.resid_var)
        s.effect_covar = _dptr(self.effect_covar)
        s.adj_times = _dptr(self.adj_times)
        s.adj_plus = _dptr(self.adj_plus)
        s.em_update = self.em_update
        if self.covar_q is not None:
            s.covar_q = C.cast(self._qptr, C.POINTER(_pd))
            s.covar_q_cols = self._qcols.ctypes.data_as(C.POINTER(C.c_int32))
        else:
            s.covar_q = None
            s.covar_q_cols = None
        if self.rows is not None:
            s.x_rows = C.cast(self._rowsptr, C.POINTER(C.POINTER(C.c_int32)))
            s.x_src_rows = self._src_rows.ctypes.data_as(C.POINTER(C.c_int32))
        else:
            s.x_rows = None
            s.x_src_rows = None


class SsProblem:
    """Host-side description of one summary-level problem."""

    def __init__(self, lds, ns, zs, L, pi, resid_var, effect_covar, adj_times, adj_plus, em_update):
        self.k = len(lds)
        mats = [_matrix_and_dtype(x, allow_i8=False) for x in lds]
        dts = {d for _, d in mats}
        if len(dts) > 1:
            mats = [(np.ascontiguousarray(m, dtype=np.float64), SB_F64) for m, _ in mats]
        self.lds = [m for m, _ in mats]
        self.ld_dtype = mats[0][1]
        self.zs = [_f64(z).reshape(-1) for z in zs]
        self.ns = _f64(ns).reshape(-1)
        self.p = int(self.lds[0].shape[0])
        self.n = np.asarray([self.p] * self.k, dtype=np.int32)
        self.L = int(L)
        self.pi = None if pi is None else _f64(pi).reshape(-1)
        self.resid_var = _f64(resid_var).reshape(-1)
        self.effect_covar = _f64(effect_covar, (self.L, self.k, self.k))
        self.adj_times = _f64(adj_times, (self.k, self.k))
        self.adj_plus = _f64(adj_plus, (self.k, self.k))
        self.em_update = int(bool(em_update))
        self._ldptr = (C.c_void_p * self.k)(*[x.ctypes.data for x in self.lds])
        self._zptr = (_pd * self.k)(*[_dptr(z) for z in self.zs])

    @property
    def cost(self) -> float:
        return float(self.k) * self.p * self.p * self.L

    def fill(self, s: SbSsProblem):
        s.k, s.p, s.L = self.k, self.p, self.L
        s.ns = _dptr(self.ns)
        s.ld = C.cast(self._ldptr, C.POINTER(C.c_void_p))
        s.z = C.cast(self._zptr, C.POINTER(_pd))
        s.ld_dtype = self.ld_dtype
        s.pi = _dptr(self.pi)
        s.resid_var = _dptr(self.resid_var)
        s.effect_covar = _dptr(self.effect_covar)
        s.adj_times = _dptr(self.adj_times)
        s.adj_plus = _dptr(self.adj_plus)
        s.em_update = self.em_update


class RawFit:
    """Numpy view of one ``sb_result``."""

    FIELDS = ("alpha", "post_mean", "post_mean_sq", "weighted_sum_covar", "kl", "log_bf", "effect_covar", "resid_var")

    def __init__(self, k, p, L, max_iter, small_only: bool = False):
        shapes = {"alpha": (L, p), "post_mean": (L, p, k), "post_mean_sq": (L, p, k, k), "weighted_sum_covar": (L, k, k),
                  "kl": (L,), "log_bf": (L, p), "effect_covar": (L, k, k), "resid_var": (k,)}
        if small_only:  # what the effect order and the convergence log need
            shapes = {name: shapes[name] for name in ("weighted_sum_covar",)}
        sizes = {name: int(np.prod(shape)) for name, shape in shapes.items()}
        buf = np.empty(sum(sizes.values()) + max_iter + 1)  # one allocation per fit, the fields are views
        pos = 0
        for name in self.FIELDS:
            if name in shapes:
                setattr(self, name, buf[pos: pos + sizes[name]].reshape(shapes[name]))
                pos += sizes[name]
            else:
                setattr(self, name, None)
        self._elbo = buf[pos:]
        self._elbo.fill(np.nan)
        self.n_iter = 0
        self.elbo_increase = True
        self.status = 0  # SB_STATUS_*: 1 = the run stopped on a NaN ELBO
        self.l_order = None  # effect order applied by the download (None: engine order)

    def fill(self, s: SbResult, effect_order: Optional[np.ndarray] = None):
        for name in self.FIELDS:
            setattr(s, name, _dptr(getattr(self, name)))
        s.elbo = _dptr(self._elbo)
        if effect_order is not None:
            self.l_order = np.ascontiguousarray(effect_order, dtype=np.int32)
            s.effect_order = self.l_order.ctypes.data_as(C.POINTER(C.c_int32))

    def finish(self, s: SbResult):
        self.n_iter = int(s.n_iter)
        self.elbo_increase = bool(s.elbo_increase)
        self.status = int(s.status)
        self.elbo = self._elbo[: self.n_iter + 1].copy()


class Engine:
    """One ``sb_ctx``: one host thread driving one GPU."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self._ctx = C.c_void_p()
        rc = self.lib.sb_init(int(device), C.byref(self._ctx))
        if rc != 0:
            msg = self.lib.sb_last_error(None)
            raise EngineError(f"sb_init(device={device}) failed ({rc}): {msg.decode() if msg else ''}")
        self.device = int(device)

    def close(self):
        if self._ctx:
            self.lib.sb_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self.lib.sb_last_error(self._ctx)
            raise EngineError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def set_stream(self, cuda_stream_handle: int):
        self._check(self.lib.sb_set_stream(self._ctx, C.c_void_p(cuda_stream_handle)), "sb_set_stream")

    @staticmethod
    def make_opts(max_iter=500, min_tol=1e-4, storage=SB_F64, team_size=0, single_phase=False, reserve_sms=0) -> SbOpts:
        o = SbOpts()
        o.max_iter, o.min_tol, o.storage, o.team_size = int(max_iter), float(min_tol), int(storage), int(team_size)
        o.flags = SB_FLAG_SINGLE_PHASE if single_phase else 0
        o.reserve_sms = int(reserve_sms)
        return o

    # ---- resident batches -------------------------------------------------------------
    def create_batch(self, problems: Sequence, opts: SbOpts) -> "Batch":
        return Batch(self, problems, opts)

    # ---- one-shot ---------------------------------------------------------------------
    def run(self, problems: Sequence, opts: SbOpts):
        """Upload, run to convergence, download.  Returns (list of RawFit, stats dict)."""
        n = len(problems)
        ss = isinstance(problems[0], SsProblem)
        arr = ((SbSsProblem if ss else SbIndProblem) * n)()
        res = (SbResult * n)()
        fits: List[RawFit] = []
        for i, pr in enumerate(problems):
            pr.fill(arr[i])
            f = RawFit(pr.k, pr.p, pr.L, opts.max_iter)
            f.fill(res[i])
            fits.append(f)
        stats = SbRunStats()
        fn = self.lib.sb_ibss_ss if ss else self.lib.sb_ibss_ind
        self._check(fn(self._ctx, arr, n, C.byref(opts), res, C.byref(stats)), "sb_ibss_ss" if ss else "sb_ibss_ind")
        for f, r in zip(fits, res):
            f.finish(r)
        return fits, stats.as_dict()


class Batch:
    """Device-resident batch (``sb_batch``)."""

    def __init__(self, engine: Engine, problems: Sequence, opts: SbOpts):
        self.engine = engine
        self.problems = list(problems)
        self.opts = opts
        n = len(self.problems)
        self.ss = isinstance(self.problems[0], SsProblem)
        arr = ((SbSsProblem if self.ss else SbIndProblem) * n)()
        for i, pr in enumerate(self.problems):
            pr.fill(arr[i])
        self._h = C.c_void_p()
        fn = engine.lib.sb_batch_create_ss if self.ss else engine.lib.sb_batch_create_ind
        engine._check(fn(engine._ctx, arr, n, C.byref(opts), C.byref(self._h)), "sb_batch_create")

    def run(self) -> dict:
        stats = SbRunStats()
        self.engine._check(self.engine.lib.sb_batch_run(self.engine._ctx, self._h, C.byref(stats)), "sb_batch_run")
        return stats.as_dict()

    def fetch(self, effect_orders: Optional[Sequence] = None, small_only: bool = False) -> List[RawFit]:
        """Download the fits.  ``effect_orders[i]`` (a permutation of 0..L-1 or None) re-orders the effects of
        problem i while the download is copied out (``_reorder_l``); ``small_only`` fetches just the ELBO trace,
        the iteration counts and ``weighted_sum_covar`` -- what decides the effect order."""
        n = len(self.problems)
        res = (SbResult * n)()
        fits = []
        for i, pr in enumerate(self.problems):
            f = RawFit(pr.k, pr.p, pr.L, self.opts.max_iter, small_only)
            f.fill(res[i], None if effect_orders is None else effect_orders[i])
            fits.append(f)
        self.engine._check(self.engine.lib.sb_batch_fetch(self.engine._ctx, self._h, res), "sb_batch_fetch")
        for f, r in zip(fits, res):
            f.finish(r)
        return fits

    def select(self, thresholds: Sequence[float]):
        """Credible-set selection on the device (``sb_batch_select``): per problem the SNP order by descending
        alpha ``[L, p]`` int32, the running sum in that order ``[L, p]`` and the set sizes ``[L]`` -- effects in
        ENGINE order."""
        n = len(self.problems)
        thr = np.ascontiguousarray(thresholds, dtype=np.float64)
        if thr.shape != (n,):
            raise ValueError("one threshold per problem")
        sel = (SbSelection * n)()
        out = []
        i32p = C.POINTER(C.c_int32)
        for i, pr in enumerate(self.problems):
            order = np.empty((pr.L, pr.p), dtype=np.int32)
            csum = np.empty((pr.L, pr.p))
            sizes = np.empty(pr.L, dtype=np.int32)
            sel[i].order, sel[i].csum, sel[i].sizes = order.ctypes.data_as(i32p), _dptr(csum), sizes.ctypes.data_as(i32p)
            out.append((order, csum, sizes))
        self.engine._check(self.engine.lib.sb_batch_select(self.engine._ctx, self._h, _dptr(thr), sel), "sb_batch_select")
        return out

    def purity(self, prob: int, index_sets: Sequence[np.ndarray]) -> np.ndarray:
        """min |corr| per (set, ancestry) on the device-resident genotypes / LD."""
        k = self.problems[prob].k
        if len(index_sets) == 0:
            return np.zeros((0, k))
        offs = np.zeros(len(index_sets) + 1, dtype=np.int32)
        offs[1:] = np.cumsum([len(s) for s in index_sets])
        idx = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.int32) for s in index_sets]), dtype=np.int32)
        out = np.empty((len(index_sets), k))
        self.engine._check(
            self.engine.lib.sb_batch_purity(
                self.engine._ctx, self._h, int(prob),
                idx.ctypes.data_as(C.POINTER(C.c_int32)), offs.ctypes.data_as(C.POINTER(C.c_int32)),
                len(index_sets), _dptr(out),
            ),
            "sb_batch_purity",
        )
        return out

    def purity_many(self, index_sets_per_problem: Sequence[Sequence[np.ndarray]]) -> List[np.ndarray]:
        """``purity`` for the sets of every problem of the batch in one launch: entry i lists the index sets of
        problem i (may be empty); returns one ``[n_sets_i, k_i]`` array per problem."""
        counts = [len(sets) for sets in index_sets_per_problem]
        flat = [np.asarray(s, dtype=np.int32) for sets in index_sets_per_problem for s in sets]
        ks = [pr.k for pr in self.problems]
        if not flat:
            return [np.zeros((0, k)) for k in ks]
        i32 = C.POINTER(C.c_int32)
        set_prob = np.repeat(np.arange(len(counts), dtype=np.int32), counts)
        offs = np.zeros(len(flat) + 1, dtype=np.int64)
        offs[1:] = np.cumsum([s.size for s in flat])
        if offs[-1] >= 2**31:
            raise EngineError("too many credible-set members for one purity call")
        offs = offs.astype(np.int32)
        idx = np.ascontiguousarray(np.concatenate(flat), dtype=np.int32)
        out = np.empty((len(flat), SB_MAX_ANCESTRIES))
        self.engine._check(
            self.engine.lib.sb_batch_purity_many(
                self.engine._ctx, self._h, len(flat), set_prob.ctypes.data_as(i32), idx.ctypes.data_as(i32),
                offs.ctypes.data_as(i32), _dptr(out),
            ),
            "sb_batch_purity_many",
        )
        bounds = np.concatenate([[0], np.cumsum(counts)])
        return [out[bounds[i]:bounds[i + 1], :k].copy() for i, k in enumerate(ks)]

    def storage_ptr(self, prob: int, ancestry: int):
        ptr, ld = C.c_void_p(), C.c_int64()
        self.engine._check(
            self.engine.lib.sb_batch_storage(self._h, prob, ancestry, C.byref(ptr), C.byref(ld)), "sb_batch_storage"
        )
        return ptr.value, ld.value

    def close(self):
        if self._h:
            self.engine.lib.sb_batch_destroy(self.engine._ctx, self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_engines = {}
_idle_engines = {}  # device -> engines released by finished worker threads, ready to be handed to the next one
_MAX_IDLE_ENGINES = 4
_engines_lock = threading.Lock()


def get_engine(device: int = 0) -> Engine:
    """Per-(thread, device) engine (one ``sb_ctx`` per host thread and GPU).  A thread that has none takes an idle
    one before creating it: an engine owns a stream, events and ~260 MB of page-locked staging, and pinning that much
    memory costs tenths of a second -- seconds when the ranks of a box all start their workers at once."""
    key = (threading.get_ident(), int(device))
    with _engines_lock:
        eng = _engines.get(key)
        if eng is None:
            idle = _idle_engines.get(int(device))
            eng = idle.pop() if idle else None
            if eng is not None:
                _engines[key] = eng
    if eng is None:
        eng = Engine(device)  # (outside the lock: sb_init takes a while)
        with _engines_lock:
            _engines[key] = eng
    return eng


_run_locks = {}
# per-thread hook called when the calling thread's chunk enters the IBSS kernel (the distributed work queue uses it
# to let the next worker of the rank pull a chunk: see batch.run_distributed)
run_hooks = threading.local()


def device_run_lock(device: int) -> threading.Lock:
    """One lock per GPU around the IBSS launch sequence of a chunk.  Worker threads that share a GPU take turns in
    the long kernel: two persistent kernels in flight at once end up finishing together, after which both workers
    do their host work at the same time and the GPU idles (measured: lock step, 22 instead of 30 genes/s).  With the
    lock the next chunk is uploaded and waiting when the current one ends, and a worker's host half always runs
    under the other worker's kernel."""
    with _engines_lock:
        lock = _run_locks.get(int(device))
        if lock is None:
            lock = _run_locks[int(device)] = threading.Lock()
        return lock


def release_thread_engines() -> None:
    """Give the engines of the calling thread back.  Short-lived worker threads (``batch.run_sharded`` starts some per
    call) must not leave their ``sb_ctx`` registered under a dead thread; up to ``_MAX_IDLE_ENGINES`` per GPU are kept
    for the next worker threads, the rest are closed."""
    me = threading.get_ident()
    surplus = []
    with _engines_lock:
        for key in [key for key in _engines if key[0] == me]:
            eng = _engines.pop(key)
            idle = _idle_engines.setdefault(key[1], [])
            (idle if len(idle) < _MAX_IDLE_ENGINES else surplus).append(eng)
    for eng in surplus:
        eng.close()


def close_idle_engines() -> None:
    """Close every pooled engine (tests; interpreter exit)."""
    with _engines_lock:
        engines = [eng for idle in _idle_engines.values() for eng in idle]
        _idle_engines.clear()
    for eng in engines:
        eng.close()


def device_count() -> int:
    return int(load_library().sb_device_count())
