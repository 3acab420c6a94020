"""TEST INFRASTRUCTURE ONLY -- run the reference's own Python source without JAX.

The reference (``/root/reference/sushie``) is plain Python whose array primitives come from
``jax==0.4.33`` / ``equinox==0.10.2`` (``setup.cfg:35-40``); neither is installed or installable
here.  This module registers stand-in ``jax`` / ``equinox`` modules backed by **numpy** (fp64, the
reference's ``jax_enable_x64`` CLI setting) so that ``sushie/infer.py``, ``sushie/infer_ss.py`` and
``sushie/utils.py`` are imported and executed UNMODIFIED from where they lie: every line of the
reference's algorithm text runs; only the third-party primitives underneath are substituted, each by
the numpy / scipy routine with the same published definition:

  jnp.<f>                          -> numpy.<f>              (results viewed as an ndarray subclass with ``.at[i].set(v)``)
  jnp.argsort                      -> numpy.argsort(kind="stable")      (jax sorts are stable)
  jnp.linalg.{inv,slogdet,svd,qr}  -> numpy.linalg (LAPACK LU / SVD / QR, as XLA's CPU backend uses)
  jax.scipy.linalg.solve_triangular / cho_solve -> scipy.linalg
  jax.scipy.stats.multivariate_normal.logpdf    -> Cholesky + triangular solve, jax's algorithm
                                                   (jax/_src/scipy/stats/multivariate_normal.py), NaN when not PD
  jax.nn.softmax                   -> exp(x - max x) / sum exp(x - max x)
  jax.lax.fori_loop                -> Python ``for``
  jax.random.PRNGKey/split/choice  -> the threefry2x32 restatement in ``oracle/threefry.py`` (checked against
                                      the Random123 known answers in tests/test_oracle.py)
  equinox.Module / filter_jit      -> plain base class / identity decorator (jit does not change values)

It is used by ``tests/golden/make_reference_golden.py`` to generate the golden vectors under
``tests/golden/ref_*.npz`` (``/root/reference`` does not exist on the GPU box) and by
``tests/test_reference_shim.py`` when ``/root/reference`` is present.  Nothing in ``sushie_b200`` imports it.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("SUSHIE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "sushie", "infer.py"))


class _At:
    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self._arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self._arr, self._idx = arr, idx

    def set(self, value):
        out = np.array(self._arr, copy=True).view(ShimArray)
        out[self._idx] = value
        return out

    def add(self, value):
        out = np.array(self._arr, copy=True).view(ShimArray)
        out[self._idx] += value
        return out


class ShimArray(np.ndarray):
    """numpy array with jax's functional-update syntax (``x.at[i].set(v)`` returns a copy)."""

    @property
    def at(self):
        return _At(self)


def _wrap(value):
    if isinstance(value, np.ndarray) and not isinstance(value, ShimArray):
        return value.view(ShimArray)
    if isinstance(value, np.generic):
        return np.asarray(value).view(ShimArray)
    if isinstance(value, tuple):
        return tuple(_wrap(v) for v in value)
    return value


def _wrapping(fn):
    def call(*args, **kwargs):
        return _wrap(fn(*args, **kwargs))

    call.__name__ = getattr(fn, "__name__", "fn")
    return call


class _NumpyNamespace(types.ModuleType):
    """``jax.numpy`` / ``jax.numpy.linalg`` look-alike: attribute access falls through to numpy."""

    def __init__(self, name, source, overrides=None):
        super().__init__(name)
        self._source = source
        self._overrides = overrides or {}

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name in self._overrides:
            return self._overrides[name]
        obj = getattr(self._source, name)
        if callable(obj) and not isinstance(obj, type):
            return _wrapping(obj)
        return obj


def _argsort(a, axis=-1, **kw):
    kw.pop("kind", None)
    kw.pop("stable", None)
    return np.argsort(a, axis=axis, kind="stable")


def _mvn_logpdf(x, mean, cov):
    """jax.scipy.stats.multivariate_normal.logpdf: L = cholesky(cov); y = L^-1 (x - mean);
    -1/2 |y|^2 - n/2 log(2 pi) - sum log diag L.  A non-PD covariance gives NaN (jax's cholesky)."""
    x, mean, cov = np.asarray(x, float), np.asarray(mean, float), np.asarray(cov, float)
    n = cov.shape[-1]
    dev = x - mean
    batch = np.broadcast_shapes(dev.shape[:-1], cov.shape[:-2])
    dev = np.broadcast_to(dev, batch + (n,)).reshape(-1, n)
    cov = np.broadcast_to(cov, batch + (n, n)).reshape(-1, n, n)
    out = np.full(dev.shape[0], np.nan)
    ok = np.all(np.isfinite(cov.reshape(cov.shape[0], -1)), axis=1)
    idx = np.nonzero(ok)[0]
    try:
        chol = np.linalg.cholesky(cov[idx])
    except np.linalg.LinAlgError:
        keep, fac = [], []
        for j in idx:
            try:
                fac.append(np.linalg.cholesky(cov[j]))
                keep.append(j)
            except np.linalg.LinAlgError:
                pass
        idx = np.asarray(keep, dtype=int)
        chol = np.asarray(fac).reshape(len(keep), n, n)
    if len(idx):
        import scipy.linalg as sla

        if len(idx) * n * n < 64:
            y = np.stack([sla.solve_triangular(chol[t], dev[idx[t]], lower=True) for t in range(len(idx))])
        else:
            y = np.linalg.solve(chol, dev[idx][..., None])[..., 0]  # triangular system, LAPACK general solve
        out[idx] = (-0.5 * np.einsum("...i,...i->...", y, y) - n / 2.0 * np.log(2 * np.pi)
                    - np.log(np.diagonal(chol, axis1=-2, axis2=-1)).sum(-1))
    return _wrap(out.reshape(batch))


def _softmax(x, axis=-1):
    x = np.asarray(x, float)
    with np.errstate(invalid="ignore", over="ignore"):
        u = np.exp(x - np.max(x, axis=axis, keepdims=True))
        return _wrap(u / np.sum(u, axis=axis, keepdims=True))


def _fori_loop(lower, upper, body, init):
    val = init
    for i in range(int(lower), int(upper)):
        val = body(i, val)
    return val


def _prng_key(seed):
    from oracle import threefry

    return threefry.prng_key(int(seed))


def _choice(key, a, shape=(), replace=True, p=None, axis=0):
    from oracle import threefry

    if replace or p is not None:
        raise NotImplementedError("shim: only choice(..., replace=False) is needed by the reference")
    n_draws = int(np.prod(shape)) if shape != () else 1
    values = np.arange(int(a)) if np.ndim(a) == 0 else np.asarray(a)  # jax: an int means arange(a)
    return _wrap(np.asarray(threefry.choice_without_replacement(key, values, n_draws)).reshape(shape))


def _split(key, num=2):
    from oracle import threefry

    return threefry.split(key, int(num))


def _permutation(key, x):
    from oracle import threefry

    x = np.arange(x) if np.ndim(x) == 0 else np.asarray(x)
    return _wrap(threefry.permutation(key, x))


class _Module:
    """equinox.Module stand-in (the reference's modules hold no fields)."""


def _identity_decorator(fn=None, **_kw):
    if fn is None:
        return lambda f: f
    return fn


_INSTALLED = False


def install() -> None:
    """Register the stand-in modules (idempotent).  Refuses to shadow a real jax."""
    global _INSTALLED
    if _INSTALLED:
        return
    try:
        import jax  # noqa: F401

        if not getattr(jax, "__refshim__", False):
            raise RuntimeError("a real jax is importable: use it (baseline/_ref) instead of the shim")
    except ImportError:
        pass
    import scipy.linalg as sla

    jnp_linalg = _NumpyNamespace("jax.numpy.linalg", np.linalg)
    jnp = _NumpyNamespace("jax.numpy", np, {
        "linalg": jnp_linalg,
        "argsort": _wrapping(_argsort),
        "array": _wrapping(lambda obj, dtype=None, **kw: np.array(obj, dtype=dtype)),
        "asarray": _wrapping(lambda obj, dtype=None, **kw: np.asarray(obj, dtype=dtype)),
        "newaxis": None, "pi": np.pi, "inf": np.inf, "nan": np.nan,
        "ndarray": np.ndarray,
    })
    jax = types.ModuleType("jax")
    jax.__refshim__ = True
    jax.__path__ = []
    jax.numpy = jnp
    jax.Array = np.ndarray
    lax = types.ModuleType("jax.lax")
    lax.fori_loop = _fori_loop
    nn = types.ModuleType("jax.nn")
    nn.softmax = _softmax
    random = types.ModuleType("jax.random")
    random.PRNGKey = _prng_key
    random.choice = _choice
    random.split = _split
    random.permutation = _permutation
    typing_m = types.ModuleType("jax.typing")
    typing_m.ArrayLike = object
    jsp = types.ModuleType("jax.scipy")
    jsp.__path__ = []
    jsp_stats = types.ModuleType("jax.scipy.stats")
    mvn = types.ModuleType("jax.scipy.stats.multivariate_normal")
    mvn.logpdf = _mvn_logpdf
    jsp_stats.multivariate_normal = mvn
    jsp_linalg = types.ModuleType("jax.scipy.linalg")
    jsp_linalg.solve_triangular = _wrapping(sla.solve_triangular)
    jsp_linalg.cho_solve = _wrapping(sla.cho_solve)
    jsp.stats, jsp.linalg = jsp_stats, jsp_linalg

    class _Config:
        def update(self, *a, **k):
            pass

    jax.config = _Config()
    jax.lax, jax.nn, jax.random, jax.typing, jax.scipy = lax, nn, random, typing_m, jsp
    eqx = types.ModuleType("equinox")
    eqx.Module = _Module
    eqx.filter_jit = _identity_decorator
    mods = {
        "jax": jax, "jax.numpy": jnp, "jax.numpy.linalg": jnp_linalg, "jax.lax": lax, "jax.nn": nn,
        "jax.random": random, "jax.typing": typing_m, "jax.scipy": jsp, "jax.scipy.stats": jsp_stats,
        "jax.scipy.stats.multivariate_normal": mvn, "jax.scipy.linalg": jsp_linalg, "equinox": eqx,
    }
    # third-party imports at the top of sushie/utils.py and sushie/io.py that the fine-mapping path never calls
    for name, attrs in {
        "glimix_core": (), "glimix_core.lmm": ("LMM",), "numpy_sugar": (), "numpy_sugar.linalg": ("economic_qs",),
        "pandas_plink": ("read_plink",), "cyvcf2": ("VCF",), "bgen_reader": ("open_bgen",),
    }.items():
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
            continue
        except Exception:
            pass
        m = types.ModuleType(name)
        m.__path__ = []
        for a in attrs:
            setattr(m, a, _unavailable(name, a))
        mods[name] = m
    sys.modules.update(mods)
    _INSTALLED = True


def _unavailable(mod, name):
    def fn(*a, **k):
        raise RuntimeError(f"{mod}.{name} is not available in the reference shim")

    return fn


def load_reference(submodules=("log", "utils", "infer", "infer_ss")):
    """Import ``sushie.<submodule>`` from ``REFERENCE_ROOT`` without running ``sushie/__init__.py`` (which pulls
    in the whole CLI).  Returns a namespace with the loaded sub-modules."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    install()
    pkg = sys.modules.get("sushie")
    if pkg is None or not getattr(pkg, "__refshim__", False):
        pkg = types.ModuleType("sushie")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "sushie")]
        pkg.__refshim__ = True
        sys.modules["sushie"] = pkg
    ns = types.SimpleNamespace()
    for sub in submodules:
        mod = importlib.import_module(f"sushie.{sub}")
        setattr(pkg, sub, mod)
        setattr(ns, sub, mod)
    return ns
