"""Individual-level SuShiE on the GPU: drop-in for ``sushie.infer`` of the reference.

``infer_sushie`` keeps the reference signature, validation, defaults and result layout
(``/root/reference/sushie/infer.py:162-587``).  The host does only what the reference
does outside its jitted function (validation, covariate residualisation, phenotype
centring/scaling, prior set-up); genotype standardisation, the whole IBSS loop with its
stopping rule, and the purity Gram matrices run in the CUDA engine behind the C ABI
(``include/sushie_b200.h``).  There is no CPU fallback.

Deliberate differences from the reference:
  * the caller's ``Xs`` / ``ys`` lists are not mutated;
  * arrays in the result are numpy, not ``jax.Array``;
  * at most ``SB_MAX_ANCESTRIES`` (6) ancestries (one kernel instantiation per count).
"""
from __future__ import annotations

import sys
import time
from typing import List, Optional, Sequence

import numpy as np

from . import _capi, config, cs as _cs, utils
from ._common import (
    Posterior,
    Prior,
    StagedResult,
    SushieResult,
    _PriorAdjustor,
    _reorder_l,
    build_priors,
    check_scalar_options,
    check_snp_counts,
    effect_orders,
    log_convergence,
    resolve_pi,
    resolve_resid_var,
)

__all__ = [
    "Prior",
    "Posterior",
    "SushieResult",
    "infer_sushie",
    "infer_sushie_batch",
    "make_cs",
    "_PriorAdjustor",
    "_reorder_l",
]


class _Prepared:
    """One validated problem: engine arguments + what the epilogue needs."""

    def __init__(self, problem, pi, sample_size, max_iter, min_tol, cs_args, no_reorder, hard_calls=False,
                 lazy_tables=False):
        self.problem = problem
        self.lazy_tables = lazy_tables  # the caller will not read `cs` / `alphas` (built on first access if it does)
        # device storage class of the genotypes: False = dense matrix at `config.precision`; "i8" = integers in
        # 0..127, not residualised (one byte per entry); "b2" = additionally only 0 / 1 / 2 and inside the limits of
        # the bit-plane traversal (look-up tables instead of fp64 multiply-adds on widened genotypes)
        self.hard_calls = hard_calls
        self.pi = pi
        self.sample_size = sample_size
        self.max_iter = max_iter
        self.min_tol = min_tol
        self.cs_args = cs_args
        self.no_reorder = no_reorder


class _HardCallCache:
    """``utils.as_hard_calls`` once per genotype matrix of an ``infer_sushie_batch`` call: a gene's full fit and the
    row views of its cross-validation folds look at the same memory, so the check (a full pass) and the int8 copy
    are shared -- and the engine sees one pointer, which it uploads once (``sb_ind_problem.x_rows``)."""

    def __init__(self):
        self._seen = {}

    def packed(self, base: np.ndarray):
        """(int8 matrix or None, largest entry) of a plain 2-D array."""
        key = utils.array_key(base)
        hit = self._seen.get(key)
        if hit is None:
            packed = utils.as_hard_calls(base)
            hit = (packed, int(packed.max()) if packed is not None and packed.size else 0, base)  # (keeps `base` alive)
            self._seen[key] = hit
        return hit[0], hit[1]

    def __call__(self, m):
        """``m`` (array or ``utils.RowView``) as hard calls -> (int8 array / view or None, largest entry)."""
        if isinstance(m, utils.RowView):
            packed, top = self.packed(m.base)
            return (None if packed is None else utils.RowView(packed, m.rows)), top
        m = np.asarray(m)
        if m.ndim != 2:
            return None, 0
        return self.packed(m)


def _prepare(
    Xs, ys, covar, L, no_scale, no_regress, no_update, pi, resid_var, effect_var, rho,
    max_iter, min_tol, threshold, purity, purity_method, max_select, min_snps, no_reorder, seed, hard_cache=None,
) -> _Prepared:
    if len(Xs) == len(ys):
        n_pop = len(Xs)
    else:
        raise ValueError(
            f"The number of geno ({len(Xs)}) and pheno ({len(ys)}) data does not match. Check your input."
        )
    for idx in range(n_pop):
        if Xs[idx].shape[0] != ys[idx].shape[0]:
            raise ValueError(
                f"Ancestry {idx + 1}: The sample size of geno ({Xs[idx].shape[0]}) "
                + f"and pheno ({ys[idx].shape[0]}) data does not match. Check your input."
            )
    for idx in range(1, n_pop):
        if Xs[idx - 1].shape[1] != Xs[idx].shape[1]:
            raise ValueError(
                f"Ancestry {idx} and ancestry {idx} do not have "
                + f"the same number of SNPs ({Xs[idx - 1].shape[1]} vs {Xs[idx].shape[1]})."
            )
    check_scalar_options(L, min_tol, threshold, purity, max_select, min_snps, summary=False)
    if n_pop > _capi.SB_MAX_ANCESTRIES:
        raise ValueError(
            f"sushie_b200 supports at most {_capi.SB_MAX_ANCESTRIES} ancestries per fit ({n_pop} given)."
        )
    n_snps = Xs[0].shape[1]
    pi = resolve_pi(pi, n_snps, summary=False)

    # covariates: QR residualisation on the host, as the reference does before its jitted loop -- except that hard
    # calls 0 / 1 / 2 stay hard calls: their residualisation (I - Q Q^T) G is applied on the device as a low-rank
    # correction of the bit-plane products (SB_B2 storage), so `--covar` runs do not fall back to the dense matrix
    mats: List[np.ndarray] = list(Xs)
    phen: List[np.ndarray] = [np.asarray(y, dtype=float) for y in ys]
    hard_calls = False
    covar_q = None

    hard_cache = hard_cache or _HardCallCache()

    def bit_plane_ok(packed, tops):
        return (config.bit_planes and n_snps <= _capi.B2_MAX_SNPS and max(m.shape[0] for m in packed) <= _capi.B2_MAX_ROWS
                and all(top <= 2 for top in tops))

    def dense(m):
        return np.asarray(m)  # (materialises a row view)

    auto = config.genotype_storage == "auto"
    if covar is not None and not no_regress and auto:
        packed, tops = zip(*[hard_cache(m) for m in mats])
        cov = [np.asarray(c, dtype=float) for c in covar]
        n_basis = max(1 + (c.reshape(c.shape[0], -1).shape[1]) for c in cov)
        if all(m is not None for m in packed) and bit_plane_ok(packed, tops) and n_basis <= _capi.B2_MAX_COVAR_BASIS:
            covar_q = [utils.covar_basis(c) for c in cov]
            for idx in range(n_pop):
                _, phen[idx] = utils.regress_covar(None, phen[idx], cov[idx], True)  # phenotype only
            mats = list(packed)
            hard_calls = "b2"
    if covar is not None and covar_q is None:
        for idx in range(n_pop):
            mats[idx], phen[idx] = utils.regress_covar(
                np.asarray(dense(mats[idx]), dtype=float), phen[idx], np.asarray(covar[idx], dtype=float), no_regress
            )
    # hard calls that were not residualised stay one byte per entry (or bit planes) on the device
    if auto and covar_q is None and (covar is None or no_regress):
        packed, tops = zip(*[hard_cache(m) for m in mats])
        if all(m is not None for m in packed):
            mats = list(packed)
            hard_calls = "b2" if bit_plane_ok(packed, tops) else "i8"
    # (row views that reach this point un-materialised -- hard calls or not -- are gathered on the device)
    for idx in range(n_pop):
        y = phen[idx] - np.mean(phen[idx])
        if not no_scale:
            y = y / np.std(y)
        phen[idx] = np.squeeze(y)

    given_resid = resolve_resid_var(resid_var, n_pop)
    if given_resid is None:
        given_resid = [np.var(y, ddof=1) for y in phen]
    check_snp_counts(n_snps, min_snps, L)
    effect_covar, adj = build_priors(n_pop, L, effect_var, rho, no_update, summary=False)

    standardize = _capi.SB_CENTER | (0 if no_scale else _capi.SB_SCALE)
    problem = _capi.IndProblem(
        mats, phen, L, pi, given_resid, effect_covar, adj.times, adj.plus, em_update=not no_update,
        standardize=standardize, covar_q=covar_q,
    )
    sample_size = np.asarray([m.shape[0] for m in mats])[:, np.newaxis]
    cs_args = dict(threshold=threshold, purity=purity, purity_method=purity_method, max_select=max_select, seed=seed)
    return _Prepared(problem, pi, sample_size, max_iter, min_tol, cs_args, no_reorder, hard_calls)


def _stage(prep: _Prepared, fit, selection=None) -> StagedResult:
    """Host half of the epilogue of one fit before the purity launch: convergence log, set selection."""
    log_convergence(fit.n_iter, prep.max_iter, prep.min_tol, fit.elbo, fit.elbo_increase)
    return StagedResult(fit, prep.pi, prep.sample_size, prep.no_reorder, **prep.cs_args, selection=selection,
                        lazy_tables=prep.lazy_tables)


def _run_prepared(preps: Sequence[_Prepared], device: int = 0, defer: bool = False) -> list:
    """Fit a list of prepared problems that share max_iter / min_tol on one GPU.

    ``defer=True`` returns one zero-argument callable per problem (the host-only result assembly) instead of
    the finished ``SushieResult`` -- the work queue runs them while the GPU is busy with the next chunk."""
    eng = _capi.get_engine(device)
    storage = config.storage_code()
    if config.genotype_storage == "auto" and all(
        p.hard_calls and p.problem.x_dtype == _capi.SB_I8 and p.problem.p <= _capi.I8_MAX_SNPS for p in preps
    ):
        storage = _capi.SB_B2 if all(p.hard_calls == "b2" for p in preps) else _capi.SB_I8
    opts = eng.make_opts(preps[0].max_iter, preps[0].min_tol, storage, config.team_size, config.single_phase,
                         config.reserve_for(device))
    return run_chunk(eng, preps, opts, _stage, defer)


def run_chunk(eng, preps, opts, stage, defer: bool) -> list:
    """One chunk of loci on one GPU:  upload -> IBSS -> effect order of every fit (from the small
    ``weighted_sum_covar`` download) -> credible-set selection on the device (sort + running sum + cut, one
    launch per width class) -> download of the posteriors straight into the effect order -> purity of every
    set of the chunk in one launch -> results (``stage`` / finishers; the pandas tables are built on first
    access).  ``config.trace_host`` prints the wall time of each phase."""
    t = [time.perf_counter()]

    def lap():
        t.append(time.perf_counter())

    with eng.create_batch([p.problem for p in preps], opts) as batch:
        lap()
        with _capi.device_run_lock(eng.device):  # one chunk at a time in the long kernel (see there)
            t_wait = time.perf_counter()
            hook = getattr(_capi.run_hooks, "on_start", None)
            if hook is not None:
                hook()
            batch.run()
            t_run = time.perf_counter()
            # (inside the lock: launched later, the sort kernels would queue behind the next chunk's IBSS kernel)
            selections = batch.select([p.cs_args["threshold"] for p in preps])
        lap()
        orders = effect_orders(batch.fetch(small_only=True), [p.no_reorder for p in preps])
        lap()
        fits = batch.fetch(orders)
        lap()
        staged = [stage(p, f, sel) for p, f, sel in zip(preps, fits, selections)]
        lap()
        min_abs = batch.purity_many([s.purity_sets for s in staged])  # one launch for the whole chunk
        lap()
    finishers = [s.finisher(m) for s, m in zip(staged, min_abs)]
    if config.trace_host:
        d = np.diff(t) * 1e3
        print(
            f"[sushie_b200] chunk of {len(preps)} fits: upload+column statistics {d[0]:.0f} ms, "
            f"wait for the GPU {1e3 * (t_wait - t[1]):.0f} ms, IBSS {1e3 * (t_run - t_wait):.0f} ms, "
            f"device set selection {1e3 * (t[2] - t_run):.0f} ms, effect order {d[2]:.0f} ms, download {d[3]:.0f} ms, "
            f"staging {d[4]:.0f} ms, purity {d[5]:.0f} ms",
            file=sys.stderr, flush=True,
        )
    if defer:
        return finishers
    t0 = time.perf_counter()
    out = [finish() for finish in finishers]
    if config.trace_host:
        print(f"[sushie_b200] chunk of {len(preps)} fits: results {1e3 * (time.perf_counter() - t0):.0f} ms",
              file=sys.stderr, flush=True)
    return out


def infer_sushie(
    Xs: List[np.ndarray],
    ys: List[np.ndarray],
    covar: utils.ListArrayOrNone = None,
    L: int = 10,
    no_scale: bool = False,
    no_regress: bool = False,
    no_update: bool = False,
    pi: np.ndarray = None,
    resid_var: utils.ListFloatOrNone = None,
    effect_var: utils.ListFloatOrNone = None,
    rho: utils.ListFloatOrNone = None,
    max_iter: int = 500,
    min_tol: float = 1e-4,
    threshold: float = 0.95,
    purity: float = 0.5,
    purity_method: str = "weighted",
    max_select: int = 250,
    min_snps: int = 100,
    no_reorder: bool = False,
    seed: int = 12345,
) -> SushieResult:
    """Fine-map one locus from individual-level data (same arguments as the reference)."""
    prep = _prepare(
        Xs, ys, covar, L, no_scale, no_regress, no_update, pi, resid_var, effect_var, rho,
        max_iter, min_tol, threshold, purity, purity_method, max_select, min_snps, no_reorder, seed,
    )
    device = (config.devices or [0])[0]
    return _run_prepared([prep], device)[0]


def infer_sushie_batch(problems: Sequence[dict], devices: Optional[Sequence[int]] = None, **common) -> List[SushieResult]:
    """Throughput entry point: many loci in one call.

    ``problems`` is a sequence of dicts with at least ``Xs`` and ``ys`` (plus any
    per-locus ``infer_sushie`` keyword); ``common`` holds keywords shared by all loci.
    One extra keyword, per locus or common: ``lazy_tables=True`` leaves the ``cs`` / ``alphas`` frames of that
    result to be built on first access (for fits whose tables are never read: cross-validation folds).
    Loci are sharded over ``devices`` by a host work queue (``sushie_b200.batch``);
    results come back in input order."""
    from . import batch as _batch

    t0 = time.perf_counter()
    preps = []
    hard_cache = _HardCallCache()  # a gene's full fit and its fold views share one hard-call check and one upload
    for pr in problems:
        kw = dict(common)
        kw.update(pr)
        preps.append(_prepare_kw(hard_cache=hard_cache, **kw))
    if config.trace_host:
        print(f"[sushie_b200] host prologue of {len(preps)} fits: {1e3 * (time.perf_counter() - t0):.0f} ms",
              file=sys.stderr, flush=True)
    return _batch.run_sharded(preps, _run_prepared, devices)


def _prepare_kw(
    Xs, ys, covar=None, L=10, no_scale=False, no_regress=False, no_update=False, pi=None, resid_var=None,
    effect_var=None, rho=None, max_iter=500, min_tol=1e-4, threshold=0.95, purity=0.5, purity_method="weighted",
    max_select=250, min_snps=100, no_reorder=False, seed=12345, hard_cache=None, lazy_tables=False,
) -> _Prepared:
    prep = _prepare(
        Xs, ys, covar, L, no_scale, no_regress, no_update, pi, resid_var, effect_var, rho,
        max_iter, min_tol, threshold, purity, purity_method, max_select, min_snps, no_reorder, seed, hard_cache,
    )
    prep.lazy_tables = bool(lazy_tables)
    return prep


def make_cs(
    alpha: np.ndarray,
    log_bf: np.ndarray,
    ns: np.ndarray,
    Xs: np.ndarray = None,
    lds: np.ndarray = None,
    threshold: float = 0.9,
    purity: float = 0.5,
    purity_method: str = "weighted",
    max_select: int = 500,
    seed: int = 12345,
):
    """Credible sets from a posterior (reference ``make_cs``, ``infer.py:835-1008``).

    ``Xs`` ([k][n_i, p], centred/scaled) or ``lds`` ([k, p, p]) is uploaded once and the
    purity Gram matrices are computed on the GPU."""
    if Xs is None and lds is None:
        raise ValueError("Both Xs and lds are None. Please specify at least one of them.")
    alpha = np.asarray(alpha, dtype=float)
    n_l, p = alpha.shape
    ns_arr = np.asarray(ns, dtype=float).reshape(-1)
    k = len(Xs) if Xs is not None else len(lds)
    eye = np.stack([np.eye(k) * 1e-3])
    eng = _capi.get_engine((config.devices or [0])[0])
    opts = eng.make_opts(1, 1e-4, _capi.SB_F64, 1)
    if Xs is not None:
        mats = [np.asarray(x, dtype=float)[: int(n)] for x, n in zip(Xs, ns_arr)]  # drop zero padding rows
        holder = _capi.IndProblem(mats, [np.zeros(m.shape[0]) for m in mats], 1, None, np.ones(k), eye,
                                  np.ones((k, k)), np.zeros((k, k)), True, standardize=0)
    else:
        holder = _capi.SsProblem([np.asarray(m, dtype=float) for m in lds], ns_arr, [np.zeros(p)] * k, 1, None,
                                 np.ones(k), eye, np.ones((k, k)), np.zeros((k, k)), True)
    with eng.create_batch([holder], opts) as batch:
        return _cs.build_tables(
            alpha, log_bf, ns_arr, lambda sets: batch.purity(0, sets), threshold, purity, purity_method,
            max_select, seed,
        )
