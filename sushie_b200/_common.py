"""Host-side pieces shared by ``infer_sushie`` and ``infer_sushie_ss``.

Everything here mirrors behaviour of the reference that is part of the output contract
(SURVEY.md appendix A): argument validation texts (``infer.py:241-401``,
``infer_ss.py:111-259``), prior construction (``infer.py:367-445``), effect re-ordering
(``infer.py:811-832``) and the result types (``infer.py:30-101``).  Arrays are numpy where
the reference returns ``jax.Array``.
"""
from __future__ import annotations

import math
from typing import List, NamedTuple, Optional, Tuple

import numpy as np
import pandas as pd

from . import log

__all__ = ["Prior", "Posterior", "SushieResult", "_PriorAdjustor"]


class Prior(NamedTuple):
    """Prior parameters: ``pi`` [p], ``resid_var`` [k,1], ``effect_covar`` [L,k,k]."""

    pi: np.ndarray
    resid_var: np.ndarray
    effect_covar: np.ndarray


class Posterior(NamedTuple):
    """Posterior parameters: ``alpha`` [L,p], ``post_mean`` [L,p,k] (alpha-weighted),
    ``post_mean_sq`` [L,p,k,k] (alpha-weighted), ``weighted_sum_covar`` [L,k,k],
    ``kl`` [L], ``log_bf`` [L,p]."""

    alpha: np.ndarray
    post_mean: np.ndarray
    post_mean_sq: np.ndarray
    weighted_sum_covar: np.ndarray
    kl: np.ndarray
    log_bf: np.ndarray


class SushieResult:
    """Inference result; field names, order and shapes follow the reference's ``SushieResult`` named tuple
    (``infer.py:67-101``): attribute access, positional construction, unpacking, indexing, ``_fields``,
    ``_asdict`` and ``_replace`` behave as there.

    One difference: the two pandas tables (``cs``, ``alphas``) may be handed over as a zero-argument ``tables``
    callable and are then built on first access.  ``infer_sushie`` / ``infer_sushie_ss`` and the batch entry points
    build them before returning, as the reference does; a batch problem marked ``lazy_tables=True`` (the CLI marks
    cross-validation folds, whose tables nobody reads) skips the ``DataFrame`` assembly until someone asks.
    Everything the tables are made of (set selection, purity, PIPs) is always computed up front."""

    _fields = ("priors", "posteriors", "pip_all", "pip_cs", "cs", "alphas", "sample_size", "elbo", "elbo_increase",
               "l_order")
    __slots__ = ("priors", "posteriors", "pip_all", "pip_cs", "_cs", "_alphas", "sample_size", "elbo",
                 "elbo_increase", "l_order", "_tables")

    def __init__(self, priors, posteriors, pip_all, pip_cs, cs=None, alphas=None, sample_size=None, elbo=None,
                 elbo_increase=True, l_order=None, *, tables=None):
        if tables is None and (cs is None or alphas is None):
            raise TypeError("SushieResult needs `cs` and `alphas`, or a `tables` callable that builds them")
        self.priors, self.posteriors, self.pip_all, self.pip_cs = priors, posteriors, pip_all, pip_cs
        self._cs, self._alphas, self._tables = cs, alphas, tables
        self.sample_size, self.elbo, self.elbo_increase, self.l_order = sample_size, elbo, elbo_increase, l_order

    def _build(self) -> None:
        if self._tables is not None:
            self._cs, self._alphas = self._tables()
            self._tables = None

    @property
    def cs(self) -> pd.DataFrame:
        self._build()
        return self._cs

    @property
    def alphas(self) -> pd.DataFrame:
        self._build()
        return self._alphas

    def __iter__(self):
        return (getattr(self, name) for name in self._fields)

    def __len__(self):
        return len(self._fields)

    def __getitem__(self, item):
        return tuple(self)[item]

    def _asdict(self) -> dict:
        return {name: getattr(self, name) for name in self._fields}

    def _replace(self, **changes) -> "SushieResult":
        fields = self._asdict()
        unknown = set(changes) - set(fields)
        if unknown:
            raise ValueError(f"Got unexpected field names: {sorted(unknown)}")
        fields.update(changes)
        return SushieResult(**fields)

    def __reduce__(self):
        return (SushieResult, tuple(self))

    def __repr__(self) -> str:
        return "SushieResult(" + ", ".join(f"{name}={getattr(self, name)!r}" for name in self._fields) + ")"


class _PriorAdjustor(NamedTuple):
    times: np.ndarray
    plus: np.ndarray


_CLI_HINT = " Specify a valid value using '--{flag}' for command-line usage or '{kw}=' for in-Python function calls."


def check_scalar_options(L, min_tol, threshold, purity, max_select, min_snps, summary: bool) -> None:
    """Scalar argument checks common to both entry points (same texts as the reference)."""
    if L <= 0:
        raise ValueError(f"Inferred L ({L}) is invalid, choose a positive L.")
    if min_tol > 0.1:
        log.logger.warning(f"Minimum intolerance ({min_tol}) is greater than 0.1. Inference may not be accurate.")
    if not 0 < threshold < 1:
        raise ValueError(
            f"Credible set PIP threshold ({threshold}) must be greater than 0 and less than 1."
            + _CLI_HINT.format(flag="threshold", kw="threshold")
        )
    if not 0 < purity < 1:
        raise ValueError(
            f"Purity threshold ({purity}) must be greater than or equal to 0 and less than 1. "
            + _CLI_HINT.format(flag="purity", kw="purity")
        )
    if max_select <= 0:
        raise ValueError(
            "The maximum selected number of SNPs for purity must be greater than 0."
            + _CLI_HINT.format(flag="max-select", kw="max_select")
        )
    if min_snps <= 0:
        if summary:
            raise ValueError("The minimum number of SNPs to fine-map is invalid. Choose a positive integer.")
        raise ValueError(
            "The minimum number of SNPs to fine-map must be greater than 0."
            + _CLI_HINT.format(flag="min-snps", kw="min_snps")
        )


def check_snp_counts(n_snps: int, min_snps: int, L: int) -> None:
    if min_snps < L:
        raise ValueError(
            f"The number of minimum common SNPs across ancestries ({min_snps}) is less than inferred L ({L})."
            + " Specify a larger value using '--min-snps' for command-line usage"
            + " or 'min_snps=' for in-Python function calls."
        )
    if n_snps < min_snps:
        raise ValueError(
            f"The number of common SNPs across ancestries ({n_snps}) is less than minimum common"
            + " number of SNPs (100) specified."
            + " Users can specify a smaller value using '--min-snps' for command-line usage"
            + " or 'min_snps=' for in-Python function calls."
        )


def resolve_pi(pi, n_snps: int, summary: bool) -> Optional[np.ndarray]:
    """``None`` keeps the uniform default (1/p); otherwise validate and normalise.

    The two entry points normalise differently: individual level only when the sum exceeds
    one (``infer.py:313``), summary level whenever it differs from one (``infer_ss.py:189``)."""
    if pi is None:
        return None
    pi = np.asarray(pi)
    if not (pi > 0).all():
        raise ValueError("Prior probability/weights must be all positive values.")
    if pi.shape[0] != n_snps:
        raise ValueError(f"Prior probability/weights ({pi.shape[0]}) does not match the number of SNPs ({n_snps}).")
    total = np.sum(pi)
    if (total != 1) if summary else (total > 1):
        log.logger.debug("Prior probability/weights sum is not equal to 1. Will normalize to sum to 1.")
        pi = pi.astype(float) / total
    return np.asarray(pi, dtype=float)


def resolve_resid_var(resid_var, n_pop: int) -> Optional[List[float]]:
    if resid_var is None:
        return None
    if len(resid_var) != n_pop:
        raise ValueError(
            f"Number of specified residual prior ({len(resid_var)}) does not match ancestry number ({n_pop})."
        )
    resid_var = [float(i) for i in resid_var]
    if np.any(np.array(resid_var) <= 0):
        raise ValueError(f"The input of residual prior ({resid_var}) is invalid (<0). Check your input.")
    return resid_var


def build_priors(n_pop: int, L: int, effect_var, rho, no_update: bool, summary: bool):
    """Initial effect covariance [L,k,k] and the ``_PriorAdjustor`` (``infer.py:367-445``)."""
    given_var, given_rho = effect_var, rho
    if effect_var is None:
        effect_var = [1e-3] * n_pop
    else:
        if len(effect_var) != n_pop:
            raise ValueError(
                f"Number of specified effect prior ({len(effect_var)}) does not match ancestry number ({n_pop})."
            )
        effect_var = [float(i) for i in effect_var]
        if np.any(np.array(effect_var) <= 0):
            raise ValueError(f"The effect size prior variance ({effect_var}) must be positive.")

    n_rho = math.comb(n_pop, 2)
    if rho is None:
        rho = [0.1] * n_rho
    else:
        if n_pop == 1:
            log.logger.debug("Running single-ancestry SuShiE. The '--rho' parameter is specified but will be ignored.")
        if (len(rho) != n_rho) and n_pop != 1:
            raise ValueError(f"Number of specified rho ({len(rho)}) does not match expected" + f" number {n_rho}.")
        rho = [float(i) for i in rho]
        # the individual-level entry point rejects |rho| == 1, the summary-level one accepts it
        bad = (np.abs(np.array(rho)) > 1) if summary else (np.abs(np.array(rho)) >= 1)
        if np.any(bad):
            raise ValueError(f"Effect size prior correlation ({rho}) must be between -1 and 1 (inclusive).")

    c0 = np.diag(np.asarray(effect_var, dtype=float))
    ct = 0
    for col in range(n_pop):
        for row in range(1, n_pop):
            if col < row:
                cov = rho[ct] * np.sqrt(effect_var[row] * effect_var[col])
                c0[row, col] = cov
                c0[col, row] = cov
                ct += 1

    eye = np.eye(n_pop)
    if no_update:
        if given_var is None and given_rho is not None and n_pop != 1:
            adj = _PriorAdjustor(times=eye, plus=c0 - np.diag(np.diag(c0)))
            log.logger.info("No updates on the prior effect correlation rho while updating prior effect variance.")
        elif given_var is not None and given_rho is None and n_pop != 1:
            adj = _PriorAdjustor(times=np.ones((n_pop, n_pop)) - eye, plus=c0 * eye)
            log.logger.info("No updates on the prior effect variance while updating prior effect correlation rho.")
        else:
            adj = _PriorAdjustor(times=np.zeros((n_pop, n_pop)), plus=c0.copy())
            log.logger.info("No updates on the prior effect size variance/covariance matrix.")
    else:
        adj = _PriorAdjustor(times=np.ones((n_pop, n_pop)), plus=np.zeros((n_pop, n_pop)))
    return np.stack([c0] * L), adj


def _reorder_l(priors: Prior, posteriors: Posterior) -> Tuple[Prior, Posterior, np.ndarray]:
    """Sort effects by descending nuclear norm of ``weighted_sum_covar`` (stable)."""
    with np.errstate(all="ignore"):
        try:
            norm = np.linalg.svd(posteriors.weighted_sum_covar, compute_uv=False).sum(axis=1)
        except np.linalg.LinAlgError:
            norm = np.full(posteriors.weighted_sum_covar.shape[0], np.nan)
    l_order = np.argsort(-norm, kind="stable")
    priors = priors._replace(effect_covar=priors.effect_covar[l_order])
    posteriors = Posterior(*[field[l_order] for field in posteriors])
    return priors, posteriors, l_order


def log_convergence(n_iter: int, max_iter: int, min_tol: float, elbo: np.ndarray, elbo_increase: bool) -> None:
    """The reference's per-fit log lines (``infer.py:523-548``); users grep for them."""
    cur = elbo[-1] if len(elbo) else float("nan")
    if not elbo_increase:
        log.logger.warning(
            f"Optimization finished after {n_iter} iterations."
            + f" ELBO decreased. Final ELBO score: {cur}. Return last iteration's results."
            + " It can be precision issue,"
            + " and adding 'import jax; jax.config.update('jax_enable_x64', True)' may fix it."
            + " If this issue keeps rising for many genes, contact the developers."
        )
        return
    digits = len(str(min_tol)) - str(min_tol).find(".") - 1
    last = elbo[-2] if len(elbo) > 1 else -np.inf
    if abs(cur - last) < min_tol:
        log.logger.info(
            f"Optimization concludes after {n_iter} iterations. Final ELBO score: {cur:.{digits}f}."
            + f" Reach minimum tolerance threshold {min_tol}.",
        )
    elif n_iter == max_iter:
        log.logger.info(
            f"Optimization concludes after {n_iter} iterations. Final ELBO score: {cur:.{digits}f}."
            + f" Reach maximum iteration threshold {max_iter}.",
        )


def effect_orders(fits, no_reorder: List[bool]) -> List[Optional[np.ndarray]]:
    """``_reorder_l``'s effect order of every fit of a chunk from its ``weighted_sum_covar`` (None where
    ``no_reorder``): the same SVD and stable argsort as :func:`_reorder_l`, but one LAPACK-batched call per
    (L, k) shape instead of one numpy call per fit.  The big posterior arrays are then downloaded straight into
    this order (``sb_result.effect_order``)."""
    out: List[Optional[np.ndarray]] = [None] * len(fits)
    groups = {}
    for i, (fit, skip) in enumerate(zip(fits, no_reorder)):
        if not skip:
            groups.setdefault(fit.weighted_sum_covar.shape, []).append(i)
    for shape, ids in groups.items():
        stack = np.concatenate([fits[i].weighted_sum_covar for i in ids])
        with np.errstate(all="ignore"):
            try:
                norms = np.linalg.svd(stack, compute_uv=False).sum(axis=1).reshape(len(ids), shape[0])
            except np.linalg.LinAlgError:  # some fit of the group is not finite: decide one by one
                norms = None
        for row, i in enumerate(ids):
            if norms is not None:
                norm = norms[row]
            else:
                with np.errstate(all="ignore"):
                    try:
                        norm = np.linalg.svd(fits[i].weighted_sum_covar, compute_uv=False).sum(axis=1)
                    except np.linalg.LinAlgError:
                        norm = np.full(shape[0], np.nan)
            out[i] = np.argsort(-norm, kind="stable").astype(np.int32)
    return out


def _reordered(fit, pi, no_reorder):
    k = fit.resid_var.shape[0]
    L, p = fit.alpha.shape
    if pi is None:
        pi = np.ones(p) / float(p)
    priors = Prior(pi=pi, resid_var=fit.resid_var.reshape(k, 1), effect_covar=fit.effect_covar)
    posteriors = Posterior(
        alpha=fit.alpha,
        post_mean=fit.post_mean,
        post_mean_sq=fit.post_mean_sq,
        weighted_sum_covar=fit.weighted_sum_covar,
        kl=fit.kl,
        log_bf=fit.log_bf,
    )
    l_order = np.arange(L)
    done = getattr(fit, "l_order", None)  # the download has put the effects in this order already
    if done is not None:
        log.logger.debug("Reordering effects based on Frobenius norm of effect size covariance prior.")
        l_order = np.asarray(done, dtype=np.int64)
    elif not no_reorder:
        log.logger.debug("Reordering effects based on Frobenius norm of effect size covariance prior.")
        priors, posteriors, l_order = _reorder_l(priors, posteriors)
    return priors, posteriors, l_order


def assemble_result(fit, pi, sample_size, no_reorder, cs_builder) -> SushieResult:
    """RawFit (engine output) -> SushieResult: re-order effects, then credible sets."""
    priors, posteriors, l_order = _reordered(fit, pi, no_reorder)
    log.logger.debug("Computing credible sets.")
    cs, full_alphas, pip_all, pip_cs = cs_builder(posteriors.alpha, posteriors.log_bf)
    log.logger.debug("Inference and credible set computation complete. Beginning to write results.")
    return SushieResult(
        priors, posteriors, pip_all, pip_cs, cs, full_alphas, sample_size, fit.elbo, fit.elbo_increase, l_order
    )


class StagedResult:
    """Epilogue of one fit in two halves.  Construction does the effect re-ordering (unless the download did it)
    and the credible-set selection -- from the device's sort / running sum when ``selection`` is given, else on
    the host; ``purity_sets`` is what the purity kernel has to look at (one index list per effect).  ``finisher``
    takes the min |corr| per (effect, ancestry) and returns the rest as a zero-argument callable, so a batch
    entry point can ask for the purity of a whole chunk of loci in one launch, release the device batch and
    start the next chunk before results are assembled."""

    def __init__(self, fit, pi, sample_size, no_reorder, threshold, purity, purity_method, max_select, seed,
                 selection=None, lazy_tables=False):
        from . import cs as _cs

        if purity_method not in ("weighted", "max", "min"):
            raise ValueError(f"Invalid purity method {purity_method}. Choose from 'weighted', 'max', or 'min'.")
        self._fit, self._sample_size = fit, sample_size
        self._purity, self._purity_method = purity, purity_method
        self._lazy_tables = lazy_tables
        self._priors, self._posteriors, self._l_order = _reordered(fit, pi, no_reorder)
        log.logger.debug("Computing credible sets.")
        if selection is not None:
            self._selection = _cs.selection_from_device(*selection, self._l_order, max_select, seed)
        else:
            self._selection = _cs.select_sets(self._posteriors.alpha, threshold, max_select, seed)
        self.purity_sets = self._selection[3]

    def finisher(self, min_abs):
        from . import cs as _cs

        min_abs = np.asarray(min_abs, dtype=float)

        def finish() -> SushieResult:
            # purity, kept sets and PIPs now (cheap, and the reference's log line needs them); the two pandas
            # tables here too (the batch entry points run this on a finisher thread, under the GPU's next chunk)
            # unless the caller said it will not look at them (`lazy_tables`: cross-validation folds)
            alpha, log_bf = self._posteriors.alpha, self._posteriors.log_bf
            selection = self._selection
            purities, kept, pip_all, pip_cs = _cs.summarise_sets(
                alpha, self._sample_size, min_abs, self._purity, self._purity_method, selection)

            def tables():
                return _cs.frames(alpha, log_bf, selection, purities, kept, pip_all, pip_cs)

            log.logger.debug("Inference and credible set computation complete. Beginning to write results.")
            result = SushieResult(
                self._priors, self._posteriors, pip_all, pip_cs, None, None, self._sample_size, self._fit.elbo,
                self._fit.elbo_increase, self._l_order, tables=tables,
            )
            if not self._lazy_tables:
                result._build()
            return result

        return finish
