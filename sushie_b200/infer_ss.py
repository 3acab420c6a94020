"""Summary-level SuShiE on the GPU: drop-in for ``sushie.infer_ss`` of the reference.

``infer_sushie_ss(lds, ns, zs, ...)`` keeps the reference signature (note the argument
order), validation texts and result layout (``/root/reference/sushie/infer_ss.py:26-434``).
``XtX = ns * LD`` is never materialised: the engine streams the stored LD and applies the
sample-size factor in the kernel; ``Xty`` follows ``infer_ss.py:342-343`` on device.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import _capi, config, infer, utils
from ._common import (
    StagedResult,
    SushieResult,
    build_priors,
    check_scalar_options,
    check_snp_counts,
    log_convergence,
    resolve_pi,
    resolve_resid_var,
)

__all__ = ["infer_sushie_ss", "infer_sushie_ss_batch"]


def _prepare_ss(
    lds, ns, zs, L=10, no_update=False, pi=None, resid_var=None, effect_var=None, rho=None, max_iter=500,
    min_tol=1e-4, threshold=0.95, purity=0.5, purity_method="weighted", max_select=250, min_snps=100,
    no_reorder=False, seed=12345, lazy_tables=False,
) -> "infer._Prepared":
    ns = np.asarray(ns)
    n_pop = ns.shape[0]
    if len(lds) != n_pop:
        raise ValueError(
            f"The number of LD matrices ({len(lds)}) does not match the number of ancestries ({n_pop})."
        )
    if not all(ld.shape == lds[0].shape for ld in lds):
        raise ValueError("LD matrices do not have the same shape. Check your input.")
    if lds[0].shape[0] != lds[0].shape[1]:
        raise ValueError("LD matrices are not square matrices. Check your input.")
    if zs is None:
        raise ValueError("Z scores are not provided. Check your input.")
    if len(zs) != n_pop:
        raise ValueError(f"The number of Z scores ({len(zs)}) does not match the number of ancestries ({n_pop}).")
    if not all(z.shape == zs[0].shape for z in zs):
        raise ValueError("Z scores across ancestries do not have the same shape. Check your input.")
    if zs[0].shape[0] != lds[0].shape[0]:
        raise ValueError("Z scores do not have the same number of SNPs as the LD matrices. Check your input.")
    check_scalar_options(L, min_tol, threshold, purity, max_select, min_snps, summary=True)
    if n_pop > _capi.SB_MAX_ANCESTRIES:
        raise ValueError(
            f"sushie_b200 supports at most {_capi.SB_MAX_ANCESTRIES} ancestries per fit ({n_pop} given)."
        )
    n_snps = lds[0].shape[0]
    pi = resolve_pi(pi, n_snps, summary=True)
    given_resid = resolve_resid_var(resid_var, n_pop)
    if given_resid is None:
        given_resid = [1.0] * n_pop
    check_snp_counts(n_snps, min_snps, L)
    effect_covar, adj = build_priors(n_pop, L, effect_var, rho, no_update, summary=True)
    problem = _capi.SsProblem(
        list(lds), np.asarray(ns, dtype=float).reshape(-1), list(zs), L, pi, given_resid, effect_covar,
        adj.times, adj.plus, em_update=not no_update,
    )
    cs_args = dict(threshold=threshold, purity=purity, purity_method=purity_method, max_select=max_select, seed=seed)
    return infer._Prepared(problem, pi, ns, max_iter, min_tol, cs_args, no_reorder, lazy_tables=bool(lazy_tables))


def _stage_ss(prep, fit, selection=None) -> StagedResult:
    log_convergence(fit.n_iter, prep.max_iter, prep.min_tol, fit.elbo, fit.elbo_increase)
    return StagedResult(fit, prep.pi, prep.sample_size, prep.no_reorder, **prep.cs_args, selection=selection,
                        lazy_tables=prep.lazy_tables)


def _run_prepared_ss(preps: Sequence["infer._Prepared"], device: int = 0, defer: bool = False) -> list:
    """Summary-level twin of ``infer._run_prepared`` (``defer``: see there)."""
    eng = _capi.get_engine(device)
    opts = eng.make_opts(preps[0].max_iter, preps[0].min_tol, config.storage_code(), config.team_size, config.single_phase,
                         config.reserve_for(device))
    return infer.run_chunk(eng, preps, opts, _stage_ss, defer)


def infer_sushie_ss(
    lds: np.ndarray,
    ns: np.ndarray,
    zs: np.ndarray,
    L: int = 10,
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
    """Fine-map one locus from LD matrices and z-scores (same arguments as the reference)."""
    prep = _prepare_ss(
        lds, ns, zs, L, no_update, pi, resid_var, effect_var, rho, max_iter, min_tol, threshold, purity,
        purity_method, max_select, min_snps, no_reorder, seed,
    )
    device = (config.devices or [0])[0]
    return _run_prepared_ss([prep], device)[0]


def infer_sushie_ss_batch(problems: Sequence[dict], devices: Optional[Sequence[int]] = None, **common):
    """Many summary-level loci in one call (dicts with ``lds``, ``ns``, ``zs`` + keywords)."""
    from . import batch as _batch

    preps = []
    for pr in problems:
        kw = dict(common)
        kw.update(pr)
        preps.append(_prepare_ss(**kw))
    return _batch.run_sharded(preps, _run_prepared_ss, devices)
