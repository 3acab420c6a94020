"""Numeric utilities on the inference path (host side, numpy/scipy).

Mirrors the reference's ``sushie/utils.py`` for the functions the IBSS path and its
callers use: ``make_pip`` (``utils.py:37-51``), ``rint`` (``utils.py:54-69``), ``ols``
(``utils.py:72-111``) and ``regress_covar`` (``utils.py:114-136``).  ``estimate_her`` (a
third-party LMM, ``utils.py:139-210``) is outside the hot path and not provided.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
from scipy import linalg as sla
from scipy import stats

__all__ = ["make_pip", "rint", "ols", "regress_covar", "ListFloatOrNone", "ListArrayOrNone", "IntOrNone"]

ListFloatOrNone = Optional[List[float]]
ListArrayOrNone = Optional[List[np.ndarray]]
IntOrNone = Optional[int]


def make_pip(alpha: np.ndarray) -> np.ndarray:
    """Posterior inclusion probability across effects: 1 - prod_l (1 - alpha_l)."""
    alpha = np.asarray(alpha, dtype=float)
    with np.errstate(divide="ignore"):
        return -np.expm1(np.log1p(-alpha).sum(axis=0))


def rint(y_val: np.ndarray) -> np.ndarray:
    """Rank inverse-normal transform."""
    y_val = np.asarray(y_val)
    n_pt = y_val.shape[0]
    return stats.norm.ppf(stats.rankdata(y_val) / (n_pt + 1))


def ols(X: np.ndarray, y: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """OLS of each column of ``y`` on ``[1, X]`` by reduced QR.

    Returns (residuals, adjusted r-squared, p-values of the coefficients)."""
    X = np.asarray(X, dtype=float)
    design = np.column_stack([np.ones(X.shape[0]), X.reshape(X.shape[0], -1)])
    y = np.reshape(np.asarray(y, dtype=float), (len(y), -1))
    q, r = np.linalg.qr(design, mode="reduced")
    qty = q.T @ y
    beta = sla.solve_triangular(r, qty)
    n_obs, n_par = q.shape
    df = n_obs - n_par
    residual = y - q @ qty
    rss = np.sum(residual ** 2, axis=0)
    sigma = np.sqrt(rss / df)
    # diag((R^T R)^-1) through a Cholesky-style solve with the upper factor
    xtx_inv_diag = np.diag(sla.cho_solve((r, False), np.eye(n_par)))
    se = np.sqrt(xtx_inv_diag)[:, None] @ sigma[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        t_scores = beta / se
    p_value = np.asarray(2 * stats.t.sf(np.abs(t_scores), df=df))
    tss = np.sum((y - np.mean(y, axis=0)) ** 2, axis=0)
    r_sq = 1 - rss / tss
    adj_r = 1 - (1 - r_sq) * (n_obs - 1) / df
    return residual, adj_r, p_value


def covar_basis(covar: np.ndarray) -> np.ndarray:
    """Orthonormal basis Q of ``[1, covar]`` -- the reduced-QR factor ``ols`` residualises with: ``y - Q Q^T y``."""
    covar = np.asarray(covar, dtype=float)
    design = np.column_stack([np.ones(covar.shape[0]), covar.reshape(covar.shape[0], -1)])
    q, _ = np.linalg.qr(design, mode="reduced")
    return np.ascontiguousarray(q)


def regress_covar(X: np.ndarray, y: np.ndarray, covar: np.ndarray, no_regress: bool) -> Tuple[np.ndarray, np.ndarray]:
    """Residualise the phenotype (and, unless ``no_regress``, the genotypes) on covariates."""
    y, _, _ = ols(covar, y)
    if not no_regress:
        X, _, _ = ols(covar, X)
    return X, y


def as_hard_calls(X) -> "np.ndarray | None":
    """``X`` as a C-contiguous int8 matrix when every entry is an integer in 0..127 (hard-call
    genotypes 0/1/2), else ``None``.  Used to pick the one-byte device storage."""
    X = np.asarray(X)
    if X.ndim != 2 or X.size == 0:
        return None
    if X.dtype.kind in "iu":
        if X.min() < 0 or X.max() > 127:
            return None
        return np.ascontiguousarray(X, dtype=np.int8)
    if X.dtype.kind != "f":
        return None
    head = X[: min(4, X.shape[0])]
    if not np.array_equal(head, np.rint(head)):  # residualised / dosage data: reject without a full pass
        return None
    lo, hi = X.min(), X.max()
    if not (np.isfinite(lo) and np.isfinite(hi)) or lo < 0 or hi > 127:
        return None
    Xi = X.astype(np.int8)
    if not np.array_equal(Xi, X):
        return None
    return np.ascontiguousarray(Xi)
