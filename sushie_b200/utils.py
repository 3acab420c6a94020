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
from scipy import special, stats

__all__ = ["make_pip", "rint", "ols", "regress_covar", "RowView", "ListFloatOrNone", "ListArrayOrNone", "IntOrNone"]

ListFloatOrNone = Optional[List[float]]
ListArrayOrNone = Optional[List[np.ndarray]]
IntOrNone = Optional[int]


def make_pip(alpha: np.ndarray) -> np.ndarray:
    """Posterior inclusion probability across effects: 1 - prod_l (1 - alpha_l)."""
    alpha = np.asarray(alpha, dtype=float)
    with np.errstate(divide="ignore"):
        return -np.expm1(np.log1p(-alpha).sum(axis=0))


def rint(y_val: np.ndarray) -> np.ndarray:
    """Rank inverse-normal transform: ``norm.ppf(rankdata(y) / (n + 1))`` (average ranks for ties).

    One-dimensional finite input takes a direct path (stable argsort + ``ndtri``, the function ``norm.ppf``
    evaluates) that returns the same bits as the scipy calls at a fifth of their cost -- cross-validation
    transforms every fold's phenotypes of every gene (``cli.py:351-356``)."""
    y_val = np.asarray(y_val)
    n_pt = y_val.shape[0]
    if y_val.ndim != 1 or n_pt == 0 or y_val.dtype.kind not in "fiu" or (y_val.dtype.kind == "f" and np.isnan(y_val).any()):
        return stats.norm.ppf(stats.rankdata(y_val) / (n_pt + 1))
    sorter = np.argsort(y_val, kind="stable")
    ranks = np.empty(n_pt, dtype=float)
    ys = y_val[sorter]
    first = np.empty(n_pt, dtype=bool)  # first entry of every run of equal values
    first[0] = True
    np.not_equal(ys[1:], ys[:-1], out=first[1:])
    if first.all():
        ranks[sorter] = np.arange(1, n_pt + 1, dtype=float)
    else:
        dense = np.cumsum(first)  # 1-based run number of every sorted entry
        bounds = np.append(np.flatnonzero(first), n_pt)  # start of every run (+ the end)
        ranks[sorter] = 0.5 * (bounds[dense] + bounds[dense - 1] + 1)
    return special.ndtri(ranks / (n_pt + 1))


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


class RowView:
    """Rows ``rows`` (in that order) of the 2-D array ``base``, not materialised.

    A cross-validation fold is a row subset of its gene's shuffled genotype matrix (``cli.py:323-375``).  Passing
    the folds to ``infer_sushie_batch`` as views lets the engine upload the gene's matrix once and gather every
    fold's rows on the device (``sb_ind_problem.x_rows``) instead of building and uploading five 80 % copies.
    ``np.asarray(view)`` gives the materialised matrix, so anything else can treat it as an array."""

    __slots__ = ("base", "rows")
    ndim = 2

    def __init__(self, base, rows):
        base = np.asarray(base)
        if base.ndim != 2:
            raise ValueError("RowView needs a 2-D base array")
        rows = np.ascontiguousarray(rows, dtype=np.int32).reshape(-1)
        if rows.size and (rows.min() < 0 or rows.max() >= base.shape[0]):
            raise IndexError("RowView row index outside the base array")
        self.base, self.rows = base, rows

    @property
    def shape(self):
        return (int(self.rows.size), int(self.base.shape[1]))

    @property
    def dtype(self):
        return self.base.dtype

    def __len__(self):
        return int(self.rows.size)

    def __array__(self, dtype=None, copy=None):
        out = self.base[self.rows]
        return out if dtype is None else out.astype(dtype, copy=False)

    def __getitem__(self, key):
        return np.asarray(self)[key]


def array_key(a: np.ndarray):
    """Identity of the memory an array looks at (two views of one genotype matrix compare equal)."""
    return (a.__array_interface__["data"][0], a.shape, a.strides, a.dtype.str)
