"""TEST INFRASTRUCTURE ONLY -- pandas restatement of the reference's ``make_cs``.

Follows ``/root/reference/sushie/infer.py:835-1008`` step by step (same pandas calls:
``sort_values`` with its default sort kind, ``cumsum``, ``merge``), with the JAX PRNG
sub-sampling replaced by the threefry restatement in ``oracle/threefry.py``.
PARITY: pinned against the reference's own ``make_cs`` source executed through ``tests/refshim``
(``tests/test_reference_shim.py``: three purity methods, sub-sampling; identical tables) -- see the header of
``oracle/ibss.py`` for what that pin covers and what it does not (the third-party primitives).
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from . import threefry
from .ibss import make_pip


def make_cs(
    alpha,
    log_bf,
    ns,
    Xs=None,
    lds=None,
    threshold=0.9,
    purity=0.5,
    purity_method="weighted",
    max_select=500,
    seed=12345,
):
    """Credible sets, purity, alphas table, PIPs.  ``Xs`` is a list of centred/scaled
    [n_i, p] genotype matrices (the reference stacks them zero-padded, which leaves the
    Gram matrices unchanged); ``lds`` is [k, p, p]."""
    if Xs is None and lds is None:
        raise ValueError("Both Xs and lds are None. Please specify at least one of them.")
    alpha = np.asarray(alpha)
    log_bf = np.asarray(log_bf)
    ns = np.asarray(ns, dtype=float).reshape(-1, 1)
    key = threefry.prng_key(seed)
    n_l, _ = alpha.shape
    t_alpha = pd.DataFrame(alpha.T).reset_index()

    cs = pd.DataFrame(columns=["CSIndex", "SNPIndex", "alpha", "c_alpha"])
    full_alphas = t_alpha[["index"]]

    for ldx in range(n_l):
        tmp_pd = t_alpha[["index", ldx]].sort_values(ldx, ascending=False).reset_index(drop=True)
        tmp_pd["csum"] = tmp_pd[[ldx]].cumsum()
        n_row = tmp_pd[tmp_pd.csum < threshold].shape[0]
        if n_row == tmp_pd.shape[0]:
            select_idx = np.arange(n_row)
        else:
            select_idx = np.arange(n_row + 1)

        tmp_cs = (
            tmp_pd.iloc[select_idx, :]
            .assign(CSIndex=(ldx + 1))
            .rename(columns={"csum": "c_alpha", "index": "SNPIndex", ldx: "alpha"})
        )
        tmp_pd["in_cs"] = (tmp_pd.index.values <= np.max(select_idx)) * 1
        tmp_pd = tmp_pd.drop(["csum"], axis=1).rename(
            columns={"in_cs": f"in_cs_l{ldx + 1}", ldx: f"alpha_l{ldx + 1}"}
        )
        full_alphas = full_alphas.merge(tmp_pd, how="left", on="index")

        snp_idx = tmp_cs.SNPIndex.values.astype("int64")
        if len(snp_idx) > max_select:
            snp_idx = threefry.choice_without_replacement(key, snp_idx, max_select)

        if Xs is not None:
            ld = np.stack([X[:, snp_idx].T @ X[:, snp_idx] for X in Xs]) / ns[:, None]
        else:
            ld = np.asarray(lds)[:, snp_idx, :][:, :, snp_idx]

        min_abs = np.min(np.abs(ld), axis=(1, 2))
        if purity_method == "weighted":
            ss_weight = ns / np.sum(ns)
            avg_corr = np.sum(min_abs[:, None] * ss_weight)
        elif purity_method == "max":
            avg_corr = np.max(min_abs)
        elif purity_method == "min":
            avg_corr = np.min(min_abs)
        else:
            raise ValueError(
                f"Invalid purity method {purity_method}. Choose from 'weighted', 'max', or 'min'."
            )

        full_alphas[f"purity_l{ldx + 1}"] = avg_corr
        if avg_corr > purity:
            # the reference concatenates onto an empty frame whose column order
            # (CSIndex, SNPIndex, alpha, c_alpha) wins (infer.py:898,970)
            tmp_cs = tmp_cs[["CSIndex", "SNPIndex", "alpha", "c_alpha"]]
            cs = pd.concat([cs, tmp_cs], ignore_index=True) if len(cs) else tmp_cs.reset_index(drop=True)
            full_alphas[f"kept_l{ldx + 1}"] = 1
        else:
            full_alphas[f"kept_l{ldx + 1}"] = 0
        full_alphas[f"log_bf_l{ldx + 1}"] = log_bf[ldx, :]

    pip_all = make_pip(alpha)
    kept = cs.CSIndex.unique().astype(int) - 1
    pip_cs = make_pip(alpha[kept,])
    snp_in_cs = cs.SNPIndex.values.astype(int)
    cs["pip_all"] = np.array([pip_all[i] for i in snp_in_cs], dtype=float)
    cs["pip_cs"] = np.array([pip_cs[i] for i in snp_in_cs], dtype=float)
    full_alphas["pip_all"] = pip_all
    full_alphas["pip_cs"] = pip_cs
    full_alphas = full_alphas.rename(columns={"index": "SNPIndex"})
    return cs, full_alphas, pip_all, pip_cs
