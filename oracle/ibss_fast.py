"""TEST / BENCH INFRASTRUCTURE ONLY -- the individual-level IBSS loop of ``oracle/ibss.py`` in "fast form".

``oracle/ibss.py`` restates the reference operation by operation (three einsums over X per single-effect update,
ERSS through ``X`` and ``X**2``, LAPACK inverse / slogdet / Cholesky per SNP) and is the parity oracle.  This module
computes the SAME quantities with the algebraic restructurings of SURVEY.md 7.3 so that the CPU baseline arm of
``bench.py`` (``--impl reference`` and ``cpu_baseline``) finishes inside the driver's window:

  * ``X b_l`` is cached per effect, so a single-effect update traverses X twice instead of three times
    (``infer.py:640,653,676``);
  * ERSS needs no pass over X: ``|y - X sum_l b_l|^2`` is the running residual, ``sum_n (X b_l)^2`` the cached vector
    and ``sum_n X^2 E[b^2] = sum_p XtX E[b^2]`` (``infer.py:799-808``);
  * the per-SNP ``k x k`` posterior (``infer.py:704-719``) is a closed-form LDL^T over arrays of length p instead of
    batched LAPACK calls.

Same stopping rule, same outputs; ``tests/test_oracle.py::test_fast_form_equals_literal_form`` holds it to the literal
form (same iteration count, ELBO to 1e-9 relative, alpha to 1e-9).  It is a faster CPU implementation than the
literal one, i.e. a *stronger* baseline for the GPU path to be compared with.  Never imported by ``sushie_b200``.
"""
from __future__ import annotations

import math

import numpy as np

from . import ibss

LOG_2PI = math.log(2.0 * math.pi)


def _posterior(r, d, log_pi, C):
    """r, d: [k][p] (``X^T r / s2`` and ``XtX / s2``); C: [k, k].  Returns alpha [p], mu [k][p] (unweighted),
    S [k][k][p], log_bf [p], weighted_sum_covar [k, k], kl (scalar)."""
    k = len(r)
    with np.errstate(all="ignore"):
        Cinv = np.linalg.inv(C)
        _, logdetC = np.linalg.slogdet(C)
        # A = diag(d) + Cinv = Lm D Lm^T
        Lm = [[None] * k for _ in range(k)]
        D, Dinv = [None] * k, [None] * k
        for j in range(k):
            dj = Cinv[j, j] + d[j]
            for t in range(j):
                dj = dj - Lm[j][t] * Lm[j][t] * D[t]
            dj = np.where(dj > 0.0, dj, np.nan)
            D[j], Dinv[j] = dj, 1.0 / dj
            for i in range(j + 1, k):
                s = Cinv[i, j]
                for t in range(j):
                    s = s - Lm[i][t] * Lm[j][t] * D[t]
                Lm[i][j] = s * Dinv[j]
        logdetA = np.log(D[0])
        for j in range(1, k):
            logdetA = logdetA + np.log(D[j])
        # Li = Lm^-1 (unit lower)
        Li = [[None] * k for _ in range(k)]
        for j in range(k):
            for i in range(j + 1, k):
                s = -Lm[i][j]
                for t in range(j + 1, i):
                    s = s - Lm[i][t] * Li[t][j]
                Li[i][j] = s
        one = np.ones_like(D[0])

        def li(t, a):
            return one if t == a else Li[t][a]

        S = [[None] * k for _ in range(k)]
        for a in range(k):
            for b in range(a, k):
                s = 0.0
                for t in range(b, k):
                    s = s + li(t, a) * Dinv[t] * li(t, b)
                S[a][b] = S[b][a] = s
        mu = [sum(S[a][b] * r[b] for b in range(k)) for a in range(k)]
        q = sum(r[a] * mu[a] for a in range(k))
        log_bf = 0.5 * (k * LOG_2PI - logdetA + q)
        w = log_pi + log_bf
        u = np.exp(w - np.max(w))
        alpha = u / np.sum(u)
        M = [[S[a][b] + mu[a] * mu[b] for b in range(k)] for a in range(k)]
        wsc = np.array([[np.sum(alpha * M[a][b]) for b in range(k)] for a in range(k)])
        tr = sum(Cinv[a, b] * S[b][a] for a in range(k) for b in range(k))
        quad = sum(mu[a] * Cinv[a, b] * mu[b] for a in range(k) for b in range(k))
        kl_beta = 0.5 * (tr - k + quad + logdetC + logdetA)
        kl = np.nansum(alpha * (np.log(alpha) - log_pi)) + alpha @ kl_beta
    return alpha, mu, M, log_bf, wsc, kl


def fit_individual(Xs, ys, L=10, max_iter=500, min_tol=1e-4, no_update=False):
    """``ibss.fit_individual`` with default priors (no covariates, scaled data), fast form.  Returns an object with
    ``n_iter``, ``elbo``, ``elbo_increase``, ``alpha`` [L, p], ``post_mean`` [L, p, k], ``effect_covar``, ``resid_var``."""
    k = len(Xs)
    Xs, ys = ibss.prepare_individual(Xs, ys)
    n = [x.shape[0] for x in Xs]
    p = Xs[0].shape[1]
    XtX = [np.sum(x * x, axis=0) for x in Xs]
    Xt = [np.ascontiguousarray(x.T) for x in Xs]  # both traversals as row-major matrix-vector products
    log_pi = np.log(np.ones(p) / float(p))
    resid_var = np.asarray([np.var(y, ddof=1) for y in ys])
    c0 = ibss.build_effect_covar(k, [1e-3] * k, [0.1] * math.comb(k, 2))
    adj_times, adj_plus = ibss.build_adjustor(k, c0, no_update, None, None)
    ecov = np.stack([c0] * L)
    state = dict(
        alpha=np.zeros((L, p)), pm=np.zeros((L, k, p)), msq_diag=np.zeros((L, k, p)), wsc=np.zeros((L, k, k)),
        kl=np.zeros(L), log_bf=np.zeros((L, p)),
    )
    xb = [np.zeros((L, n[i])) for i in range(k)]
    resid = [ys[i].copy() for i in range(k)]  # y - sum_l X b_l
    elbo = [-np.inf]
    inc = True
    n_iter = 0
    for it in range(max_iter):
        prev = (ecov.copy(), resid_var.copy(), {a: b.copy() for a, b in state.items()}, [x.copy() for x in xb],
                [r.copy() for r in resid])
        for l in range(L):
            r_l = [resid[i] + xb[i][l] for i in range(k)]
            rz = [(Xt[i] @ r_l[i]) / resid_var[i] for i in range(k)]
            dz = [XtX[i] / resid_var[i] for i in range(k)]
            first = _posterior(rz, dz, log_pi, ecov[l])
            new_c = first[4]
            if no_update:
                new_c = new_c * adj_times + adj_plus
            alpha, mu, M, log_bf, wsc, kl = _posterior(rz, dz, log_pi, new_c)
            ecov[l] = new_c
            state["alpha"][l], state["log_bf"][l], state["wsc"][l], state["kl"][l] = alpha, log_bf, wsc, kl
            for i in range(k):
                state["pm"][l, i] = alpha * mu[i]
                state["msq_diag"][l, i] = alpha * M[i][i]
                xb[i][l] = Xs[i] @ state["pm"][l, i]
                resid[i] = r_l[i] - xb[i][l]
        erss = np.zeros(k)
        for i in range(k):
            erss[i] = resid[i] @ resid[i] + np.sum(XtX[i] * state["msq_diag"][:, i, :]) - np.sum(xb[i] * xb[i])
        with np.errstate(all="ignore"):
            exp_ll = np.sum(-(0.5 * np.asarray(n)) * np.log(2 * np.pi * resid_var) - (0.5 / resid_var) * erss)
        cur = float(exp_ll - np.sum(state["kl"]))
        resid_var = erss / np.asarray(n)
        n_iter = it + 1
        last = elbo[it]
        elbo.append(cur)
        with np.errstate(invalid="ignore"):
            inc = bool(cur >= last or np.isclose(cur, last, atol=1e-8))
        if not inc:
            ecov, resid_var, state, xb, resid = prev
            break
        if abs(cur - last) < min_tol:
            break

    class Fit:
        pass

    out = Fit()
    out.n_iter, out.elbo, out.elbo_increase = n_iter, np.asarray(elbo), inc
    out.alpha = state["alpha"]
    out.post_mean = np.transpose(state["pm"], (0, 2, 1))
    out.weighted_sum_covar, out.kl, out.log_bf = state["wsc"], state["kl"], state["log_bf"]
    out.effect_covar, out.resid_var = ecov, resid_var
    return out
