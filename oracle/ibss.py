"""TEST INFRASTRUCTURE ONLY -- numpy-fp64 restatement of SuShiE's IBSS loop.

This module is the *oracle* the CUDA path is checked against.  It is never imported by
the product package ``sushie_b200`` (only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it).

PARITY: the reference (``/root/reference/sushie``) is pure Python on top of ``jax==0.4.33`` /
``equinox==0.10.2`` (``setup.cfg:35-40``), neither of which is installed or installable in this
environment, and the reference's own tests pin no numbers (``tests/test_infer.py:57,93`` only assert
``res is not None``).  The oracle is therefore pinned against the REFERENCE'S OWN SOURCE executed through
``tests/refshim`` (numpy-backed stand-ins for the jax / equinox primitives; ``tests/test_reference_shim.py``
compares it with this module on 20+ option mixes: identical iteration counts, effect order and credible sets,
ELBO to 1e-10 relative) and the committed goldens under ``tests/golden/`` are outputs of that run
(``tests/golden/make_golden.py``).  What stays unpinned is the third-party arithmetic underneath (XLA's LU /
Cholesky / reductions versus LAPACK / numpy) -- rounding-level differences.  Further anchors: the documented
causal SNPs of the bundled example, rs1886340 and rs10914958 (``docs/manual.rst:88``), which come out with
PIP = 1, and independent scipy identities (``tests/test_oracle.py``).  Everything below follows the reference
operation by operation ("literal form"); each function cites the lines it restates.

Primitive mapping (third-party, restated from their published definitions):
  jnp.linalg.inv            -> numpy.linalg.inv (LU)
  jnp.linalg.slogdet        -> numpy.linalg.slogdet (LU)
  jax.scipy.stats.multivariate_normal.logpdf
                            -> Cholesky of the covariance + triangular solve, NaN when
                               the covariance is not positive definite (as in JAX)
  jax.nn.softmax            -> exp(x - max x) / sum exp(x - max x)
  jnp.nansum                -> numpy.nansum
  jnp.isclose               -> numpy.isclose (rtol=1e-5 default, atol as given)
"""
from __future__ import annotations

import math
from typing import List, NamedTuple, Optional, Sequence

import numpy as np

LOG_2PI = math.log(2.0 * math.pi)


class OraclePrior(NamedTuple):
    pi: np.ndarray  # [p]
    resid_var: np.ndarray  # [k, 1]
    effect_covar: np.ndarray  # [L, k, k]


class OraclePosterior(NamedTuple):
    alpha: np.ndarray  # [L, p]
    post_mean: np.ndarray  # [L, p, k]
    post_mean_sq: np.ndarray  # [L, p, k, k]
    weighted_sum_covar: np.ndarray  # [L, k, k]
    kl: np.ndarray  # [L]
    log_bf: np.ndarray  # [L, p]


class OracleFit(NamedTuple):
    """What the IBSS loop returns before credible sets are built."""

    priors: OraclePrior
    posteriors: OraclePosterior
    sample_size: np.ndarray  # [k, 1]
    elbo: np.ndarray  # trace with leading -inf
    elbo_increase: bool
    l_order: np.ndarray  # [L]
    n_iter: int
    Xs: Optional[List[np.ndarray]]  # centred/scaled genotypes (individual level) or None
    lds: Optional[np.ndarray]  # [k, p, p] (summary level) or None


# --------------------------------------------------------------------------------------
# small helpers


def ols_residual(X: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Residual of y on [1, X] by reduced QR.  Reference ``utils.py:72-111`` (only the
    residual is needed on the inference path, ``utils.py:131-136``)."""
    Xi = np.concatenate([np.ones((X.shape[0], 1)), X], axis=1)
    y2 = np.reshape(y, (len(y), -1))
    q, _ = np.linalg.qr(Xi, mode="reduced")
    return y2 - q @ (q.T @ y2)


def regress_covar(X, y, covar, no_regress):
    """Reference ``utils.py:114-136``."""
    y = ols_residual(covar, y)
    if not no_regress:
        X = ols_residual(covar, X)
    return X, y


def make_pip(alpha: np.ndarray) -> np.ndarray:
    """Reference ``utils.py:37-51``."""
    with np.errstate(divide="ignore"):
        return -np.expm1(np.sum(np.log1p(-alpha), axis=0))


def _mvn_logpdf_zero_mean(x: np.ndarray, cov: np.ndarray) -> np.ndarray:
    """log N(x; 0, cov) batched over the leading axis, JAX's algorithm: Cholesky factor,
    triangular solve, -(k/2) log 2pi - sum log diag(L) - |L^-1 x|^2 / 2.  Non-PD -> NaN."""
    p, k = x.shape
    out = np.full(p, np.nan)
    with np.errstate(invalid="ignore"):
        ok = np.all(np.isfinite(cov.reshape(p, -1)), axis=1)
    try:
        chol = np.linalg.cholesky(cov[ok])
        good = np.nonzero(ok)[0]
    except np.linalg.LinAlgError:
        # fall back to per-matrix factorisation so one bad SNP only poisons itself
        good_list, chol_list = [], []
        for j in np.nonzero(ok)[0]:
            try:
                chol_list.append(np.linalg.cholesky(cov[j]))
                good_list.append(j)
            except np.linalg.LinAlgError:
                pass
        good = np.asarray(good_list, dtype=int)
        chol = np.asarray(chol_list).reshape(len(good_list), k, k)
    if len(good):
        # forward substitution, batched
        yv = np.zeros((len(good), k))
        xb = x[good]
        for i in range(k):
            acc = xb[:, i].copy()
            for j in range(i):
                acc -= chol[:, i, j] * yv[:, j]
            yv[:, i] = acc / chol[:, i, i]
        logdet_half = np.sum(np.log(np.diagonal(chol, axis1=1, axis2=2)), axis=1)
        out[good] = -0.5 * np.sum(yv * yv, axis=1) - 0.5 * k * LOG_2PI - logdet_half
    return out


def _softmax(x: np.ndarray) -> np.ndarray:
    with np.errstate(invalid="ignore", over="ignore"):
        u = np.exp(x - np.max(x))
        return u / np.sum(u)


def kl_categorical(alpha, pi):
    """Reference ``infer.py:748-752``."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.nansum(alpha * (np.log(alpha) - np.log(pi)))


def kl_mvn(m0, sigma0, sigma1):
    """KL(N(m0, sigma0) || N(0, sigma1)) per SNP.  Reference ``infer.py:755-783``."""
    k = sigma1.shape[0]
    sigma1_inv = np.linalg.inv(sigma1)
    diff = 0.0 - m0
    p1 = np.einsum("km,pmk->p", sigma1_inv, sigma0) - k
    p2 = np.einsum("pk,km,pm->p", diff, sigma1_inv, diff)
    with np.errstate(divide="ignore", invalid="ignore"):
        _, sld1 = np.linalg.slogdet(sigma1)
        _, sld0 = np.linalg.slogdet(sigma0)
    return 0.5 * (p1 + p2 + (sld1 - sld0))


# --------------------------------------------------------------------------------------
# the single-effect posterior (reference infer.py:692-745)


def compute_posterior(rTZDinv, inv_shat2_diag, pi, prior_covar):
    """One call of the reference's ``_compute_posterior`` for a single effect.

    rTZDinv [p,k]; inv_shat2_diag [p,k] is the diagonal the reference expands to
    [p,k,k] (``infer.py:680-683``).  Returns the per-effect rows and the EM prior
    (``weighted_sum_covar``)."""
    p, k = rTZDinv.shape
    inv_shat2 = np.eye(k)[None, :, :] * inv_shat2_diag[:, None, :]
    with np.errstate(all="ignore"):
        post_covar = np.linalg.inv(inv_shat2 + np.linalg.inv(prior_covar))
        post_mean = np.einsum("pkm,pm->pk", post_covar, rTZDinv)
        post_mean_sq = post_covar + np.einsum("pk,pm->pkm", post_mean, post_mean)
        log_bf = -1.0 * _mvn_logpdf_zero_mean(post_mean, post_covar)
        alpha = _softmax(np.log(pi) + log_bf)
        w_mean = post_mean * alpha[:, None]
        w_mean_sq = post_mean_sq * alpha[:, None, None]
        wsc = np.sum(w_mean_sq, axis=0)
        kl = kl_categorical(alpha, pi) + alpha @ kl_mvn(post_mean, post_covar, prior_covar)
    return dict(
        alpha=alpha,
        post_mean=w_mean,
        post_mean_sq=w_mean_sq,
        weighted_sum_covar=wsc,
        kl=kl,
        log_bf=log_bf,
    )


def _ser(rTZDinv, inv_shat2_diag, pi, effect_covar_l, adj_times, adj_plus, em):
    """Prior update + posterior of one single-effect regression.

    ``_EMOptFunc`` (``infer.py:128-140``): C <- weighted_sum_covar of a first posterior
    pass.  ``_NoopOptFunc`` (``infer.py:143-159``): C <- that * times + plus.  Then the
    posterior proper is computed under the new C (``infer.py:685-687``)."""
    first = compute_posterior(rTZDinv, inv_shat2_diag, pi, effect_covar_l)
    new_c = first["weighted_sum_covar"]
    if not em:
        new_c = new_c * adj_times + adj_plus
    post = compute_posterior(rTZDinv, inv_shat2_diag, pi, new_c)
    return new_c, post


def _store(post: dict, l: int, rows: dict):
    for name, val in rows.items():
        post[name][l] = val


# --------------------------------------------------------------------------------------
# prior set-up shared by both entry points (reference infer.py:367-479, infer_ss.py:225-337)


def build_effect_covar(k: int, effect_var: Sequence[float], rho: Sequence[float]) -> np.ndarray:
    c0 = np.diag(np.asarray(effect_var, dtype=float))
    ct = 0
    for col in range(k):
        for row in range(1, k):
            if col < row:
                two_sd = np.sqrt(effect_var[row] * effect_var[col])
                c0[row, col] = rho[ct] * two_sd
                c0[col, row] = rho[ct] * two_sd
                ct += 1
    return c0


def build_adjustor(k, c0, no_update, user_effect_var, user_rho):
    if no_update:
        if user_effect_var is None and user_rho is not None and k != 1:
            return np.eye(k), c0 - np.diag(np.diag(c0))
        if user_effect_var is not None and user_rho is None and k != 1:
            return np.ones((k, k)) - np.eye(k), c0 * np.eye(k)
        return np.zeros((k, k)), c0.copy()
    return np.ones((k, k)), np.zeros((k, k))


def _zero_posterior(L, p, k):
    return dict(
        alpha=np.zeros((L, p)),
        post_mean=np.zeros((L, p, k)),
        post_mean_sq=np.zeros((L, p, k, k)),
        weighted_sum_covar=np.zeros((L, k, k)),
        kl=np.zeros(L),
        log_bf=np.zeros((L, p)),
    )


def _copy_state(effect_covar, resid_var, post):
    return effect_covar.copy(), resid_var.copy(), {n: v.copy() for n, v in post.items()}


def _converge_loop(step, L, p, k, effect_covar, resid_var, max_iter, min_tol):
    """Host loop of the reference (``infer.py:496-548`` == ``infer_ss.py:346-395``)."""
    post = _zero_posterior(L, p, k)
    elbo = [-np.inf]
    elbo_increase = True
    n_iter = 0
    for o_iter in range(max_iter):
        prev = _copy_state(effect_covar, resid_var, post)
        effect_covar, resid_var, post, cur = step(effect_covar, resid_var, post)
        n_iter = o_iter + 1
        last = elbo[o_iter]
        elbo.append(cur)
        with np.errstate(invalid="ignore"):
            elbo_increase = bool(cur >= last or np.isclose(cur, last, atol=1e-8))
        if not elbo_increase:
            effect_covar, resid_var, post = prev
            break
        if abs(cur - last) < min_tol:
            break
    return effect_covar, resid_var, post, np.asarray(elbo), elbo_increase, n_iter


def reorder_l(effect_covar, post):
    """Reference ``infer.py:811-832``: descending nuclear norm of weighted_sum_covar."""
    with np.errstate(all="ignore"):
        try:
            norm = np.sum(np.linalg.svd(post["weighted_sum_covar"], compute_uv=False), axis=1)
        except np.linalg.LinAlgError:
            norm = np.full(post["weighted_sum_covar"].shape[0], np.nan)
    order = np.argsort(-norm, kind="stable")
    return effect_covar[order], {n: v[order] for n, v in post.items()}, order


# --------------------------------------------------------------------------------------
# individual level (reference infer.py:162-635)


def prepare_individual(Xs, ys, covar=None, no_scale=False, no_regress=False):
    """Reference ``infer.py:319-340``: covariate residualisation, centring, scaling,
    default residual variance.  Works on copies (the reference mutates its inputs)."""
    Xs = [np.array(x, dtype=float) for x in Xs]
    ys = [np.array(y, dtype=float) for y in ys]
    k = len(Xs)
    if covar is not None:
        for i in range(k):
            Xs[i], ys[i] = regress_covar(Xs[i], ys[i], np.asarray(covar[i], dtype=float), no_regress)
    for i in range(k):
        Xs[i] = Xs[i] - np.mean(Xs[i], axis=0)
        ys[i] = ys[i] - np.mean(ys[i])
        if not no_scale:
            Xs[i] = Xs[i] / np.std(Xs[i], axis=0)
            ys[i] = ys[i] / np.std(ys[i])
        ys[i] = np.squeeze(ys[i])
    return Xs, ys


def _erss(Xs, ys, beta, beta_sq):
    """Reference ``infer.py:799-808``.  beta, beta_sq: [k, p, L]."""
    out = np.zeros(len(Xs))
    for i, (X, y) in enumerate(zip(Xs, ys)):
        mu = X @ beta[i]  # [n, L]
        mu2 = (X ** 2) @ beta_sq[i]
        out[i] = np.sum((y - np.sum(mu, axis=1)) ** 2) + np.sum(mu2 - mu ** 2)
    return out


def update_effects(Xs, ys, XtXs, ns, pi, effect_covar, resid_var, post, adj_times, adj_plus, em):
    """One outer IBSS iteration.  Reference ``infer.py:590-689``."""
    effect_covar = effect_covar.copy()
    post = {n: v.copy() for n, v in post.items()}
    L, p, k = post["post_mean"].shape
    lsum = np.sum(post["post_mean"], axis=0)  # [p, k]
    residual = [ys[i] - Xs[i] @ lsum[:, i] for i in range(k)]
    for l in range(L):
        residual_l = [residual[i] + Xs[i] @ post["post_mean"][l][:, i] for i in range(k)]
        Xty = np.stack([Xs[i].T @ residual_l[i] for i in range(k)])  # [k, p]
        rTZDinv = (Xty / resid_var).T
        inv_shat2_diag = (XtXs / resid_var).T
        new_c, rows = _ser(rTZDinv, inv_shat2_diag, pi, effect_covar[l], adj_times, adj_plus, em)
        effect_covar[l] = new_c
        _store(post, l, rows)
        residual = [residual_l[i] - Xs[i] @ post["post_mean"][l][:, i] for i in range(k)]
    beta = post["post_mean"].T  # [k, p, L]
    beta_sq = np.diagonal(post["post_mean_sq"], axis1=2, axis2=3).T  # [k, p, L]
    erss = _erss(Xs, ys, beta, beta_sq)
    with np.errstate(all="ignore"):
        exp_ll = np.sum(
            -(0.5 * ns) * np.log(2 * np.pi * resid_var) - (0.5 / resid_var) * erss[:, None]
        )
    new_resid_var = erss[:, None] / ns
    elbo = exp_ll - np.sum(post["kl"])
    return effect_covar, new_resid_var, post, float(elbo)


def fit_individual(
    Xs,
    ys,
    covar=None,
    L=10,
    no_scale=False,
    no_regress=False,
    no_update=False,
    pi=None,
    resid_var=None,
    effect_var=None,
    rho=None,
    max_iter=500,
    min_tol=1e-4,
    no_reorder=False,
) -> OracleFit:
    """Reference ``infer_sushie`` minus validation and ``make_cs`` (``infer.py:300-556``)."""
    k = len(Xs)
    Xs, ys = prepare_individual(Xs, ys, covar, no_scale, no_regress)
    p = Xs[0].shape[1]
    if pi is None:
        pi = np.ones(p) / float(p)
    else:
        pi = np.asarray(pi, dtype=float)
        if np.sum(pi) > 1:  # infer.py:313-317
            pi = pi / np.sum(pi)
    if resid_var is None:
        resid_var = [np.var(y, ddof=1) for y in ys]
    resid_var = np.asarray([float(v) for v in resid_var])[:, None]
    user_effect_var, user_rho = effect_var, rho
    if effect_var is None:
        effect_var = [1e-3] * k
    if rho is None:
        rho = [0.1] * math.comb(k, 2)
    c0 = build_effect_covar(k, [float(v) for v in effect_var], [float(v) for v in rho])
    adj_times, adj_plus = build_adjustor(k, c0, no_update, user_effect_var, user_rho)
    effect_covar = np.stack([c0] * L)
    ns = np.asarray([X.shape[0] for X in Xs], dtype=float)[:, None]
    XtXs = np.stack([np.sum(X ** 2, axis=0) for X in Xs])  # [k, p]

    def step(ec, rv, post):
        return update_effects(Xs, ys, XtXs, ns, pi, ec, rv, post, adj_times, adj_plus, not no_update)

    effect_covar, resid_var, post, elbo, inc, n_iter = _converge_loop(
        step, L, p, k, effect_covar, resid_var, max_iter, min_tol
    )
    order = np.arange(L)
    if not no_reorder:
        effect_covar, post, order = reorder_l(effect_covar, post)
    return OracleFit(
        OraclePrior(pi, resid_var, effect_covar),
        OraclePosterior(**post),
        ns.astype(int),
        elbo,
        inc,
        order,
        n_iter,
        Xs,
        None,
    )


# --------------------------------------------------------------------------------------
# summary level (reference infer_ss.py:26-575)


def _erss_ss(Xtys, XtXs, ns, beta, beta_sq):
    """Reference ``infer_ss.py:555-575``.  beta, beta_sq: [k, p, L]."""
    ebeta = np.sum(beta, axis=2)
    term_1 = np.squeeze(ns, axis=1)
    term_2 = -2 * np.sum(ebeta * Xtys, axis=1)
    term_3 = np.einsum("kp,kpq,kq->k", ebeta, XtXs, ebeta)
    tr_beta = np.transpose(beta, (0, 2, 1))
    tr_beta_sq = np.transpose(beta_sq, (0, 2, 1))
    term_4 = -1 * np.einsum("klp,kpq,klq->k", tr_beta, XtXs, tr_beta)
    term_5 = np.einsum("kp,klp->k", np.diagonal(XtXs, axis1=1, axis2=2), tr_beta_sq)
    return term_1 + term_2 + term_3 + term_4 + term_5


def update_effects_ss(Xtys, XtXs, ns, pi, effect_covar, resid_var, post, adj_times, adj_plus, em):
    """One outer iteration on summary statistics.  Reference ``infer_ss.py:437-537``."""
    effect_covar = effect_covar.copy()
    post = {n: v.copy() for n, v in post.items()}
    L, p, k = post["post_mean"].shape
    lsum = np.sum(post["post_mean"], axis=0)
    residual = Xtys - np.einsum("kpq,kq->kp", XtXs, lsum.T)
    diag = np.diagonal(XtXs, axis1=1, axis2=2)
    for l in range(L):
        residual_l = residual + np.einsum("kpq,kq->kp", XtXs, post["post_mean"][l].T)
        rTZDinv = (residual_l / resid_var).T
        inv_shat2_diag = (diag / resid_var).T
        new_c, rows = _ser(rTZDinv, inv_shat2_diag, pi, effect_covar[l], adj_times, adj_plus, em)
        effect_covar[l] = new_c
        _store(post, l, rows)
        residual = residual_l - np.einsum("kpq,kq->kp", XtXs, post["post_mean"][l].T)
    beta = post["post_mean"].T
    beta_sq = np.diagonal(post["post_mean_sq"], axis1=2, axis2=3).T
    erss = _erss_ss(Xtys, XtXs, ns, beta, beta_sq)
    with np.errstate(all="ignore"):
        exp_ll = np.sum(
            -(0.5 * ns) * np.log(2 * np.pi * resid_var) - (0.5 / resid_var) * erss[:, None]
        )
    new_resid_var = erss[:, None] / ns
    elbo = exp_ll - np.sum(post["kl"])
    return effect_covar, new_resid_var, post, float(elbo)


def fit_summary(
    lds,
    ns,
    zs,
    L=10,
    no_update=False,
    pi=None,
    resid_var=None,
    effect_var=None,
    rho=None,
    max_iter=500,
    min_tol=1e-4,
    no_reorder=False,
) -> OracleFit:
    """Reference ``infer_sushie_ss`` minus validation and ``make_cs`` (``infer_ss.py:176-402``)."""
    ns = np.asarray(ns, dtype=float).reshape(-1, 1)
    k = ns.shape[0]
    lds = np.asarray(lds, dtype=float)
    zs = np.asarray(zs, dtype=float)
    p = lds.shape[1]
    if pi is None:
        pi = np.ones(p) / float(p)
    else:
        pi = np.asarray(pi, dtype=float)
        if np.sum(pi) != 1:  # infer_ss.py:189-193
            pi = pi / np.sum(pi)
    if resid_var is None:
        resid_var = [1] * k
    resid_var = np.asarray([float(v) for v in resid_var])[:, None]
    user_effect_var, user_rho = effect_var, rho
    if effect_var is None:
        effect_var = [1e-3] * k
    if rho is None:
        rho = [0.1] * math.comb(k, 2)
    c0 = build_effect_covar(k, [float(v) for v in effect_var], [float(v) for v in rho])
    adj_times, adj_plus = build_adjustor(k, c0, no_update, user_effect_var, user_rho)
    effect_covar = np.stack([c0] * L)
    sigma2 = ns / (ns + zs ** 2)  # infer_ss.py:342
    Xtys = np.sqrt(ns) * np.sqrt(sigma2) * zs  # infer_ss.py:343
    XtXs = ns[:, :, None] * lds  # infer_ss.py:344

    def step(ec, rv, post):
        return update_effects_ss(Xtys, XtXs, ns, pi, ec, rv, post, adj_times, adj_plus, not no_update)

    effect_covar, resid_var, post, elbo, inc, n_iter = _converge_loop(
        step, L, p, k, effect_covar, resid_var, max_iter, min_tol
    )
    order = np.arange(L)
    if not no_reorder:
        effect_covar, post, order = reorder_l(effect_covar, post)
    return OracleFit(
        OraclePrior(pi, resid_var, effect_covar),
        OraclePosterior(**post),
        ns.astype(int),
        elbo,
        inc,
        order,
        n_iter,
        None,
        lds,
    )
