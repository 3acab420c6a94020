"""Oracle self-checks (CPU): the numpy restatement vs its committed goldens, independent
scipy identities, and the external anchors the reference documents."""
import math
import os

import numpy as np
import pytest
from scipy import stats

from oracle import credible_sets, ibss, threefry


def test_threefry_known_answers():
    # Random123 / JAX test-suite vectors for threefry2x32 (20 rounds)
    kat = [
        ((0x0, 0x0), (0x0, 0x0), (0x6B200159, 0x99BA4EFE)),
        ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
        ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0)),
    ]
    for key, ctr, want in kat:
        a, b = threefry.threefry2x32(key, np.array([ctr[0]], np.uint32), np.array([ctr[1]], np.uint32))
        assert (int(a[0]), int(b[0])) == want


def test_jax_split_documented_value():
    # jax.random.split(jax.random.PRNGKey(0)) as printed in the JAX documentation
    got = threefry.split(threefry.prng_key(0))
    assert got.tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]


def test_choice_is_a_subset_permutation():
    vals = np.arange(1000, 3000)
    out = threefry.choice_without_replacement(threefry.prng_key(12345), vals, 250)
    assert len(out) == 250 and len(set(out.tolist())) == 250 and set(out.tolist()) <= set(vals.tolist())
    again = threefry.choice_without_replacement(threefry.prng_key(12345), vals, 250)
    assert np.array_equal(out, again)


def test_posterior_matches_scipy():
    """log_bf = -logpdf(mu; 0, S) and KL(N(mu,S)||N(0,C)) vs scipy / closed forms."""
    rng = np.random.default_rng(0)
    k, p = 3, 7
    A = rng.standard_normal((k, k))
    C = A @ A.T * 1e-2 + 1e-3 * np.eye(k)
    r = rng.standard_normal((p, k)) * 5
    d = rng.uniform(50, 500, size=(p, k))
    pi = np.ones(p) / p
    out = ibss.compute_posterior(r, d, pi, C)
    for j in range(p):
        S = np.linalg.inv(np.diag(d[j]) + np.linalg.inv(C))
        mu = S @ r[j]
        want = -stats.multivariate_normal.logpdf(mu, mean=np.zeros(k), cov=S)
        assert math.isclose(out["log_bf"][j], want, rel_tol=1e-10)
    # alpha sums to one, weighted_sum_covar is symmetric PSD
    assert math.isclose(out["alpha"].sum(), 1.0, rel_tol=1e-12)
    ev = np.linalg.eigvalsh(out["weighted_sum_covar"])
    assert np.all(ev > 0)
    # KL of a Gaussian against itself is zero
    assert abs(ibss.kl_mvn(np.zeros((1, k)), C[None], C)[0]) < 1e-12


def test_make_pip():
    a = np.array([[0.5, 0.0, 1.0], [0.5, 0.1, 0.0]])
    np.testing.assert_allclose(ibss.make_pip(a), [0.75, 0.1, 1.0])


@pytest.mark.parametrize("tag,k,kw", [
    ("k1", 1, dict(L=5)),
    ("k2rho", 2, dict(L=4, no_update=True, rho=[0.5])),
    ("k2var", 2, dict(L=4, no_update=True, effect_var=[0.01, 0.02])),
    ("k2noupd", 2, dict(L=4, no_update=True)),
])
def test_oracle_reproduces_committed_goldens(golden_syn, tag, k, kw):
    g = golden_syn
    Xs = [g[f"{tag}_X{i}"].astype(float) for i in range(k)]
    ys = [g[f"{tag}_y{i}"] for i in range(k)]
    fit = ibss.fit_individual(Xs, ys, min_tol=1e-4, **kw)
    assert fit.n_iter == int(g[f"{tag}_n_iter"])
    np.testing.assert_allclose(fit.elbo[1:], g[f"{tag}_elbo"][1:], rtol=1e-10, equal_nan=True)
    np.testing.assert_allclose(ibss.make_pip(fit.posteriors.alpha), g[f"{tag}_pip_all"], atol=1e-10)
    assert fit.elbo_increase == bool(g[f"{tag}_elbo_increase"])


def test_bundled_summary_golden_and_known_causal_snps(golden_ss):
    """Config 2 (k=2, CLI tolerance): oracle re-run == golden; documented causal SNPs found."""
    g = golden_ss
    fit = ibss.fit_summary(g["k2_lds"], g["k2_ns"], g["k2_zs"], L=10, min_tol=1e-3)
    assert fit.n_iter == int(g["k2_tol3_n_iter"]) == 87
    pip = ibss.make_pip(fit.posteriors.alpha)
    np.testing.assert_allclose(pip, g["k2_tol3_pip_all"], atol=1e-10)
    top = set(g["k2_snp"][np.argsort(-pip)[:2]])
    assert top == {"rs1886340", "rs10914958"}  # docs/manual.rst:88


def test_bundled_individual_golden_anchor(golden_ind):
    g = golden_ind
    assert int(g["tol3_n_iter"]) == 89 and int(g["tol4_n_iter"]) == 289
    top = set(g["snp"][np.argsort(-g["tol4_pip_all"])[:2]])
    assert top == {"rs1886340", "rs10914958"}
    assert g["X0"].shape == (489, 112) and g["X1"].shape == (639, 112)


def test_elbo_monotone_and_invariants(golden_syn):
    g = golden_syn
    for tag in ("k1", "k2", "k3", "k4", "k3null"):
        e = g[f"{tag}_elbo"][1:]
        assert np.all(np.diff(e) >= -1e-8 - 1e-5 * np.abs(e[:-1]))
        np.testing.assert_allclose(g[f"{tag}_alpha"].sum(axis=1), 1.0, atol=1e-10)
        assert np.all((g[f"{tag}_pip_all"] >= 0) & (g[f"{tag}_pip_all"] <= 1))


def test_oracle_make_cs_schema(golden_syn):
    g = golden_syn
    Xs = [g[f"k2_X{i}"].astype(float) for i in range(2)]
    ys = [g[f"k2_y{i}"] for i in range(2)]
    Xstd, _ = ibss.prepare_individual(Xs, ys)
    ns = np.array([[x.shape[0]] for x in Xs])
    cs, alphas, pip_all, pip_cs = credible_sets.make_cs(g["k2_alpha"], g["k2_log_bf"], ns, Xs=Xstd, threshold=0.95,
                                                        purity=0.5, max_select=250)
    assert list(cs.columns) == ["CSIndex", "SNPIndex", "alpha", "c_alpha", "pip_all", "pip_cs"]
    L = g["k2_alpha"].shape[0]
    want = ["SNPIndex"]
    for i in range(1, L + 1):
        want += [f"alpha_l{i}", f"in_cs_l{i}", f"purity_l{i}", f"kept_l{i}", f"log_bf_l{i}"]
    want += ["pip_all", "pip_cs"]
    assert list(alphas.columns) == want
    assert sorted(set(cs.SNPIndex.astype(int))) == sorted(set(g["k2_cs_snp"].tolist()))


@pytest.mark.parametrize("seed,k,n,p,kw", [
    (1, 1, (150,), 120, dict(L=3)),
    (2, 2, (120, 100), 140, dict(L=4)),
    (3, 3, (90, 80, 100), 130, dict(L=3, min_tol=1e-3)),
    (4, 4, (60, 70, 80, 50), 110, dict(L=2)),
    (5, 2, (120, 100), 140, dict(L=3, no_update=True)),
])
def test_fast_form_equals_literal_form(seed, k, n, p, kw):
    """oracle/ibss_fast.py (the timed CPU arm of bench.py: cached X b_l, ERSS without a pass over X, closed-form
    per-SNP posterior) computes what the literal restatement oracle/ibss.py computes."""
    from oracle import ibss_fast

    rng = np.random.default_rng(seed)
    Xs = [rng.binomial(2, rng.uniform(0.1, 0.5, size=p), size=(ni, p)).astype(float) for ni in n]
    for X in Xs:
        X[0, X.std(axis=0) == 0] = 1.0
    beta = np.zeros(p)
    beta[[7, 60]] = [0.8, -0.6]
    ys = [((X - X.mean(0)) / X.std(0)) @ beta + rng.standard_normal(X.shape[0]) for X in Xs]
    fast = ibss_fast.fit_individual(Xs, ys, **kw)
    lit = ibss.fit_individual(Xs, ys, no_reorder=True, **kw)
    assert fast.n_iter == lit.n_iter and fast.elbo_increase == lit.elbo_increase
    np.testing.assert_allclose(fast.elbo[1:], lit.elbo[1:], rtol=1e-9)
    np.testing.assert_allclose(fast.alpha, lit.posteriors.alpha, atol=1e-9)
    np.testing.assert_allclose(fast.post_mean, lit.posteriors.post_mean, atol=1e-9)
    np.testing.assert_allclose(fast.weighted_sum_covar, lit.posteriors.weighted_sum_covar, rtol=1e-7, atol=1e-14)
    np.testing.assert_allclose(fast.kl, lit.posteriors.kl, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(fast.log_bf, lit.posteriors.log_bf, rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(fast.effect_covar, lit.priors.effect_covar, rtol=1e-7, atol=1e-14)
    np.testing.assert_allclose(fast.resid_var, np.ravel(lit.priors.resid_var), rtol=1e-10)
