"""The oracle and the committed goldens against the REFERENCE'S OWN SOURCE (CPU).

``/root/reference/sushie/{infer,infer_ss,utils,cli}.py`` are imported unmodified and executed through
``tests/refshim`` (numpy-backed stand-ins for jax / equinox, see its docstring).  These tests run wherever the
reference tree exists (the build container); on a machine without it only the golden-based checks at the bottom run.
"""
import os
import types

import numpy as np
import pytest

import refshim
from oracle import credible_sets, ibss

needs_reference = pytest.mark.skipif(not refshim.reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    import warnings

    warnings.filterwarnings("ignore")
    return refshim.load_reference(("log", "utils", "infer", "infer_ss", "io", "cli"))


def _locus(seed, k, n, p, causal=(3, 40), h=0.9):
    rng = np.random.default_rng(seed)
    maf = rng.uniform(0.1, 0.5, size=p)
    Xs, ys = [], []
    beta = np.zeros(p)
    beta[list(causal)] = rng.choice([-1.0, 1.0], size=len(causal)) * h
    for a in range(k):
        g = rng.binomial(2, maf, size=(n[a], p)).astype(float)
        for j in np.nonzero(g.std(axis=0) == 0)[0]:
            g[0, j] = 1.0 if g[0, j] != 1.0 else 0.0
        ys.append(((g - g.mean(0)) / g.std(0)) @ beta + rng.standard_normal(n[a]))
        Xs.append(g)
    return Xs, ys


def _same_fit(res, fit, cs_out=None):
    """Reference SushieResult vs OracleFit: rounding-level agreement of everything the loop returns."""
    assert len(res.elbo) - 1 == fit.n_iter
    assert bool(res.elbo_increase) == bool(fit.elbo_increase)
    np.testing.assert_allclose(np.asarray(res.elbo, float)[1:], fit.elbo[1:], rtol=1e-10, equal_nan=True)
    assert np.array_equal(np.asarray(res.l_order), fit.l_order)
    for name in ("alpha", "post_mean", "post_mean_sq", "weighted_sum_covar", "kl", "log_bf"):
        np.testing.assert_allclose(np.asarray(getattr(res.posteriors, name)), getattr(fit.posteriors, name),
                                   rtol=1e-7, atol=1e-10, err_msg=name)
    np.testing.assert_allclose(np.asarray(res.priors.effect_covar), fit.priors.effect_covar, rtol=1e-7, atol=1e-14)
    np.testing.assert_allclose(np.asarray(res.priors.resid_var), fit.priors.resid_var, rtol=1e-10)
    if cs_out is not None:
        cs, alphas, pip_all, pip_cs = cs_out
        assert np.array_equal(res.cs.SNPIndex.values.astype(int), cs.SNPIndex.values.astype(int))
        assert np.array_equal(res.cs.CSIndex.values.astype(int), cs.CSIndex.values.astype(int))
        assert list(res.alphas.columns) == list(alphas.columns)
        np.testing.assert_allclose(res.alphas.values.astype(float), alphas.values.astype(float), atol=1e-9)
        np.testing.assert_allclose(np.asarray(res.pip_all), pip_all, atol=1e-12)
        np.testing.assert_allclose(np.asarray(res.pip_cs), pip_cs, atol=1e-12)


IND_CASES = [
    (11, 1, (130,), 140, dict(L=3)),
    (12, 2, (100, 120), 150, dict(L=4)),
    (13, 3, (80, 90, 70), 128, dict(L=3, min_tol=1e-3)),
    (14, 2, (100, 120), 150, dict(L=3, no_update=True)),
    (15, 2, (100, 120), 150, dict(L=3, no_update=True, rho=[0.4])),
    (16, 2, (100, 120), 150, dict(L=3, no_update=True, effect_var=[0.02, 0.01])),
    (17, 2, (100, 120), 150, dict(L=2, no_scale=True, max_iter=40)),
    (18, 3, (60, 70, 80), 110, dict(L=2, resid_var=[0.9, 1.1, 1.0], max_iter=7)),
    (19, 2, (90, 110), 120, dict(L=3, pi="random")),
    (20, 4, (50, 60, 70, 80), 105, dict(L=2)),
]


@needs_reference
@pytest.mark.parametrize("seed,k,n,p,kw", IND_CASES)
def test_reference_source_vs_oracle_individual(ref, seed, k, n, p, kw):
    kw = dict(kw)
    Xs, ys = _locus(seed, k, n, p)
    if kw.get("pi") == "random":
        pi = np.random.default_rng(5).uniform(0.5, 2.0, size=p)
        kw["pi"] = pi / pi.sum()
    res = ref.infer.infer_sushie([x.copy() for x in Xs], [y.copy() for y in ys], **kw)
    fit = ibss.fit_individual(Xs, ys, **kw)
    cs_out = credible_sets.make_cs(fit.posteriors.alpha, fit.posteriors.log_bf, fit.sample_size, Xs=fit.Xs,
                                   threshold=0.95, purity=0.5, purity_method="weighted", max_select=250, seed=12345)
    _same_fit(res, fit, cs_out)


@needs_reference
@pytest.mark.parametrize("no_regress", [False, True])
def test_reference_source_vs_oracle_covariates(ref, no_regress):
    Xs, ys = _locus(31, 2, (100, 120), 110)
    rng = np.random.default_rng(2)
    covar = [np.column_stack([rng.standard_normal(x.shape[0]), rng.integers(0, 2, x.shape[0])]).astype(float) for x in Xs]
    res = ref.infer.infer_sushie([x.copy() for x in Xs], [y.copy() for y in ys], [c.copy() for c in covar], L=3,
                                 no_regress=no_regress)
    fit = ibss.fit_individual(Xs, ys, covar=covar, L=3, no_regress=no_regress)
    _same_fit(res, fit)


@needs_reference
@pytest.mark.parametrize("kw", [dict(L=3), dict(L=2, no_update=True, effect_var=[0.02, 0.01]), dict(L=1),
                                dict(L=4, max_iter=2), dict(L=3, min_tol=1e-3)])
def test_reference_source_vs_oracle_summary(ref, kw):
    Xs, ys = _locus(21, 2, (150, 170), 140)
    lds, zs = [], []
    for X, y in zip(Xs, ys):
        Xstd = (X - X.mean(0)) / X.std(0)
        lds.append(Xstd.T @ Xstd / X.shape[0])
        zs.append(Xstd.T @ ((y - y.mean()) / y.std()) / np.sqrt(X.shape[0]))
    lds, zs, ns = np.stack(lds), np.stack(zs), np.array([[150], [170]])
    res = ref.infer_ss.infer_sushie_ss(lds.copy(), ns.copy(), zs.copy(), **kw)
    fit = ibss.fit_summary(lds, ns, zs, **kw)
    cs_out = credible_sets.make_cs(fit.posteriors.alpha, fit.posteriors.log_bf, fit.sample_size, lds=fit.lds,
                                   threshold=0.95, purity=0.5, purity_method="weighted", max_select=250, seed=12345)
    _same_fit(res, fit, cs_out)


@needs_reference
@pytest.mark.parametrize("method,max_select", [("weighted", 500), ("max", 40), ("min", 25)])
def test_reference_make_cs_vs_oracle_and_product_selection(ref, golden_syn, method, max_select):
    """``make_cs`` of the reference (pandas sort / merge rounds, sub-sampled purity) vs the oracle's restatement and
    the product's host-side set selection (``sushie_b200.cs.select_sets``)."""
    from sushie_b200 import cs as product_cs

    g = golden_syn
    Xs = [g[f"k3_X{i}"].astype(float) for i in range(3)]
    ys = [g[f"k3_y{i}"] for i in range(3)]
    Xstd, _ = ibss.prepare_individual(Xs, ys)
    n_max = max(x.shape[0] for x in Xstd)
    padded = np.stack([np.pad(x, ((0, n_max - x.shape[0]), (0, 0))) for x in Xstd])
    alpha, log_bf = g["k3_alpha"], g["k3_log_bf"]
    ns = np.array([[x.shape[0]] for x in Xs])
    got = ref.infer.make_cs(alpha, log_bf, ns, Xs=padded, threshold=0.9, purity=0.5, purity_method=method,
                            max_select=max_select, seed=12345)
    want = credible_sets.make_cs(alpha, log_bf, ns, Xs=Xstd, threshold=0.9, purity=0.5, purity_method=method,
                                 max_select=max_select, seed=12345)
    assert list(got[1].columns) == list(want[1].columns)
    np.testing.assert_allclose(got[1].values.astype(float), want[1].values.astype(float), atol=1e-9)
    assert np.array_equal(got[0].SNPIndex.values.astype(int), want[0].SNPIndex.values.astype(int))
    assert np.array_equal(got[0].CSIndex.values.astype(int), want[0].CSIndex.values.astype(int))
    # product: same members, same sub-sample
    orders, csums, sizes, purity_idx = product_cs.select_sets(alpha, 0.9, max_select, 12345)
    for ldx in range(alpha.shape[0]):
        in_cs = got[1][f"in_cs_l{ldx + 1}"].values.astype(int)
        assert set(np.nonzero(in_cs)[0].tolist()) == set(orders[ldx][: sizes[ldx]].tolist())
        assert len(purity_idx[ldx]) == min(sizes[ldx], max_select)


@needs_reference
def test_committed_goldens_are_what_the_reference_source_gives(ref, golden_syn):
    """Guards against stale fixtures: re-run two of the synthetic cases and compare with the committed file."""
    for tag, k, kw in (("k2", 2, dict(L=5)), ("k2rho", 2, dict(L=4, no_update=True, rho=[0.5]))):
        Xs = [golden_syn[f"{tag}_X{i}"].astype(float) for i in range(k)]
        ys = [golden_syn[f"{tag}_y{i}"].copy() for i in range(k)]
        res = ref.infer.infer_sushie(Xs, ys, min_tol=1e-4, **kw)
        assert len(res.elbo) - 1 == int(golden_syn[f"{tag}_n_iter"])
        np.testing.assert_allclose(np.asarray(res.elbo, float), golden_syn[f"{tag}_elbo"], rtol=1e-12, equal_nan=True)
        np.testing.assert_allclose(np.asarray(res.posteriors.alpha), golden_syn[f"{tag}_alpha"], rtol=1e-9, atol=1e-300)
        assert np.array_equal(res.cs.SNPIndex.values.astype(int), golden_syn[f"{tag}_cs_snp"])


@needs_reference
def test_prepare_cv_and_scores_match_reference_source(ref):
    """The product's ``_prepare_cv`` / ``_cv_scores`` vs ``cli._prepare_cv`` / ``utils.ols`` of the reference."""
    from sushie_b200 import cli

    Xs, ys = _locus(41, 2, (103, 97), 20, causal=(3, 11))
    theirs = ref.cli._prepare_cv([x.copy() for x in Xs], [y.copy() for y in ys], 5, 777)
    ours = cli._prepare_cv([x.copy() for x in Xs], [y.copy() for y in ys], 5, 777)
    for a, b in zip(theirs, ours):
        for i in range(2):
            np.testing.assert_array_equal(np.asarray(a.train_geno[i]), b.train_geno[i])
            np.testing.assert_array_equal(np.asarray(a.valid_geno[i]), b.valid_geno[i])
            np.testing.assert_allclose(np.asarray(a.train_pheno[i]), b.train_pheno[i], atol=1e-14)
            np.testing.assert_allclose(np.asarray(a.valid_pheno[i]), b.valid_pheno[i], atol=1e-14)
    rng = np.random.default_rng(3)
    fits = [types.SimpleNamespace(posteriors=types.SimpleNamespace(post_mean=rng.standard_normal((3, 20, 2)) * 0.1))
            for _ in range(5)]
    want = []
    for idx in range(2):
        est = np.concatenate([np.asarray(f.valid_geno[idx]) @ fit.posteriors.post_mean.sum(0)[:, idx] for f, fit in zip(theirs, fits)])
        obs = np.concatenate([np.asarray(f.valid_pheno[idx]) for f in theirs])
        _, adj_r2, p_value = ref.utils.ols(est[:, None], obs[:, None])
        want.append([float(adj_r2[0]), float(p_value[1][0])])
    np.testing.assert_allclose(np.asarray(cli._cv_scores(ours, fits)), np.asarray(want), rtol=1e-9)


# ---- checks that need no reference tree (committed goldens only) -----------------------------------------


def test_oracle_follows_the_headline_golden(golden_headline):
    """Config 3 at full shape: the oracle's first 10 iterations on gene 1 reproduce the ELBO trace the reference's
    source produced (the converged fits themselves are compared on the GPU, tests/test_gpu_parity.py)."""
    import bench

    Xs, ys = bench.make_gene(1)
    fit = ibss.fit_individual([x.astype(float) for x in Xs], ys, L=bench.L_EFF, max_iter=10, min_tol=0.0)
    np.testing.assert_allclose(fit.elbo[1:], golden_headline["g1_tol4_elbo"][1:11], rtol=1e-11)
    # the CLI tolerance stops on the same trajectory, earlier
    n3 = int(golden_headline["g1_tol3_n_iter"])
    np.testing.assert_array_equal(golden_headline["g1_tol3_elbo"], golden_headline["g1_tol4_elbo"][: n3 + 1])


def test_product_fold_split_matches_config5_golden(golden_config5):
    """Config 5: the product's threefry row shuffles / fold cuts / rank-normalised phenotypes are the ones the
    reference's ``cli._prepare_cv`` produced for the same gene and seed (host logic, no GPU)."""
    import bench
    import bench_cv
    from sushie_b200 import cli

    g = golden_config5
    gene, cv_num, seed = int(g["gene"]), int(g["cv_num"]), int(g["seed"])
    Xs, ys = bench.make_gene(gene, bench.N_IND, bench_cv.gene_width(gene, 500, 4000))
    rows = [np.arange(x.shape[0])[:, None] for x in Xs]
    folds = cli._prepare_cv(rows, list(ys), cv_num, seed)
    for j, fold in enumerate(folds):
        for a in range(len(Xs)):
            np.testing.assert_array_equal(fold.train_geno[a][:, 0], g[f"fold{j}_train_rows{a}"])
            np.testing.assert_array_equal(fold.valid_geno[a][:, 0], g[f"fold{j}_valid_rows{a}"])
            np.testing.assert_allclose(fold.train_pheno[a], g[f"fold{j}_train_pheno{a}"], rtol=0, atol=1e-12)
            np.testing.assert_allclose(fold.valid_pheno[a], g[f"fold{j}_valid_pheno{a}"], rtol=0, atol=1e-12)
