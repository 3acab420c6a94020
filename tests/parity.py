"""Shared parity assertions: product result vs oracle golden.

Tolerances (BASELINE.json north_star / SURVEY.md 8c), fp64 vs fp64:
  same n_iter; max |dELBO| <= 1e-6 * |ELBO|; max |dPIP| <= 1e-4; identical credible-set
  membership; prior covariance within 1e-4 relative (Frobenius) per effect.
"""
import numpy as np

PIP_TOL = 1e-4
ELBO_RTOL = 1e-6
COV_RTOL = 1e-4


def assert_fit_matches(res, g, prefix, check_order=True):
    n_iter = int(g[f"{prefix}_n_iter"])
    assert len(res.elbo) - 1 == n_iter, f"n_iter {len(res.elbo) - 1} != oracle {n_iter}"
    assert bool(res.elbo_increase) == bool(g[f"{prefix}_elbo_increase"])
    e_ref = g[f"{prefix}_elbo"]
    assert np.isneginf(res.elbo[0])
    fin = np.isfinite(e_ref[1:])
    np.testing.assert_allclose(res.elbo[1:][fin], e_ref[1:][fin], rtol=ELBO_RTOL, atol=0)
    # PIPs are order independent
    assert np.max(np.abs(res.pip_all - g[f"{prefix}_pip_all"])) <= PIP_TOL
    assert np.max(np.abs(res.pip_cs - g[f"{prefix}_pip_cs"])) <= PIP_TOL
    # credible sets: identical membership per kept set
    ref_sets = {}
    for ci, si in zip(g[f"{prefix}_cs_index"], g[f"{prefix}_cs_snp"]):
        ref_sets.setdefault(int(ci), set()).add(int(si))
    got_sets = {}
    for ci, si in zip(res.cs.CSIndex.values, res.cs.SNPIndex.values):
        got_sets.setdefault(int(ci), set()).add(int(si))
    assert sorted(map(sorted, got_sets.values())) == sorted(map(sorted, ref_sets.values()))
    np.testing.assert_allclose(res.priors.resid_var, g[f"{prefix}_resid_var"], rtol=1e-8)
    same_order = np.array_equal(res.l_order, g[f"{prefix}_l_order"])
    if check_order and same_order:
        assert got_sets == ref_sets
        ref_c = g[f"{prefix}_effect_covar"]
        for l in range(ref_c.shape[0]):
            num = np.linalg.norm(res.priors.effect_covar[l] - ref_c[l])
            assert num <= COV_RTOL * np.linalg.norm(ref_c[l]) + 1e-300, f"effect_covar[{l}]"
        np.testing.assert_allclose(res.posteriors.alpha, g[f"{prefix}_alpha"], atol=PIP_TOL)
        np.testing.assert_allclose(res.posteriors.weighted_sum_covar, g[f"{prefix}_wsc"], rtol=1e-4, atol=1e-12)
        np.testing.assert_allclose(res.posteriors.kl, g[f"{prefix}_kl"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(res.posteriors.post_mean, g[f"{prefix}_post_mean"], atol=1e-6)
        np.testing.assert_allclose(res.posteriors.log_bf, g[f"{prefix}_log_bf"], rtol=1e-6, atol=1e-6)
        purity_cols = [c for c in res.alphas.columns if c.startswith("purity_l")]
        np.testing.assert_allclose(res.alphas[purity_cols].iloc[0].values, g[f"{prefix}_purity"], atol=1e-8)
        kept_cols = [c for c in res.alphas.columns if c.startswith("kept_l")]
        assert np.array_equal(res.alphas[kept_cols].iloc[0].values, g[f"{prefix}_kept"])
    else:
        # null effects have near-equal norms, so their order is fragile:
        # compare the order-free multiset of effect-covariance traces instead
        tr_ref = np.sort(np.trace(g[f"{prefix}_effect_covar"], axis1=1, axis2=2))
        tr_got = np.sort(np.trace(res.priors.effect_covar, axis1=1, axis2=2))
        np.testing.assert_allclose(tr_got, tr_ref, rtol=COV_RTOL)
    return same_order
