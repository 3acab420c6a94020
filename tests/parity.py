"""Shared parity assertions: product result vs golden (reference source run through tests/refshim).

Tolerances (BASELINE.json north_star / SURVEY.md 8c), fp64 vs fp64:
  same n_iter; max |dELBO| <= 1e-6 * |ELBO|; max |dPIP| <= 1e-4; identical credible-set
  membership; prior covariance within 1e-4 relative (Frobenius) per effect.

Every per-effect quantity is compared effect by effect.  The reference sorts the effects by the nuclear norm of
``weighted_sum_covar`` (``infer.py:811-832``); null effects have norms that agree to rounding, so their relative
order is not determined by the arithmetic.  Both sides are therefore mapped back to the un-sorted effect index
through ``l_order`` before comparing, and a different ``l_order`` is accepted only between effects whose norms are
tied (relative difference <= ``ORDER_TIE_RTOL``) and that hold no kept credible set.
"""
import numpy as np

PIP_TOL = 1e-4
ELBO_RTOL = 1e-6
COV_RTOL = 1e-4
ORDER_TIE_RTOL = 1e-6


def _sets_by_effect(cs_index, cs_snp, l_order):
    out = {}
    for ci, si in zip(cs_index, cs_snp):
        out.setdefault(int(l_order[int(ci) - 1]), set()).add(int(si))
    return out


def assert_fit_matches(res, g, prefix, check_order=True):
    """Returns True when the effect order equals the golden's.  ``check_order=False`` only skips the tie proof
    (used for storage modes whose rounding differs from the golden's by more than the tie band)."""
    n_iter = int(g[f"{prefix}_n_iter"])
    assert len(res.elbo) - 1 == n_iter, f"n_iter {len(res.elbo) - 1} != reference {n_iter}"
    assert bool(res.elbo_increase) == bool(g[f"{prefix}_elbo_increase"])
    e_ref = g[f"{prefix}_elbo"]
    assert np.isneginf(res.elbo[0])
    fin = np.isfinite(e_ref[1:])
    np.testing.assert_allclose(np.asarray(res.elbo)[1:][fin], e_ref[1:][fin], rtol=ELBO_RTOL, atol=0)
    # PIPs are order independent
    assert np.max(np.abs(res.pip_all - g[f"{prefix}_pip_all"])) <= PIP_TOL
    assert np.max(np.abs(res.pip_cs - g[f"{prefix}_pip_cs"])) <= PIP_TOL
    np.testing.assert_allclose(np.ravel(res.priors.resid_var), np.ravel(g[f"{prefix}_resid_var"]), rtol=1e-8)

    lo_got, lo_ref = np.asarray(res.l_order), np.asarray(g[f"{prefix}_l_order"])
    assert sorted(lo_got.tolist()) == sorted(lo_ref.tolist()) == list(range(len(lo_ref)))
    inv_got, inv_ref = np.argsort(lo_got), np.argsort(lo_ref)  # sorted position of un-sorted effect e

    # credible sets: identical membership per kept set, keyed by the un-sorted effect index
    ref_sets = _sets_by_effect(g[f"{prefix}_cs_index"], g[f"{prefix}_cs_snp"], lo_ref)
    got_sets = _sets_by_effect(res.cs.CSIndex.values, res.cs.SNPIndex.values, lo_got)
    assert got_sets == ref_sets, "credible-set membership differs"

    same_order = np.array_equal(lo_got, lo_ref)
    if not same_order and check_order:
        # only tied (null) effects may trade places, and none of them may hold a kept credible set
        with np.errstate(all="ignore"):
            norm = np.linalg.svd(g[f"{prefix}_wsc"], compute_uv=False).sum(axis=1)  # by sorted position (reference)
        moved = np.nonzero(lo_got != lo_ref)[0]
        span = norm[moved]
        assert np.all(np.isfinite(span)) and (span.max() - span.min()) <= ORDER_TIE_RTOL * abs(span.max()), (
            f"effect order differs outside a tie: l_order {lo_got} vs {lo_ref}, norms {norm}")
        assert not (set(lo_ref[moved].tolist()) & set(ref_sets)), "effects holding a kept credible set changed order"
    if not same_order:
        # sorted outputs of tied effects: positions differ, the multiset may not
        kept_pos_got = sorted(int(c) for c in set(res.cs.CSIndex.values.tolist()))
        kept_pos_ref = sorted(int(c) for c in set(np.asarray(g[f"{prefix}_cs_index"]).tolist()))
        assert kept_pos_got == kept_pos_ref, "CSIndex labels of the kept sets differ"

    def by_effect(sorted_arr, inv):
        return np.asarray(sorted_arr)[inv]

    ref_c = by_effect(g[f"{prefix}_effect_covar"], inv_ref)
    got_c = by_effect(res.priors.effect_covar, inv_got)
    for e in range(ref_c.shape[0]):
        num = np.linalg.norm(got_c[e] - ref_c[e])
        assert num <= COV_RTOL * np.linalg.norm(ref_c[e]) + 1e-300, f"effect_covar of effect {e}"
    np.testing.assert_allclose(by_effect(res.posteriors.alpha, inv_got), by_effect(g[f"{prefix}_alpha"], inv_ref),
                               atol=PIP_TOL)
    np.testing.assert_allclose(by_effect(res.posteriors.weighted_sum_covar, inv_got),
                               by_effect(g[f"{prefix}_wsc"], inv_ref), rtol=1e-4, atol=1e-12)
    np.testing.assert_allclose(by_effect(res.posteriors.kl, inv_got), by_effect(g[f"{prefix}_kl"], inv_ref),
                               rtol=1e-5, atol=1e-7)
    if f"{prefix}_weights" in g:
        np.testing.assert_allclose(np.sum(res.posteriors.post_mean, axis=0), g[f"{prefix}_weights"], atol=1e-6)
    if f"{prefix}_post_mean" in g:
        np.testing.assert_allclose(by_effect(res.posteriors.post_mean, inv_got),
                                   by_effect(g[f"{prefix}_post_mean"], inv_ref), atol=1e-6)
    if f"{prefix}_log_bf" in g:
        np.testing.assert_allclose(by_effect(res.posteriors.log_bf, inv_got), by_effect(g[f"{prefix}_log_bf"], inv_ref),
                                   rtol=1e-6, atol=1e-6)
    purity_cols = [c for c in res.alphas.columns if c.startswith("purity_l")]
    np.testing.assert_allclose(by_effect(res.alphas[purity_cols].iloc[0].values, inv_got),
                               by_effect(g[f"{prefix}_purity"], inv_ref), atol=1e-8)
    kept_cols = [c for c in res.alphas.columns if c.startswith("kept_l")]
    assert np.array_equal(by_effect(res.alphas[kept_cols].iloc[0].values, inv_got), by_effect(g[f"{prefix}_kept"], inv_ref))
    return same_order
