"""GPU parity for option combinations and edge shapes, against the oracle run in the test itself
(small loci, so the numpy oracle takes well under a second each).

Tolerances as in tests/parity.py: same n_iter, |dELBO| <= 1e-6 |ELBO|, |dPIP| <= 1e-4, prior covariance
within 1e-4 relative (order-free when the effect order is fragile)."""
import numpy as np
import pytest

from conftest import STORAGE_MODES, genotype_storage
from parity import COV_RTOL, ELBO_RTOL, PIP_TOL

pytestmark = pytest.mark.gpu


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
        gs = (g - g.mean(0)) / g.std(0)
        ys.append(gs @ beta + rng.standard_normal(n[a]))
        Xs.append(g)
    return Xs, ys


def _check(res, ref):
    from oracle import ibss

    assert len(res.elbo) - 1 == ref.n_iter
    assert bool(res.elbo_increase) == bool(ref.elbo_increase)
    fin = np.isfinite(ref.elbo[1:])
    np.testing.assert_allclose(res.elbo[1:][fin], np.asarray(ref.elbo)[1:][fin], rtol=ELBO_RTOL)
    assert np.max(np.abs(res.pip_all - ibss.make_pip(ref.posteriors.alpha))) <= PIP_TOL
    np.testing.assert_allclose(np.sum(res.posteriors.alpha, axis=1), 1.0, atol=1e-9)
    np.testing.assert_allclose(np.ravel(res.priors.resid_var), np.ravel(ref.priors.resid_var), rtol=1e-8)
    tr_ref = np.sort(np.trace(ref.priors.effect_covar, axis1=1, axis2=2))
    tr_got = np.sort(np.trace(res.priors.effect_covar, axis1=1, axis2=2))
    np.testing.assert_allclose(tr_got, tr_ref, rtol=COV_RTOL)


CASES = {
    "pi_vector": dict(k=2, n=(90, 110), p=120, kw=dict(L=3), pi=True),
    "resid_var": dict(k=2, n=(90, 110), p=120, kw=dict(L=3, resid_var=[0.8, 1.3])),
    "single_effect": dict(k=3, n=(70, 80, 60), p=101, kw=dict(L=1)),
    "not_converged": dict(k=2, n=(90, 110), p=120, kw=dict(L=4, max_iter=3)),
    "odd_p_k1": dict(k=1, n=(257,), p=333, kw=dict(L=2)),
    "many_effects": dict(k=2, n=(64, 64), p=128, kw=dict(L=12)),
    "no_update_rho": dict(k=3, n=(70, 80, 60), p=150, kw=dict(L=3, no_update=True, rho=[0.3, 0.2, 0.1])),
    "wide_dense": dict(k=2, n=(60, 50), p=2300, kw=dict(L=2)),
    "very_wide": dict(k=2, n=(60, 50), p=5000, kw=dict(L=2, max_iter=6)),   # 12 x 1 fp64 tile / 32 x 2 int8 tile
    "narrow": dict(k=3, n=(120, 90, 70), p=700, kw=dict(L=3)),              # 4-byte int8 vectors
    "k5": dict(k=5, n=(60, 70, 50, 80, 65), p=140, kw=dict(L=3)),           # more than four ancestries (infer.py:241-246
    "k6": dict(k=6, n=(50, 60, 45, 70, 55, 40), p=130, kw=dict(L=2)),       #  has no limit): kernels per count up to 6
}


@pytest.mark.parametrize("storage", STORAGE_MODES)
@pytest.mark.parametrize("name", sorted(CASES))
def test_individual_options_match_oracle(engine_lib, name, storage):
    import sushie_b200
    from oracle import ibss
    from sushie_b200 import config

    case = CASES[name]
    Xs, ys = _locus(100 + sorted(CASES).index(name), case["k"], case["n"], case["p"])
    kw = dict(case["kw"], min_tol=1e-4)
    if case.get("pi"):
        pi = np.random.default_rng(5).uniform(0.5, 2.0, size=case["p"])
        kw["pi"] = pi / pi.sum()
    with genotype_storage(storage):
        res = sushie_b200.infer_sushie([x.copy() for x in Xs], [y.copy() for y in ys], min_snps=50, **kw)
    ref = ibss.fit_individual(Xs, ys, **kw)
    _check(res, ref)


@pytest.mark.parametrize("storage", STORAGE_MODES)
def test_covariates_with_and_without_genotype_regression(engine_lib, storage):
    """Covariates (reference `utils.regress_covar`, infer.py:319-335).  With the default storage the hard calls stay
    bit planes and (I - Q Q^T) is applied on the device as a low-rank correction; "dense" / "bytes" residualise on the
    host and store the dense matrix."""
    import sushie_b200
    from oracle import ibss
    from sushie_b200 import infer

    Xs, ys = _locus(11, 2, (100, 120), 110)
    rng = np.random.default_rng(2)
    covar = [np.column_stack([rng.standard_normal(x.shape[0]), rng.integers(0, 2, x.shape[0]),
                              rng.standard_normal(x.shape[0])]).astype(float) for x in Xs]
    for no_regress in (False, True):
        for kw in (dict(L=3), dict(L=2, no_scale=True)):
            with genotype_storage(storage):
                prep = infer._prepare_kw([x.copy() for x in Xs], [y.copy() for y in ys], covar, no_regress=no_regress,
                                         min_tol=1e-4, **kw)
                assert (prep.problem.covar_q is not None) == (storage == "auto" and not no_regress)
                res = sushie_b200.infer_sushie([x.copy() for x in Xs], [y.copy() for y in ys], covar,
                                               no_regress=no_regress, min_tol=1e-4, **kw)
            ref = ibss.fit_individual(Xs, ys, covar=covar, no_regress=no_regress, min_tol=1e-4, **kw)
            _check(res, ref)


@pytest.mark.parametrize("kw", [dict(L=3), dict(L=2, no_update=True, effect_var=[0.02, 0.01]), dict(L=1),
                                dict(L=3, pi=True), dict(L=4, max_iter=2)])
def test_summary_options_match_oracle(engine_lib, kw):
    import sushie_b200
    from oracle import ibss

    kw = dict(kw)
    Xs, ys = _locus(21, 2, (150, 170), 140)
    lds, zs = [], []
    for X, y in zip(Xs, ys):
        Xs_ = (X - X.mean(0)) / X.std(0)
        lds.append(Xs_.T @ Xs_ / X.shape[0])
        zs.append(Xs_.T @ ((y - y.mean()) / y.std()) / np.sqrt(X.shape[0]))
    ns = np.array([[150], [170]])
    if kw.pop("pi", False):
        pi = np.random.default_rng(6).uniform(0.5, 2.0, size=140)
        kw["pi"] = pi / pi.sum()
    res = sushie_b200.infer_sushie_ss(np.stack(lds), ns, np.stack(zs), min_tol=1e-4, **kw)
    ref = ibss.fit_summary(np.stack(lds), ns, np.stack(zs), min_tol=1e-4, **kw)
    _check(res, ref)


def test_summary_level_very_wide_locus_fits(engine_lib):
    """p = 13 000 with fp32 LD storage uses the 16-vector register tile; a few iterations vs the oracle on a
    banded LD (cheap to build) to make sure the wide geometry computes the same thing."""
    import sushie_b200
    from oracle import ibss
    from sushie_b200 import config

    p = 13000
    idx = np.arange(p)
    rng = np.random.default_rng(3)
    lds, zs = [], []
    for rho in (0.6, 0.4):
        d = np.abs(idx[:, None] - idx[None, :])
        ld = np.where(d < 12, rho ** d, 0.0).astype(np.float32)
        z = rng.standard_normal(p)
        z[[500, 9000]] += (7.0, -6.0)
        lds.append(ld)
        zs.append(z)
    ns = np.array([[5000], [4000]])
    old = config.precision
    config.precision = 32
    try:
        res = sushie_b200.infer_sushie_ss(lds, ns, zs, L=2, max_iter=3, min_tol=0.0)
    finally:
        config.precision = old
    ref = ibss.fit_summary([ld.astype(np.float64) for ld in lds], ns, zs, L=2, max_iter=3, min_tol=0.0)
    np.testing.assert_allclose(res.elbo[1:], np.asarray(ref.elbo)[1:], rtol=1e-6)
    assert np.max(np.abs(res.pip_all - ibss.make_pip(ref.posteriors.alpha))) <= PIP_TOL


def test_loci_too_wide_for_shared_memory_vectors(engine_lib):
    """p = 9000 at fp64: neither the register tiles nor the shared-memory generic path fit, so b and the X^T r
    accumulator live in global memory (individual level) / b is read from the effect vector (summary level)."""
    import sushie_b200
    from oracle import ibss
    from sushie_b200 import config

    p = 9000
    rng = np.random.default_rng(8)
    X = rng.standard_normal((40, p))          # not hard calls: dense storage
    beta = np.zeros(p)
    beta[[100, 7000]] = (1.0, -1.0)
    y = X @ beta + rng.standard_normal(40)
    old = config.team_size
    try:
        for team in (1, 4):
            config.team_size = team
            res = sushie_b200.infer_sushie([X.copy()], [y.copy()], L=2, max_iter=4, min_tol=0.0)
            ref = ibss.fit_individual([X], [y], L=2, max_iter=4, min_tol=0.0)
            np.testing.assert_allclose(res.elbo[1:], np.asarray(ref.elbo)[1:], rtol=1e-6)
            assert np.max(np.abs(res.pip_all - ibss.make_pip(ref.posteriors.alpha))) <= PIP_TOL
        config.team_size = 8
        idx = np.arange(p)
        d = np.abs(idx[:, None] - idx[None, :])
        ld = np.where(d < 6, 0.5 ** d, 0.0)
        z = rng.standard_normal(p)
        z[[300, 8000]] += (8.0, -7.0)
        ns = np.array([[3000]])
        res = sushie_b200.infer_sushie_ss([ld], ns, [z], L=2, max_iter=3, min_tol=0.0)
        ref = ibss.fit_summary([ld], ns, [z], L=2, max_iter=3, min_tol=0.0)
        np.testing.assert_allclose(res.elbo[1:], np.asarray(ref.elbo)[1:], rtol=1e-6)
        assert np.max(np.abs(res.pip_all - ibss.make_pip(ref.posteriors.alpha))) <= PIP_TOL
    finally:
        config.team_size = old


def test_summary_level_six_ancestries(engine_lib):
    import sushie_b200
    from oracle import ibss

    k = 6
    Xs, ys = _locus(31, k, (150, 170, 140, 160, 130, 180), 120)
    lds, zs = [], []
    for X, y in zip(Xs, ys):
        Xstd = (X - X.mean(0)) / X.std(0)
        lds.append(Xstd.T @ Xstd / X.shape[0])
        zs.append(Xstd.T @ ((y - y.mean()) / y.std()) / np.sqrt(X.shape[0]))
    ns = np.array([[x.shape[0]] for x in Xs])
    res = sushie_b200.infer_sushie_ss(lds, ns, zs, L=3, min_tol=1e-4)
    ref = ibss.fit_summary(lds, ns, zs, L=3, min_tol=1e-4)
    _check(res, ref)


def test_more_ancestries_than_kernels_is_an_error(engine_lib):
    import sushie_b200

    Xs, ys = _locus(32, 7, (40,) * 7, 110)
    with pytest.raises(ValueError, match="at most 6 ancestries"):
        sushie_b200.infer_sushie(Xs, ys, L=2)
