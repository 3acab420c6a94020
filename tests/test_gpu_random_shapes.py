"""Seeded random shapes / option mixes against the oracle, both genotype storages, several team sizes.
Catches shape-dependent indexing mistakes (ragged n, p not a multiple of the vector width, L = 1, k = 4,
teams larger than the number of bands of an ancestry, ...)."""
import numpy as np
import pytest

from conftest import STORAGE_MODES, genotype_storage
from parity import COV_RTOL, ELBO_RTOL, PIP_TOL

pytestmark = pytest.mark.gpu


def _case(seed):
    rng = np.random.default_rng(1000 + seed)
    k = int(rng.integers(1, 5))
    n = [int(v) for v in rng.integers(12, 260, size=k)]
    p = int(rng.choice([101, 130, 257, 400, 777, 1025, 1300, 2049]))
    L = int(rng.integers(1, 6))
    maf = rng.uniform(0.08, 0.5, size=p)
    n_causal = int(rng.integers(0, 3))
    causal = rng.choice(p, size=n_causal, replace=False)
    beta = np.zeros(p)
    beta[causal] = rng.choice([-1.0, 1.0], size=n_causal) * rng.uniform(0.5, 1.2, size=n_causal)
    Xs, ys = [], []
    for a in range(k):
        g = rng.binomial(2, maf, size=(n[a], p)).astype(float)
        for j in np.nonzero(g.std(axis=0) == 0)[0]:
            g[0, j] = 1.0 if g[0, j] != 1.0 else 0.0
        gs = (g - g.mean(0)) / g.std(0)
        ys.append(gs @ beta + rng.standard_normal(n[a]))
        Xs.append(g)
    kw = dict(L=L, max_iter=int(rng.choice([3, 8, 40])), min_tol=float(rng.choice([0.0, 1e-4])))
    mode = int(rng.integers(0, 4))
    if mode == 1:
        kw["no_update"] = True
    elif mode == 2:
        kw["no_scale"] = True
    elif mode == 3 and k > 1:
        kw.update(no_update=True, rho=[float(v) for v in rng.uniform(0.05, 0.6, size=k * (k - 1) // 2)])
    team = int(rng.choice([1, 1, 2, 4, 8, 16, 24]))
    return Xs, ys, kw, team, dict(k=k, n=n, p=p, team=team, **kw)


@pytest.mark.parametrize("storage", STORAGE_MODES)
@pytest.mark.parametrize("seed", range(14))
def test_random_individual_level_shapes(engine_lib, seed, storage):
    import sushie_b200
    from oracle import ibss
    from sushie_b200 import config

    Xs, ys, kw, team, desc = _case(seed)
    old = config.team_size
    config.team_size = team
    try:
        with genotype_storage(storage):
            res = sushie_b200.infer_sushie([x.copy() for x in Xs], [y.copy() for y in ys], min_snps=50, **kw)
    finally:
        config.team_size = old
    ref = ibss.fit_individual(Xs, ys, **kw)
    assert len(res.elbo) - 1 == ref.n_iter, desc
    assert bool(res.elbo_increase) == bool(ref.elbo_increase), desc
    fin = np.isfinite(np.asarray(ref.elbo)[1:])
    np.testing.assert_allclose(res.elbo[1:][fin], np.asarray(ref.elbo)[1:][fin], rtol=ELBO_RTOL, err_msg=str(desc))
    assert np.max(np.abs(res.pip_all - ibss.make_pip(ref.posteriors.alpha))) <= PIP_TOL, desc
    tr_ref = np.sort(np.trace(ref.priors.effect_covar, axis1=1, axis2=2))
    tr_got = np.sort(np.trace(res.priors.effect_covar, axis1=1, axis2=2))
    np.testing.assert_allclose(tr_got, tr_ref, rtol=COV_RTOL, err_msg=str(desc))


@pytest.mark.parametrize("seed", range(8))
def test_random_summary_level_shapes(engine_lib, seed):
    import sushie_b200
    from oracle import ibss
    from sushie_b200 import config

    rng = np.random.default_rng(5000 + seed)
    k = int(rng.integers(1, 5))
    p = int(rng.choice([101, 257, 640, 1025, 1500]))
    L = int(rng.integers(1, 5))
    lds, zs, ns = [], [], []
    for a in range(k):
        n = int(rng.integers(150, 400))
        g = rng.binomial(2, rng.uniform(0.1, 0.5, size=p), size=(n, p)).astype(float)
        for j in np.nonzero(g.std(axis=0) == 0)[0]:
            g[0, j] = 1.0 if g[0, j] != 1.0 else 0.0
        gs = (g - g.mean(0)) / g.std(0)
        lds.append(gs.T @ gs / n)
        z = rng.standard_normal(p)
        z[int(rng.integers(0, p))] += 7.0
        zs.append(z)
        ns.append([n])
    kw = dict(L=L, max_iter=int(rng.choice([3, 10])), min_tol=0.0)
    team = int(rng.choice([1, 2, 4, 16, 24]))
    old = config.team_size
    config.team_size = team
    try:
        res = sushie_b200.infer_sushie_ss(lds, np.array(ns), zs, **kw)
    finally:
        config.team_size = old
    ref = ibss.fit_summary(lds, np.array(ns), zs, **kw)
    desc = dict(k=k, p=p, team=team, **kw)
    assert len(res.elbo) - 1 == ref.n_iter, desc
    fin = np.isfinite(np.asarray(ref.elbo)[1:])
    np.testing.assert_allclose(res.elbo[1:][fin], np.asarray(ref.elbo)[1:][fin], rtol=ELBO_RTOL, err_msg=str(desc))
    assert np.max(np.abs(res.pip_all - ibss.make_pip(ref.posteriors.alpha))) <= PIP_TOL, desc
