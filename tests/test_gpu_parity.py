"""GPU parity tests: the CUDA path through the Python API / C ABI vs the committed goldens
(tests/golden/*.npz: outputs of the reference's own source run through tests/refshim, produced by
tests/golden/make_golden.py)."""
import numpy as np
import pytest

from conftest import STORAGE_MODES, genotype_storage
from parity import assert_fit_matches

pytestmark = pytest.mark.gpu

SYN_CASES = {
    "k1": dict(k=1, kw=dict(L=5)),
    "k2": dict(k=2, kw=dict(L=5)),
    "k3": dict(k=3, kw=dict(L=6)),
    "k3null": dict(k=3, kw=dict(L=4)),
    "k2noupd": dict(k=2, kw=dict(L=4, no_update=True)),
    "k2rho": dict(k=2, kw=dict(L=4, no_update=True, rho=[0.5])),
    "k2var": dict(k=2, kw=dict(L=4, no_update=True, effect_var=[0.01, 0.02])),
    "k2noscale": dict(k=2, kw=dict(L=4, no_scale=True)),
    "k4": dict(k=4, kw=dict(L=4)),
}


@pytest.fixture(params=STORAGE_MODES)
def geno_storage(request):
    """Run a test with standardised fp64 storage, with bit-plane hard-call storage (the default for 0/1/2 genotypes)
    and with one-byte hard-call storage."""
    with genotype_storage(request.param) as mode:
        yield mode


def _syn_inputs(g, tag, k, as_float=True):
    Xs = [g[f"{tag}_X{i}"] for i in range(k)]
    if as_float:
        Xs = [x.astype(np.float64) for x in Xs]
    ys = [g[f"{tag}_y{i}"].copy() for i in range(k)]
    return Xs, ys


@pytest.mark.parametrize("tol_tag,tol", [("tol3", 1e-3), ("tol4", 1e-4)])
def test_bundled_individual_example(engine_lib, golden_ind, tol_tag, tol):
    """Config 1: bundled EUR+AFR genotypes + covariates, L=10."""
    import sushie_b200

    g = golden_ind
    Xs = [g["X0"].astype(np.float64), g["X1"].astype(np.float64)]
    ys = [g["y0"].copy(), g["y1"].copy()]
    covar = [g["covar0"], g["covar1"]]
    res = sushie_b200.infer_sushie(Xs, ys, covar, L=10, min_tol=tol)
    assert_fit_matches(res, g, tol_tag)
    snp = g["snp"]
    top = set(snp[np.argsort(-res.pip_all)[:2]])
    assert top == {"rs1886340", "rs10914958"}  # documented causal SNPs, docs/manual.rst:88


@pytest.mark.parametrize("tag", ["k2", "k3"])
@pytest.mark.parametrize("tol_tag,tol", [("tol3", 1e-3), ("tol4", 1e-4)])
def test_bundled_summary_example(engine_lib, golden_ss, tag, tol_tag, tol):
    """Config 2: bundled LD + GWAS z-scores."""
    import sushie_b200

    g = golden_ss
    res = sushie_b200.infer_sushie_ss(g[f"{tag}_lds"], g[f"{tag}_ns"], g[f"{tag}_zs"], L=10, min_tol=tol)
    assert_fit_matches(res, g, f"{tag}_{tol_tag}")
    top = set(g[f"{tag}_snp"][np.argsort(-res.pip_all)[:2]])
    assert top == {"rs1886340", "rs10914958"}


@pytest.mark.parametrize("tag", sorted(SYN_CASES))
def test_synthetic_small(engine_lib, golden_syn, tag, geno_storage):
    import sushie_b200

    case = SYN_CASES[tag]
    Xs, ys = _syn_inputs(golden_syn, tag, case["k"])
    res = sushie_b200.infer_sushie(Xs, ys, min_tol=1e-4, **case["kw"])
    assert_fit_matches(res, golden_syn, tag)


def test_int8_genotypes_match_float(engine_lib, golden_syn, geno_storage):
    """Hard-call int8 input is standardised on device exactly like float input."""
    import sushie_b200

    Xi, ys = _syn_inputs(golden_syn, "k3", 3, as_float=False)
    assert Xi[0].dtype == np.int8
    res = sushie_b200.infer_sushie(Xi, ys, min_tol=1e-4, L=6)
    assert_fit_matches(res, golden_syn, "k3")


@pytest.mark.parametrize("team", [1, 2, 4, 8, 16, 24])
def test_team_sizes_agree_with_oracle(engine_lib, golden_syn, team, geno_storage):
    """Cluster teams (hardware barrier, 2..16 blocks) and co-operative teams (24 blocks,
    global-memory barrier) all reproduce the oracle."""
    import sushie_b200
    from sushie_b200 import config

    old = config.team_size
    config.team_size = team
    try:
        Xs, ys = _syn_inputs(golden_syn, "k3", 3)
        res = sushie_b200.infer_sushie(Xs, ys, min_tol=1e-4, L=6)
        assert_fit_matches(res, golden_syn, "k3")
        res = sushie_b200.infer_sushie_ss(golden_ss_inputs()[0], golden_ss_inputs()[1], golden_ss_inputs()[2], L=10, min_tol=1e-3)
        assert_fit_matches(res, golden_ss_inputs()[3], "k2_tol3")
    finally:
        config.team_size = old


def golden_ss_inputs():
    import os

    from conftest import GOLDEN

    g = np.load(os.path.join(GOLDEN, "example_ss.npz"))
    return g["k2_lds"], g["k2_ns"], g["k2_zs"], g


def test_batch_mixed_shapes_matches_single(engine_lib, golden_syn, geno_storage):
    """Ragged p / n and mixed k in one device-resident batch == the same loci one by one."""
    import sushie_b200

    problems = []
    for tag in ("k1", "k2", "k3", "k3null", "k4"):
        case = SYN_CASES[tag]
        Xs, ys = _syn_inputs(golden_syn, tag, case["k"])
        problems.append(dict(Xs=Xs, ys=ys, **case["kw"]))
    out = sushie_b200.infer_sushie_batch(problems, min_tol=1e-4)
    for tag, res in zip(("k1", "k2", "k3", "k3null", "k4"), out):
        assert_fit_matches(res, golden_syn, tag)


def test_fp32_storage_is_close(engine_lib, golden_syn):
    """fp32 device storage (fp64 arithmetic) is reported, not gated at 1e-4: it must find
    the same credible sets and PIPs within 1e-2 on a well-powered locus."""
    import sushie_b200
    from sushie_b200 import config

    old = config.precision
    config.precision = 32
    try:
        Xs, ys = _syn_inputs(golden_syn, "k2", 2)
        res = sushie_b200.infer_sushie(Xs, ys, min_tol=1e-4, L=5)
    finally:
        config.precision = old
    assert np.max(np.abs(res.pip_all - golden_syn["k2_pip_all"])) < 1e-2
    assert set(res.cs.SNPIndex.values.tolist()) == set(golden_syn["k2_cs_snp"].tolist())


def test_make_cs_standalone_matches_oracle(engine_lib, golden_syn):
    """Public make_cs (device purity) vs the pandas oracle on the same posterior."""
    import sushie_b200
    from oracle import credible_sets, ibss

    Xs, ys = _syn_inputs(golden_syn, "k3", 3)
    Xstd, _ = ibss.prepare_individual(Xs, ys)
    alpha, log_bf = golden_syn["k3_alpha"], golden_syn["k3_log_bf"]
    ns = np.array([[x.shape[0]] for x in Xs])
    for method in ("weighted", "max", "min"):
        for max_select in (500, 40):
            cs_o, al_o, pa_o, pc_o = credible_sets.make_cs(alpha, log_bf, ns, Xs=Xstd, threshold=0.9, purity=0.5,
                                                          purity_method=method, max_select=max_select, seed=12345)
            cs_g, al_g, pa_g, pc_g = sushie_b200.make_cs(alpha, log_bf, ns, Xs=Xstd, threshold=0.9, purity=0.5,
                                                        purity_method=method, max_select=max_select, seed=12345)
            assert list(al_o.columns) == list(al_g.columns)
            np.testing.assert_allclose(al_g.values.astype(float), al_o.values.astype(float), atol=1e-9)
            assert np.array_equal(cs_g.SNPIndex.values, cs_o.SNPIndex.values.astype(int))
            assert np.array_equal(cs_g.CSIndex.values, cs_o.CSIndex.values.astype(int))
            np.testing.assert_allclose(pc_g, pc_o, atol=1e-12)


def test_headline_shape_fixed_iterations(engine_lib):
    """Full headline shape (k=3, n=500, p=2000, L=10): a fixed 12 outer iterations must
    match the oracle's ELBO trace and PIPs, plus size-independent invariants."""
    import sushie_b200
    from oracle import ibss

    rng = np.random.default_rng(20260001)
    k, n, p, L = 3, 500, 2000, 10
    Xs = [rng.binomial(2, rng.uniform(0.05, 0.5, size=p), size=(n, p)).astype(np.float64) for _ in range(k)]
    for X in Xs:
        X[0, X.std(axis=0) == 0] = 1.0
    causal = [150, 1200]
    ys = []
    for X in Xs:
        Xs_std = (X - X.mean(0)) / X.std(0)
        ys.append(Xs_std[:, causal] @ np.array([0.35, -0.3]) + rng.standard_normal(n))
    res = sushie_b200.infer_sushie(Xs, ys, L=L, max_iter=12, min_tol=0.0)
    ref = ibss.fit_individual(Xs, ys, L=L, max_iter=12, min_tol=0.0)
    assert len(res.elbo) == len(ref.elbo) == 13
    np.testing.assert_allclose(res.elbo[1:], ref.elbo[1:], rtol=1e-6)
    assert np.max(np.abs(res.pip_all - ibss.make_pip(ref.posteriors.alpha))) <= 1e-4
    # invariants
    assert np.all(np.diff(res.elbo[1:]) >= -1e-5 * np.abs(res.elbo[1:-1]) - 1e-8)
    np.testing.assert_allclose(res.posteriors.alpha.sum(axis=1), 1.0, atol=1e-10)
    assert np.all((res.pip_all >= 0) & (res.pip_all <= 1))
    ev = np.linalg.eigvalsh(res.posteriors.weighted_sum_covar)
    assert np.all(ev >= -1e-12)


def test_phased_schedule_matches_single_phase_and_reruns(engine_lib, golden_syn):
    """Suspend / regroup phases (loci continued by larger teams as the batch drains) must
    not change results, and a resident batch must give the same answer on every run."""
    from sushie_b200 import _capi
    from sushie_b200._common import build_priors

    eng = _capi.get_engine(0)
    tags = ["k2", "k3", "k3null", "k4", "k1", "k2noscale"] * 2
    def problems():
        out = []
        for tag in tags:
            case = SYN_CASES[tag]
            Xs, ys = _syn_inputs(golden_syn, tag, case["k"])
            k, L = case["k"], case["kw"]["L"]
            ys2 = [(y - y.mean()) / y.std() for y in ys]
            rv = [np.var(y, ddof=1) for y in ys2]
            ec, adj = build_priors(k, L, None, None, False, summary=False)
            out.append(_capi.IndProblem(Xs, ys2, L, None, rv, ec, adj.times, adj.plus, True, standardize=3))
        return out

    fits = {}
    for name, single, team in (("single", True, 1), ("phased", False, 1), ("phased2", False, 2)):
        opts = eng.make_opts(500, 1e-4, _capi.SB_F64, team, single)
        with eng.create_batch(problems(), opts) as b:
            b.run()
            first = b.fetch()
            b.run()  # second pass over the same resident batch
            second = b.fetch()
        for a, c in zip(first, second):
            assert a.n_iter == c.n_iter
            if single:  # fixed team size: bit-reproducible
                np.testing.assert_array_equal(a.elbo, c.elbo)
            else:  # the iteration at which a locus changes team size depends on timing: rounding-level only
                np.testing.assert_allclose(a.elbo[1:], c.elbo[1:], rtol=1e-11)
        fits[name] = first
    for name in ("phased", "phased2"):
        for a, c in zip(fits["single"], fits[name]):
            assert a.n_iter == c.n_iter
            np.testing.assert_allclose(a.elbo[1:], c.elbo[1:], rtol=1e-9)
            np.testing.assert_allclose(a.alpha, c.alpha, atol=1e-8)


def test_summary_level_large_crop_fp32_storage(engine_lib):
    """Config 4 crop: k=3, p=2000 AR(1) LD stored fp32 (fp64 arithmetic), co-operative team over
    many blocks; compared with the fp64 oracle fed the same fp32-rounded LD."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import bench_ss
    import sushie_b200
    from sushie_b200 import config
    from oracle import ibss

    lds, ns, zs, causal = bench_ss.make_locus(2000, np.float32)
    old_p, old_t = config.precision, config.team_size
    config.precision, config.team_size = 32, 40
    try:
        res = sushie_b200.infer_sushie_ss(lds, ns, zs, L=10, max_iter=6, min_tol=0.0)
    finally:
        config.precision, config.team_size = old_p, old_t
    ref = ibss.fit_summary([ld.astype(np.float64) for ld in lds], ns, zs, L=10, max_iter=6, min_tol=0.0)
    assert len(res.elbo) == len(ref.elbo) == 7
    np.testing.assert_allclose(res.elbo[1:], ref.elbo[1:], rtol=1e-6)
    assert np.max(np.abs(res.pip_all - ibss.make_pip(ref.posteriors.alpha))) <= 1e-4


def test_hard_call_storage_is_selected_and_wide_loci_fit(engine_lib):
    """One-byte storage through the engine API: the band geometry of the row-team traversal is reported (one, two,
    four or eight warps per row by locus width), and every width agrees with the fp64 storage run."""
    from sushie_b200 import _capi
    from sushie_b200._common import build_priors

    rng = np.random.default_rng(11)
    eng = _capi.get_engine(0)
    for p, rows in ((300, 32), (700, 16), (1500, 8), (2500, 4), (5000, 2)):
        k, n, L = 2, [90, 70], 3
        Xs = [rng.binomial(2, 0.3, size=(n[i], p)).astype(np.int8) for i in range(k)]
        for X in Xs:
            X[0, X.std(axis=0) == 0] = 1
        beta = np.zeros(p)
        beta[[5, p // 2]] = [1.0, -0.8]
        ys = []
        for X in Xs:
            y = ((X - X.mean(0)) / X.std(0)) @ beta + rng.standard_normal(X.shape[0])
            ys.append((y - y.mean()) / y.std())
        ec, adj = build_priors(k, L, None, None, False, summary=False)
        fits = {}
        for storage in (_capi.SB_F64, _capi.SB_I8, _capi.SB_B2):
            prob = _capi.IndProblem(Xs, ys, L, None, [np.var(y, ddof=1) for y in ys], ec, adj.times, adj.plus, True,
                                    standardize=_capi.SB_CENTER | _capi.SB_SCALE)
            with eng.create_batch([prob], eng.make_opts(30, 1e-4, storage, 1)) as b:
                st = b.run()
                fits[storage] = b.fetch()[0]
                if storage == _capi.SB_I8:
                    assert st["rows_per_band"] == rows
        a = fits[_capi.SB_F64]
        for hard in (_capi.SB_I8, _capi.SB_B2):
            c = fits[hard]
            assert a.n_iter == c.n_iter
            np.testing.assert_allclose(a.elbo[1:], c.elbo[1:], rtol=1e-9)
            np.testing.assert_allclose(a.alpha, c.alpha, atol=1e-8)
            np.testing.assert_allclose(a.effect_covar, c.effect_covar, rtol=1e-6, atol=1e-12)


def test_batch_sharded_over_all_visible_gpus(engine_lib, golden_syn):
    """The host work queue with one worker (own sb_ctx) per visible GPU returns the same fits, in input order,
    as a single-GPU run -- also with hard-call and dense loci mixed in one call (they get separate chunks)."""
    import sushie_b200
    from sushie_b200 import _capi, config

    n_dev = _capi.device_count()
    assert n_dev >= 1
    problems, tags = [], []
    for rep in range(3):
        for tag in ("k2", "k3", "k3null", "k2noscale"):
            case = SYN_CASES[tag]
            Xs, ys = _syn_inputs(golden_syn, tag, case["k"])
            if rep == 1:
                Xs = [x + 0.5 for x in Xs]  # not hard calls any more (same standardised matrix): dense chunk
            problems.append(dict(Xs=Xs, ys=ys, **case["kw"]))
            tags.append(tag)
    old = config.chunk_loci
    config.chunk_loci = 2  # many chunks so every worker pulls several
    try:
        out = sushie_b200.infer_sushie_batch(problems, devices=list(range(n_dev)), min_tol=1e-4)
    finally:
        config.chunk_loci = old
    assert len(out) == len(problems)
    for tag, res in zip(tags, out):
        assert_fit_matches(res, golden_syn, tag)


@pytest.mark.parametrize("storage", ["dense", "i8"])
def test_purity_of_a_whole_chunk_in_one_launch(engine_lib, storage):
    """`sb_batch_purity_many`: ragged index sets (1 ... 300 SNPs, not multiples of the 4 x 4 tile) of loci with
    different ancestry counts and widths, against numpy on the standardised genotypes."""
    from sushie_b200 import _capi
    from sushie_b200._common import build_priors

    rng = np.random.default_rng(99)
    eng = _capi.get_engine(0)
    probs, Xstd, sets = [], [], []
    for k, n, p in ((1, [90], 330), (3, [64, 33, 120], 517), (4, [50, 70, 31, 32], 300), (2, [200, 150], 1100)):
        G = [rng.binomial(2, rng.uniform(0.1, 0.5, size=p), size=(ni, p)).astype(np.int8) for ni in n]
        for g in G:
            g[0, g.std(axis=0) == 0] += 1
        Xstd.append([(g - g.mean(0)) / g.std(0) for g in G])
        ec, adj = build_priors(k, 2, None, None, False, summary=False)
        mats = G if storage == "i8" else [g.astype(np.float64) for g in G]
        probs.append(_capi.IndProblem(mats, [rng.standard_normal(ni) for ni in n], 2, None, np.ones(k), ec, adj.times,
                                      adj.plus, True, standardize=_capi.SB_CENTER | _capi.SB_SCALE))
        sizes = [1, 3, 4, 5, 37, 250, min(p, 300)] + ([1000] if p > 1000 else [])  # 1000: fewer rows per pass
        sets.append([rng.choice(p, size=m, replace=False) for m in sizes])
    sets[2] = []  # a locus without sets
    opts = eng.make_opts(2, 1e-4, _capi.SB_I8 if storage == "i8" else _capi.SB_F64, 1)
    with eng.create_batch(probs, opts) as b:
        got = b.purity_many(sets)  # (hard calls: the 1000-SNP set does not fit the integer kernel -> fp64 kernel)
        one = b.purity(1, sets[1])
        small = b.purity_many([ss[:7] for ss in sets])  # hard calls: exact integer Gram matrices (dp4a)
    assert [g.shape for g in got] == [(len(s), len(x)) for s, x in zip(sets, Xstd)]
    for res in (got, small):
        for g, ss, xs in zip(res, sets, Xstd):
            for row, idx in zip(g, ss):
                want = [np.min(np.abs(x[:, idx].T @ x[:, idx] / x.shape[0])) for x in xs]
                np.testing.assert_allclose(row, want, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(one, got[1], rtol=1e-10, atol=1e-12)


def test_integer_purity_kernel_with_covariates_and_monomorphic_columns(engine_lib):
    """Bit-plane storage with covariate-residualised hard calls: min |corr| of the residualised, scaled columns from
    the exact integer Gram matrix minus M M^T (M = G^T Q); a monomorphic column makes the purity NaN as jnp.min does."""
    from sushie_b200 import _capi, utils
    from sushie_b200._common import build_priors

    rng = np.random.default_rng(7)
    eng = _capi.get_engine(0)
    k, n, p = 2, [130, 101], 400
    G = [rng.binomial(2, rng.uniform(0.1, 0.5, size=p), size=(ni, p)).astype(np.int8) for ni in n]
    for g in G:
        g[0, g.std(axis=0) == 0] += 1
    G[1][:, 17] = 1  # monomorphic in ancestry 2
    cov = [rng.standard_normal((ni, 3)) for ni in n]
    Q = [utils.covar_basis(c) for c in cov]
    Xres = [g - q @ (q.T @ g) for g, q in zip(G, Q)]
    with np.errstate(all="ignore"):
        Xstd = [(x - x.mean(0)) / x.std(0) for x in Xres]
    ec, adj = build_priors(k, 2, None, None, False, summary=False)
    prob = _capi.IndProblem(G, [rng.standard_normal(ni) for ni in n], 2, None, np.ones(k), ec, adj.times, adj.plus, True,
                            standardize=_capi.SB_CENTER | _capi.SB_SCALE, covar_q=Q)
    clean = np.setdiff1d(np.arange(p), [17])
    sets = [rng.choice(clean, size=m, replace=False) for m in (2, 9, 64, 250)] + [np.array([3, 17, 40])]
    with eng.create_batch([prob], eng.make_opts(2, 1e-4, _capi.SB_B2, 1)) as b:
        got = b.purity(0, sets)
        # the same genotypes without covariates: the monomorphic column is 0 / 0 after standardising -> NaN purity
        plain = _capi.IndProblem(G, [rng.standard_normal(ni) for ni in n], 2, None, np.ones(k), ec, adj.times, adj.plus,
                                 True, standardize=_capi.SB_CENTER | _capi.SB_SCALE)
    with eng.create_batch([plain], eng.make_opts(2, 1e-4, _capi.SB_B2, 1)) as b:
        nan_case = b.purity(0, sets)
    for row, idx in zip(got[:-1], sets[:-1]):
        want = [np.min(np.abs(x[:, idx].T @ x[:, idx] / x.shape[0])) for x in Xstd]
        np.testing.assert_allclose(row, want, rtol=1e-9, atol=1e-11)
    # (with covariates the residual of a constant column is rounding noise, in the reference too: no claim)
    assert np.isfinite(nan_case[-1][0]) and np.isnan(nan_case[-1][1])
    assert np.all(np.isfinite(nan_case[:-1]))


def test_device_set_selection_equals_numpy(engine_lib):
    """`sb_batch_select` on fitted loci of different widths (one just past a power of two, one exactly on it): SNP
    order = descending alpha with ties by index, csum = numpy's left-to-right cumsum bit for bit, sizes = the
    reference's cut; per-problem thresholds."""
    import bench
    from sushie_b200 import _capi
    from sushie_b200._common import build_priors

    eng = _capi.get_engine(0)
    probs = []
    for g, p, L in ((0, 300, 3), (1, 1025, 10), (2, 2048, 5), (3, 4500, 10)):
        Xs, ys = bench.make_gene(g, 120, p)
        ec, adj = build_priors(3, L, None, None, False, summary=False)
        yy, rv = bench.prep_y(ys)
        probs.append(_capi.IndProblem(Xs, yy, L, None, rv, ec, adj.times, adj.plus, True,
                                      standardize=_capi.SB_CENTER | _capi.SB_SCALE))
    thr = [0.95, 0.9, 0.5, 0.99]
    with eng.create_batch(probs, eng.make_opts(6, 1e-4, _capi.SB_B2, 1)) as b:
        b.run()
        fits = b.fetch()
        sel = b.select(thr)
    for fit, (order, csum, sizes), t in zip(fits, sel, thr):
        L, p = fit.alpha.shape
        for l in range(L):
            a = fit.alpha[l]
            want = np.lexsort((np.arange(p), -a))
            np.testing.assert_array_equal(order[l], want)
            c = np.cumsum(a[want])
            np.testing.assert_array_equal(csum[l], c)
            n_row = int(np.count_nonzero(c < t))
            assert sizes[l] == (n_row if n_row == p else n_row + 1)


def test_row_view_folds_equal_materialised_folds(engine_lib, geno_storage):
    """Cross-validation folds passed as row views of the gene's matrix (one upload, rows gathered on the device) give
    the fits of the same folds passed as materialised matrices (to rounding) -- hard calls and dosages."""
    import bench
    import sushie_b200
    from sushie_b200 import cli

    Xs, ys = bench.make_gene(5, 150, 700)
    for mats in (list(Xs), [x.astype(np.float64) + 0.125 for x in Xs]):
        folds = cli._prepare_cv(list(mats), list(ys), 3, 12345)
        views = [dict(Xs=mats, ys=ys)] + [dict(Xs=f.train_geno, ys=f.train_pheno) for f in folds]
        dense = [dict(Xs=mats, ys=ys)] + [dict(Xs=[np.asarray(x) for x in f.train_geno], ys=f.train_pheno) for f in folds]
        a = sushie_b200.infer_sushie_batch(views, L=5, max_iter=30)
        b = sushie_b200.infer_sushie_batch(dense, L=5, max_iter=30)
        for ra, rb in zip(a, b):
            # rounding level, not bits: in a multi-locus batch the iteration at which a locus moves to a larger team
            # (phased schedule) depends on timing, which changes the order of a few reductions
            assert len(ra.elbo) == len(rb.elbo)
            np.testing.assert_allclose(ra.elbo[1:], rb.elbo[1:], rtol=1e-11)
            np.testing.assert_allclose(ra.posteriors.alpha, rb.posteriors.alpha, rtol=0, atol=1e-9)
            np.testing.assert_allclose(ra.posteriors.post_mean, rb.posteriors.post_mean, rtol=0, atol=1e-9)
            np.testing.assert_array_equal(ra.l_order, rb.l_order)
            assert ra.cs[["CSIndex", "SNPIndex"]].equals(rb.cs[["CSIndex", "SNPIndex"]])
            np.testing.assert_allclose(ra.pip_all, rb.pip_all, rtol=0, atol=1e-9)


# ---- BASELINE.json configurations at their real shapes ------------------------------------------------


@pytest.mark.parametrize("tol_tag,tol", [("tol4", 1e-4), ("tol3", 1e-3)])
@pytest.mark.parametrize("gene", [0, 1, 2, 3])
def test_headline_genes_to_convergence(engine_lib, golden_headline, gene, tol_tag, tol, geno_storage):
    """Config 3: genes 0..3 of the bench generator (k=3, n=500, p=2000, L=10; null gene, 1, 2, 3 causal SNPs) fitted
    TO CONVERGENCE at the API (1e-4) and CLI (1e-3) tolerances, both genotype storages, against the reference's
    own source: same iteration count, ELBO trace, PIPs, credible sets, prior covariances."""
    import bench
    import sushie_b200

    Xs, ys = bench.make_gene(gene)
    res = sushie_b200.infer_sushie([x.astype(np.float64) for x in Xs], ys, L=bench.L_EFF, min_tol=tol)
    assert_fit_matches(res, golden_headline, f"g{gene}_{tol_tag}")


def test_config5_gene_full_and_fold_fits(engine_lib, golden_config5, geno_storage):
    """Config 5: one ragged-width gene, the fit on all samples plus the 5 cross-validation training folds as ONE
    device batch; every fit and the out-of-fold scores against the reference's ``_prepare_cv`` / ``_run_cv``."""
    import bench
    import bench_cv
    import sushie_b200
    from sushie_b200 import cli

    g = golden_config5
    gene, cv_num, seed = int(g["gene"]), int(g["cv_num"]), int(g["seed"])
    p = bench_cv.gene_width(gene, 500, 4000)
    assert p == int(g["p"]) and p % 32 != 0
    Xs, ys = bench.make_gene(gene, bench.N_IND, p)
    folds = cli._prepare_cv(list(Xs), list(ys), cv_num, seed)
    for j, fold in enumerate(folds):
        for a in range(len(Xs)):
            np.testing.assert_array_equal(fold.train_geno[a], Xs[a][g[f"fold{j}_train_rows{a}"]])
            np.testing.assert_allclose(fold.train_pheno[a], g[f"fold{j}_train_pheno{a}"], rtol=0, atol=1e-12)
            np.testing.assert_allclose(fold.valid_pheno[a], g[f"fold{j}_valid_pheno{a}"], rtol=0, atol=1e-12)
    specs = [dict(Xs=Xs, ys=ys)] + [dict(Xs=f.train_geno, ys=f.train_pheno) for f in folds]
    fits = sushie_b200.infer_sushie_batch(specs, L=bench.L_EFF, min_tol=1e-4)
    assert_fit_matches(fits[0], g, "full")
    for j in range(cv_num):
        assert_fit_matches(fits[1 + j], g, f"fold{j}")
    scores = np.asarray(cli._cv_scores(folds, fits[1:]))
    np.testing.assert_allclose(scores[:, 0], g["cv_res"][:, 0], atol=1e-8)             # adjusted r-squared
    np.testing.assert_allclose(scores[:, 1], g["cv_res"][:, 1], rtol=1e-6, atol=1e-300)  # p-value


@pytest.mark.slow
def test_config4_full_size_fp32_storage(engine_lib, golden_config4):
    """Config 4 at its full size: k=3, p=10 000 LD stored fp32 (1.2 GB, fp64 arithmetic), L=10, EM on, 3 fixed
    iterations, co-operative team over the whole device, against the reference's source on the same
    fp32-rounded LD in fp64."""
    import bench_ss
    import sushie_b200
    from sushie_b200 import config

    g = golden_config4
    lds, ns, zs, causal = bench_ss.make_locus(int(g["p"]), np.float32)
    assert np.array_equal(causal, g["causal"])
    old = config.precision
    config.precision = 32
    try:
        res = sushie_b200.infer_sushie_ss(lds, ns, zs, L=10, max_iter=int(g["iters"]), min_tol=0.0)
    finally:
        config.precision = old
    assert_fit_matches(res, g, "c4")


def test_device_set_selection_of_a_locus_too_wide_for_shared_memory(engine_lib):
    """p = 17 000 pads to 32 768 sort keys (384 KB): `sb_batch_select` sorts in a global-memory scratch slot per effect.
    Same contract as the shared-memory path; a narrow locus in the same batch takes the shared-memory launch."""
    from sushie_b200 import _capi
    from sushie_b200._common import build_priors

    rng = np.random.default_rng(21)
    eng = _capi.get_engine(0)
    probs = []
    for p, n in ((17000, [40, 30]), (700, [40, 30])):
        Xs = [rng.standard_normal((ni, p)).astype(np.float32) for ni in n]
        beta = np.zeros(p)
        beta[[5, p - 3]] = [1.0, -1.0]
        ys = [x.astype(np.float64) @ beta + rng.standard_normal(x.shape[0]) for x in Xs]
        ys = [(y - y.mean()) / y.std() for y in ys]
        ec, adj = build_priors(2, 3, None, None, False, summary=False)
        probs.append(_capi.IndProblem(Xs, ys, 3, None, [np.var(y, ddof=1) for y in ys], ec, adj.times, adj.plus, True,
                                      standardize=_capi.SB_CENTER | _capi.SB_SCALE))
    with eng.create_batch(probs, eng.make_opts(3, 1e-4, _capi.SB_F32, 1)) as b:
        b.run()
        fits = b.fetch()
        sel = b.select([0.95, 0.8])
    for fit, (order, csum, sizes), t in zip(fits, sel, (0.95, 0.8)):
        L, p = fit.alpha.shape
        for l in range(L):
            a = fit.alpha[l]
            want = np.lexsort((np.arange(p), -a))
            np.testing.assert_array_equal(order[l], want)
            c = np.cumsum(a[want])
            np.testing.assert_array_equal(csum[l], c)
            n_row = int(np.count_nonzero(c < t))
            assert sizes[l] == (n_row if n_row == p else n_row + 1)
