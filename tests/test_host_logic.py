"""Host-side logic that needs no GPU: validation texts, prior construction, credible-set
selection and table assembly (with a numpy purity callback), chunk planning, C-ABI export."""
import ctypes
import os
import re

import numpy as np
import pandas as pd
import pytest

import sushie_b200
from sushie_b200 import _capi, _common, batch, config, cs as cs_mod, utils
from oracle import credible_sets, ibss, threefry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _xy(n=60, p=120, k=2, seed=0):
    rng = np.random.default_rng(seed)
    return [rng.standard_normal((n, p)) for _ in range(k)], [rng.standard_normal(n) for _ in range(k)]


def test_validation_messages_individual():
    Xs, ys = _xy()
    with pytest.raises(ValueError, match=r"The number of geno \(2\) and pheno \(1\) data does not match"):
        sushie_b200.infer_sushie(Xs, ys[:1])
    with pytest.raises(ValueError, match=r"Ancestry 1: The sample size of geno \(60\) and pheno \(59\)"):
        sushie_b200.infer_sushie(Xs, [ys[0][:59], ys[1]])
    with pytest.raises(ValueError, match=r"Inferred L \(0\) is invalid, choose a positive L\."):
        sushie_b200.infer_sushie(Xs, ys, L=0)
    with pytest.raises(ValueError, match=r"Credible set PIP threshold \(1\.5\) must be greater than 0 and less than 1\."):
        sushie_b200.infer_sushie(Xs, ys, threshold=1.5)
    with pytest.raises(ValueError, match=r"Purity threshold \(0\) must be greater than or equal to 0 and less than 1\."):
        sushie_b200.infer_sushie(Xs, ys, purity=0)
    with pytest.raises(ValueError, match=r"maximum selected number of SNPs for purity must be greater than 0"):
        sushie_b200.infer_sushie(Xs, ys, max_select=0)
    with pytest.raises(ValueError, match=r"minimum number of SNPs to fine-map must be greater than 0"):
        sushie_b200.infer_sushie(Xs, ys, min_snps=0)
    with pytest.raises(ValueError, match=r"minimum common SNPs across ancestries \(5\) is less than inferred L \(10\)"):
        sushie_b200.infer_sushie(Xs, ys, min_snps=5)
    with pytest.raises(ValueError, match=r"common SNPs across ancestries \(120\) is less than minimum common number of SNPs \(100\)"):
        sushie_b200.infer_sushie(Xs, ys, min_snps=200)
    with pytest.raises(ValueError, match=r"Prior probability/weights must be all positive values\."):
        sushie_b200.infer_sushie(Xs, ys, pi=np.zeros(120))
    with pytest.raises(ValueError, match=r"Prior probability/weights \(3\) does not match the number of SNPs \(120\)"):
        sushie_b200.infer_sushie(Xs, ys, pi=np.ones(3))
    with pytest.raises(ValueError, match=r"Number of specified residual prior \(1\) does not match ancestry number \(2\)"):
        sushie_b200.infer_sushie(Xs, ys, resid_var=[1.0])
    with pytest.raises(ValueError, match=r"The input of residual prior \(\[1\.0, -1\.0\]\) is invalid"):
        sushie_b200.infer_sushie(Xs, ys, resid_var=[1.0, -1.0])
    with pytest.raises(ValueError, match=r"Number of specified effect prior \(3\) does not match ancestry number \(2\)"):
        sushie_b200.infer_sushie(Xs, ys, effect_var=[1, 2, 3])
    with pytest.raises(ValueError, match=r"The effect size prior variance \(\[1\.0, 0\.0\]\) must be positive"):
        sushie_b200.infer_sushie(Xs, ys, effect_var=[1, 0])
    with pytest.raises(ValueError, match=r"Number of specified rho \(2\) does not match expected number 1\."):
        sushie_b200.infer_sushie(Xs, ys, rho=[0.1, 0.2])
    with pytest.raises(ValueError, match=r"Effect size prior correlation \(\[1\.0\]\) must be between -1 and 1"):
        sushie_b200.infer_sushie(Xs, ys, rho=[1.0])  # |rho| == 1 rejected at individual level (infer.py:398)


def test_validation_messages_summary():
    p = 120
    lds = [np.eye(p), np.eye(p)]
    zs = [np.zeros(p), np.zeros(p)]
    ns = np.array([[100], [120]])
    with pytest.raises(ValueError, match=r"number of LD matrices \(1\) does not match the number of ancestries \(2\)"):
        sushie_b200.infer_sushie_ss(lds[:1], ns, zs)
    with pytest.raises(ValueError, match="LD matrices do not have the same shape"):
        sushie_b200.infer_sushie_ss([np.eye(p), np.eye(p + 1)], ns, zs)
    with pytest.raises(ValueError, match="LD matrices are not square matrices"):
        sushie_b200.infer_sushie_ss([np.ones((p, p + 1))] * 2, ns, zs)
    with pytest.raises(ValueError, match="Z scores are not provided"):
        sushie_b200.infer_sushie_ss(lds, ns, None)
    with pytest.raises(ValueError, match=r"number of Z scores \(1\) does not match"):
        sushie_b200.infer_sushie_ss(lds, ns, zs[:1])
    with pytest.raises(ValueError, match="Z scores do not have the same number of SNPs as the LD matrices"):
        sushie_b200.infer_sushie_ss(lds, ns, [np.zeros(p - 1)] * 2)
    with pytest.raises(ValueError, match="minimum number of SNPs to fine-map is invalid. Choose a positive integer"):
        sushie_b200.infer_sushie_ss(lds, ns, zs, min_snps=0)
    # |rho| == 1 is accepted at summary level (infer_ss.py:256) and then needs a GPU: EngineError, not ValueError
    with pytest.raises(_capi.EngineError):
        if _capi.device_count() > 0:
            raise _capi.EngineError("GPU present: covered by the gpu suite")
        sushie_b200.infer_sushie_ss(lds, ns, zs, rho=[1.0])


def test_no_cpu_fallback_without_gpu():
    """The product path must fail loudly when no CUDA device is usable."""
    if _capi.device_count() > 0:
        pytest.skip("a GPU is present")
    Xs, ys = _xy()
    with pytest.raises(_capi.EngineError, match="sb_init"):
        sushie_b200.infer_sushie(Xs, ys, L=2)


def test_prior_construction_matches_oracle():
    for k, ev, rho, no_update in [(1, None, None, False), (2, None, [0.5], True), (3, [0.1, 0.2, 0.3], None, True),
                                  (3, [0.1, 0.2, 0.3], [0.1, 0.2, 0.3], True), (4, None, None, False)]:
        ec, adj = _common.build_priors(k, 3, ev, rho, no_update, summary=False)
        ev_o = [1e-3] * k if ev is None else ev
        rho_o = [0.1] * (k * (k - 1) // 2) if rho is None else rho
        c0 = ibss.build_effect_covar(k, [float(v) for v in ev_o], [float(v) for v in rho_o])
        t, pl = ibss.build_adjustor(k, c0, no_update, ev, rho)
        np.testing.assert_array_equal(ec, np.stack([c0] * 3))
        np.testing.assert_array_equal(adj.times, t)
        np.testing.assert_array_equal(adj.plus, pl)


def test_pi_normalisation_rules():
    pi = np.full(10, 0.05)  # sums to 0.5
    np.testing.assert_array_equal(_common.resolve_pi(pi, 10, summary=False), pi)       # infer.py:313 only if > 1
    np.testing.assert_allclose(_common.resolve_pi(pi, 10, summary=True), pi / 0.5)     # infer_ss.py:189 if != 1
    np.testing.assert_allclose(_common.resolve_pi(pi * 4, 10, summary=False).sum(), 1.0)


def test_threefry_product_matches_oracle():
    for seed in (0, 12345, 2**40 + 7):
        for n, m in ((10, 3), (251, 250), (1900, 250)):
            vals = np.random.default_rng(n).permutation(5000)[:n]
            a = cs_mod.jax_choice_no_replace(seed, vals, m)
            b = threefry.choice_without_replacement(threefry.prng_key(seed), vals, m)
            assert np.array_equal(a, b)


def _numpy_purity(Xstd):
    def f(sets):
        out = np.empty((len(sets), len(Xstd)))
        for s, idx in enumerate(sets):
            for k, X in enumerate(Xstd):
                out[s, k] = np.min(np.abs(X[:, idx].T @ X[:, idx] / X.shape[0]))
        return out
    return f


@pytest.mark.parametrize("method", ["weighted", "max", "min"])
@pytest.mark.parametrize("max_select", [500, 30])
def test_build_tables_matches_oracle_make_cs(golden_syn, method, max_select):
    """Credible-set selection / table assembly vs the pandas oracle (purity supplied by numpy
    here; on the GPU box the same function is fed by the purity kernel)."""
    g = golden_syn
    Xs = [g[f"k3_X{i}"].astype(float) for i in range(3)]
    ys = [g[f"k3_y{i}"] for i in range(3)]
    Xstd, _ = ibss.prepare_individual(Xs, ys)
    ns = np.array([[x.shape[0]] for x in Xs])
    alpha, log_bf = g["k3_alpha"], g["k3_log_bf"]
    cs_o, al_o, pa_o, pc_o = credible_sets.make_cs(alpha, log_bf, ns, Xs=Xstd, threshold=0.9, purity=0.4,
                                                  purity_method=method, max_select=max_select, seed=7)
    cs_g, al_g, pa_g, pc_g = cs_mod.build_tables(alpha, log_bf, ns, _numpy_purity(Xstd), 0.9, 0.4, method,
                                                max_select, 7)
    assert list(al_g.columns) == list(al_o.columns)
    np.testing.assert_allclose(al_g.values.astype(float), al_o.values.astype(float), atol=1e-12)
    assert list(cs_g.columns) == list(cs_o.columns)
    np.testing.assert_allclose(cs_g.values.astype(float), cs_o.values.astype(float), atol=1e-12)
    np.testing.assert_allclose(pa_g, pa_o, atol=1e-15)
    np.testing.assert_allclose(pc_g, pc_o, atol=1e-15)


def test_build_tables_invalid_method():
    with pytest.raises(ValueError, match="Invalid purity method"):
        cs_mod.build_tables(np.ones((1, 3)) / 3, np.zeros((1, 3)), np.array([10]), lambda s: np.ones((1, 1)), 0.9,
                            0.5, "median", 10, 1)


def test_utils_ols_and_rint():
    rng = np.random.default_rng(1)
    X = rng.standard_normal((80, 3))
    y = X @ [1.0, -2.0, 0.5] + 0.1 * rng.standard_normal(80)
    resid, adj_r, pval = utils.ols(X, y)
    design = np.column_stack([np.ones(80), X])
    beta = np.linalg.lstsq(design, y, rcond=None)[0]
    np.testing.assert_allclose(resid[:, 0], y - design @ beta, atol=1e-10)
    assert adj_r[0] > 0.99 and pval.shape == (4, 1)
    np.testing.assert_allclose(ibss.ols_residual(X, y), resid, atol=1e-10)
    r = utils.rint(np.array([3.0, 1.0, 2.0]))
    assert r[1] < r[2] < r[0] and abs(r[2]) < 1e-12


def test_plan_chunks_covers_everything_once():
    class P:  # minimal stand-in for a prepared problem
        def __init__(self, p, tol):
            self.problem = type("X", (), dict(n=[100, 100], p=p, L=10, k=2, cost=200.0 * p * 10))()
            self.max_iter, self.min_tol = 500, tol
    preps = [P(100 + 7 * i, 1e-4 if i % 3 else 1e-3) for i in range(40)]
    chunks = batch.plan_chunks(preps, chunk_bytes=3_000_000, chunk_loci=0)
    flat = sorted(i for c in chunks for i in c)
    assert flat == list(range(40))
    for c in chunks:  # launch-compatible problems only
        assert len({(preps[i].max_iter, preps[i].min_tol) for i in c}) == 1
    q = batch.ChunkQueue(len(chunks))
    got = []
    while (c := q.next()) is not None:
        got.append(c)
    assert sorted(got) == list(range(len(chunks)))


def test_plan_chunks_memory_budget_beyond_int32():
    """Sample counts arrive as an int32 array (`IndProblem.n`); the running byte total of a chunk must not wrap at
    2 GiB, or the memory budget would never cut a chunk."""
    class P:
        def __init__(self):
            self.problem = type("X", (), dict(n=np.asarray([500, 500, 500], dtype=np.int32), p=4000, L=10, k=3,
                                              cost=1500.0 * 4000 * 10))()
            self.max_iter, self.min_tol = 500, 1e-4
    preps = [P() for _ in range(120)]
    one = batch.estimate_bytes(preps[0])
    assert type(one) is int and one > 60_000_000
    chunks = batch.plan_chunks(preps, chunk_bytes=3 << 30, chunk_loci=0)
    assert len(chunks) >= 2 and sorted(i for c in chunks for i in c) == list(range(120))
    for c in chunks:
        assert len(c) * one <= 3 << 30


class _FakePrep:
    def __init__(self, p=200):
        self.problem = type("X", (), dict(n=np.asarray([100, 100], dtype=np.int32), p=p, L=10, k=2, cost=200.0 * p * 10))()
        self.max_iter, self.min_tol = 500, 1e-4


def test_auto_chunking_gives_every_worker_a_chunk(monkeypatch):
    monkeypatch.setattr(config, "chunk_loci", 0)
    preps = [_FakePrep(200 + i % 17) for i in range(1300)]
    assert len(batch.plan_chunks(preps, n_workers=1)) == 1       # one worker: nothing to balance
    two = batch.plan_chunks(preps, n_workers=2)                  # 1300 // 296 = 4 chunks of 325
    assert sorted(len(c) for c in two) == [325, 325, 325, 325]
    eight = batch.plan_chunks(preps, n_workers=8)                # small chunks rather than idle GPUs
    assert len(eight) == 8 and sorted(i for c in eight for i in c) == list(range(1300))
    assert len(batch.plan_chunks(preps[:5], n_workers=8)) == 5
    # several workers per GPU (they take turns in the IBSS kernel): chunks are for pipelining big batches, a small
    # group stays whole on its device instead of running its loci one chunk after the other
    assert len(batch.plan_chunks(preps[:6], n_workers=3, n_devices=1)) == 1
    assert len(batch.plan_chunks(preps[:500], n_workers=3, n_devices=1)) == 1
    assert sorted(len(c) for c in batch.plan_chunks(preps[:1000], n_workers=3, n_devices=1)) == [333, 333, 334]
    assert sorted(len(c) for c in batch.plan_chunks(preps[:6], n_workers=6, n_devices=2)) == [3, 3]


@pytest.mark.parametrize("devices", [[0], [0, 1, 2]])
def test_run_sharded_defers_table_assembly_and_keeps_input_order(monkeypatch, devices):
    monkeypatch.setattr(config, "chunk_loci", 7)
    preps = [_FakePrep(100 + i) for i in range(40)]
    seen = []

    def runner(chunk, device, defer=False):
        seen.append((device, len(chunk), defer))
        done = [lambda p=p: ("fit", p.problem.p) for p in chunk]
        return done if defer else [f() for f in done]

    out = batch.run_sharded(preps, runner, devices)
    assert out == [("fit", 100 + i) for i in range(40)]
    assert all(d for _, _, d in seen) and sum(n for _, n, _ in seen) == 40

    def bad_runner(chunk, device, defer=False):
        def boom():
            raise RuntimeError("table assembly failed")
        return [boom for _ in chunk]

    with pytest.raises(RuntimeError, match="table assembly failed"):
        batch.run_sharded(preps, bad_runner, devices)


def test_c_abi_exports_every_declared_symbol(engine_lib):
    header = open(os.path.join(ROOT, "include", "sushie_b200.h")).read()
    declared = set(re.findall(r"\b(sb_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_capi.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(engine_lib, name), name
    assert engine_lib.sb_version() == 110
    # struct layouts on the Python side must match the C side (sizes are a cheap proxy)
    structs = [("sb_opts", _capi.SbOpts, 40), ("sb_result", _capi.SbResult, 96), ("sb_run_stats", _capi.SbRunStats, 72),
               ("sb_ind_problem", _capi.SbIndProblem, 128), ("sb_ss_problem", _capi.SbSsProblem, 96),
               ("sb_selection", _capi.SbSelection, 24)]
    for _, py_struct, size in structs:
        assert ctypes.sizeof(py_struct) == size
    # ... and the sizes a C compiler gives the header itself (the header is plain C: it must compile as such)
    import shutil
    import subprocess
    import tempfile

    gcc = shutil.which("gcc")
    if gcc:
        with tempfile.TemporaryDirectory() as tmp:
            src = os.path.join(tmp, "sizes.c")
            with open(src, "w") as fh:
                fh.write('#include <stdio.h>\n#include "sushie_b200.h"\nint main(void){printf("'
                         + " ".join(["%zu"] * len(structs)) + '\\n",'
                         + ",".join(f"sizeof({name})" for name, _, _ in structs) + ");return 0;}\n")
            exe = os.path.join(tmp, "sizes")
            subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
            got = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
        assert [int(v) for v in got] == [size for _, _, size in structs]


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under sushie_b200/ may reference it."""
    pkg = os.path.join(ROOT, "sushie_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), fn


@pytest.mark.parametrize("no_reorder", [True, False])
def test_staged_epilogue_equals_one_shot_assembly(no_reorder):
    """`StagedResult` (set selection -> purity for the whole chunk -> deferred tables) must give exactly what the
    one-shot `assemble_result` + `build_tables` path gives."""
    import types

    rng = np.random.default_rng(3)
    L, p, k = 4, 400, 3
    alpha = rng.dirichlet(np.ones(p) * 0.02, size=L)
    alpha[3] = rng.dirichlet(np.ones(p) * 5.0)  # a diffuse effect: its set is sub-sampled to max_select
    fit = types.SimpleNamespace(
        alpha=alpha, post_mean=rng.standard_normal((L, p, k)), post_mean_sq=rng.random((L, p, k, k)),
        weighted_sum_covar=np.stack([np.eye(k) * w for w in (0.3, 2.0, 1.0, 0.1)]), kl=rng.random(L),
        log_bf=rng.standard_normal((L, p)), effect_covar=np.stack([np.eye(k) * w for w in (0.3, 2.0, 1.0, 0.1)]),
        resid_var=np.ones(k), elbo=np.array([-np.inf, -3.0, -2.0]), elbo_increase=True,
    )
    ns = np.array([[100], [120], [90]])
    seen = []

    def min_abs(sets):
        seen.append([np.asarray(s).copy() for s in sets])
        return np.array([[0.2 + 0.1 * i + 0.01 * a for a in range(k)] for i in range(len(sets))])

    cs_args = dict(threshold=0.9, purity=0.35, purity_method="weighted", max_select=50, seed=7)
    one = _common.assemble_result(
        fit, None, ns, no_reorder, lambda a, lbf: cs_mod.build_tables(a, lbf, ns, min_abs, **cs_args))
    staged = _common.StagedResult(fit, None, ns, no_reorder, **cs_args)
    assert max(len(s) for s in staged.purity_sets) == 50
    two = staged.finisher(min_abs(staged.purity_sets))()
    assert all(np.array_equal(a, b) for a, b in zip(seen[0], seen[1]))
    pd_testing = pytest.importorskip("pandas.testing")
    pd_testing.assert_frame_equal(one.cs, two.cs)
    pd_testing.assert_frame_equal(one.alphas, two.alphas)
    np.testing.assert_array_equal(one.pip_all, two.pip_all)
    np.testing.assert_array_equal(one.pip_cs, two.pip_cs)
    np.testing.assert_array_equal(one.l_order, two.l_order)
    np.testing.assert_array_equal(one.posteriors.alpha, two.posteriors.alpha)
    with pytest.raises(ValueError, match="Invalid purity method"):
        _common.StagedResult(fit, None, ns, True, 0.9, 0.5, "median", 50, 7)
    # the tables exist when the finisher returns, unless the problem was marked `lazy_tables`
    assert two._tables is None
    lazy = _common.StagedResult(fit, None, ns, no_reorder, **cs_args, lazy_tables=True)
    three = lazy.finisher(min_abs(lazy.purity_sets))()
    assert three._tables is not None
    np.testing.assert_array_equal(three.pip_cs, two.pip_cs)  # (PIPs never wait for the tables)
    pd_testing.assert_frame_equal(three.alphas, two.alphas)
    assert three._tables is None


def test_plan_chunks_properties_under_random_inputs():
    """Property check (hypothesis): every problem lands in exactly one chunk, chunks never mix launch
    configurations or storage kinds, and explicit caps are respected, for any worker count."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")

    class P:
        def __init__(self, p, tol, hard):
            self.problem = type("X", (), dict(n=np.asarray([50, 60], dtype=np.int32), p=p, L=5, k=2, cost=110.0 * p * 5))()
            self.max_iter, self.min_tol, self.hard_calls = 500, tol, hard

    @hyp.settings(max_examples=300, deadline=None)
    @hyp.given(
        st.lists(st.tuples(st.integers(100, 3000), st.sampled_from([1e-3, 1e-4]), st.booleans()), min_size=1, max_size=80),
        st.integers(1, 8), st.integers(0, 9),
    )
    def check(items, n_workers, cap):
        preps = [P(*it) for it in items]
        chunks = batch.plan_chunks(preps, chunk_bytes=1 << 40, chunk_loci=cap, n_workers=n_workers)
        assert sorted(i for c in chunks for i in c) == list(range(len(preps)))
        for c in chunks:
            assert len({(preps[i].min_tol, preps[i].hard_calls) for i in c}) == 1
            if cap:
                assert len(c) <= cap
        if not cap:
            groups = {(p.min_tol, p.hard_calls) for p in preps}
            sizes = {g: sum(1 for p in preps if (p.min_tol, p.hard_calls) == g) for g in groups}
            for g, n in sizes.items():  # with several workers every group is cut into at least min(n, workers) chunks
                n_chunks = sum(1 for c in chunks if (preps[c[0]].min_tol, preps[c[0]].hard_calls) == g)
                assert n_chunks >= (min(n, n_workers) if n_workers > 1 else 1)

    check()


# ---- device-side set selection, effect order at download time, row views, lazy tables ---------------------------


def _device_selection_emulated(alpha, threshold):
    """What `sb_batch_select` computes, in numpy: descending alpha with ties by ascending SNP index, the running
    sum accumulated left to right, the set size."""
    L, p = alpha.shape
    order = np.stack([np.lexsort((np.arange(p), -alpha[l])) for l in range(L)]).astype(np.int32)
    csum = np.cumsum(np.take_along_axis(alpha, order.astype(np.int64), axis=1), axis=1)
    n_rows = np.count_nonzero(csum < threshold, axis=1)
    sizes = np.where(n_rows == p, n_rows, n_rows + 1).astype(np.int32)
    return order, csum, sizes


@pytest.mark.parametrize("permuted", [False, True])
def test_selection_from_device_equals_host_selection(permuted):
    rng = np.random.default_rng(11)
    L, p = 5, 700
    alpha = rng.dirichlet(np.ones(p) * 0.03, size=L)
    alpha[2] = rng.dirichlet(np.ones(p) * 8.0)  # diffuse: sub-sampled
    alpha[4] = 0.0                              # the "initial state returned" case: every SNP in the set
    l_order = np.array([3, 0, 4, 2, 1]) if permuted else np.arange(L)
    host = cs_mod.select_sets(alpha[l_order], 0.9, 60, 5)
    dev = cs_mod.selection_from_device(*_device_selection_emulated(alpha, 0.9), l_order, 60, 5)
    assert dev[2] == host[2]
    for l in range(L):
        if l_order[l] == 4:
            continue  # all-equal alphas: the tie order is the sort's business
        np.testing.assert_array_equal(dev[0][l], host[0][l])
        np.testing.assert_array_equal(dev[1][l], host[1][l])
        np.testing.assert_array_equal(dev[3][l], host[3][l])
    assert dev[2][list(l_order).index(4)] == p and len(dev[3][list(l_order).index(4)]) == 60


def test_effect_orders_equal_reorder_l():
    import types

    rng = np.random.default_rng(5)
    fits, want = [], []
    for L, k in ((10, 3), (10, 3), (4, 2), (10, 3), (6, 1)):
        a = rng.standard_normal((L, k, k))
        wsc = a @ np.transpose(a, (0, 2, 1))
        wsc[1] = wsc[0]  # an exact tie: the stable sort keeps index order
        fits.append(types.SimpleNamespace(weighted_sum_covar=wsc))
        post = _common.Posterior(np.zeros((L, 3)), np.zeros((L, 3, k)), np.zeros((L, 3, k, k)), wsc, np.zeros(L),
                                 np.zeros((L, 3)))
        want.append(_common._reorder_l(_common.Prior(None, None, np.zeros((L, k, k))), post)[2])
    fits[3].weighted_sum_covar = np.full((10, 3, 3), np.nan)  # non-finite fit: every norm NaN -> index order
    got = _common.effect_orders(fits, [False, True, False, False, False])
    assert got[1] is None
    for i in (0, 2, 4):
        np.testing.assert_array_equal(got[i], want[i])
        assert got[i].dtype == np.int32
    np.testing.assert_array_equal(got[3], np.arange(10))


def test_sushie_result_behaves_like_the_reference_named_tuple():
    import pickle

    calls = []

    def tables():
        calls.append(1)
        return pd.DataFrame({"CSIndex": [1]}), pd.DataFrame({"SNPIndex": [0, 1]})

    pri = _common.Prior(np.ones(2) / 2, np.ones((1, 1)), np.ones((1, 1, 1)))
    post = _common.Posterior(*[np.zeros(1)] * 6)
    res = _common.SushieResult(pri, post, np.zeros(2), np.zeros(2), None, None, np.array([[5]]), np.array([-np.inf, -1.0]),
                               True, np.arange(1), tables=tables)
    assert res._fields == ("priors", "posteriors", "pip_all", "pip_cs", "cs", "alphas", "sample_size", "elbo",
                           "elbo_increase", "l_order")
    assert not calls and res.priors is pri and len(res) == 10
    assert list(res.cs.columns) == ["CSIndex"] and list(res.alphas.columns) == ["SNPIndex"] and calls == [1]
    priors, posteriors, pip_all, pip_cs, cs, alphas, ns, elbo, inc, l_order = res  # unpacks like the named tuple
    assert cs is res.cs and res[4] is cs and res[-1] is l_order and calls == [1]
    assert set(res._asdict()) == set(res._fields)
    twin = res._replace(elbo_increase=False)
    assert twin.elbo_increase is False and twin.cs is cs and res.elbo_increase is True
    back = pickle.loads(pickle.dumps(res))
    pd.testing.assert_frame_equal(back.alphas, res.alphas)
    with pytest.raises(ValueError):
        res._replace(nope=1)
    with pytest.raises(TypeError):
        _common.SushieResult(pri, post, np.zeros(2), np.zeros(2))
    eager = _common.SushieResult(pri, post, np.zeros(2), np.zeros(2), cs, alphas, ns, elbo, True, l_order)
    assert eager.cs is cs


def test_block_built_frames_equal_the_dict_constructor(monkeypatch):
    try:
        import pandas.api.internals as internals
        assert cs_mod._FRAME_FROM_BLOCKS is getattr(internals, "create_dataframe_from_blocks", None)
    except ImportError:
        assert cs_mod._FRAME_FROM_BLOCKS is None
    rng = np.random.default_rng(2)
    L, p = 3, 41
    alpha = rng.dirichlet(np.ones(p) * 0.05, size=L)
    log_bf = rng.standard_normal((L, p))
    selection = cs_mod.select_sets(alpha, 0.9, 10, 3)
    purities, kept = np.array([0.9, 0.2, 0.7]), np.array([True, False, True])
    pip_all, pip_cs = utils.make_pip(alpha), utils.make_pip(alpha[[0, 2]])

    # the reference's schema written out column by column with the dict constructor
    cols = {"SNPIndex": np.arange(p)}
    rows = []
    for l in range(L):
        order, csum, size = selection[0][l], selection[1][l], selection[2][l]
        in_cs = np.zeros(p, dtype=np.int64)
        in_cs[order[:size]] = 1
        cols[f"alpha_l{l + 1}"] = alpha[l]
        cols[f"in_cs_l{l + 1}"] = in_cs
        cols[f"purity_l{l + 1}"] = np.full(p, purities[l])
        cols[f"kept_l{l + 1}"] = np.full(p, int(kept[l]), dtype=np.int64)
        cols[f"log_bf_l{l + 1}"] = log_bf[l]
        if kept[l]:
            rows.append(pd.DataFrame({"CSIndex": np.full(size, l + 1, dtype=np.int64), "SNPIndex": order[:size].astype(np.int64),
                                      "alpha": alpha[l][order[:size]], "c_alpha": csum[:size]}))
    cols["pip_all"], cols["pip_cs"] = pip_all, pip_cs
    want_alphas = pd.DataFrame(cols)
    want_cs = pd.concat(rows, ignore_index=True)
    want_cs["pip_all"] = pip_all[want_cs.SNPIndex.values]
    want_cs["pip_cs"] = pip_cs[want_cs.SNPIndex.values]

    for use_blocks in (True, False):  # the block path and the older-pandas path give the same frames
        if not use_blocks:
            monkeypatch.setattr(cs_mod, "_FRAME_FROM_BLOCKS", None)
        elif cs_mod._FRAME_FROM_BLOCKS is None:
            continue
        cs, alphas = cs_mod.frames(alpha, log_bf, selection, purities, kept, pip_all, pip_cs)
        pd.testing.assert_frame_equal(alphas, want_alphas)  # values, dtypes, column order, index
        pd.testing.assert_frame_equal(cs, want_cs)
        alphas["extra"] = 1.0  # the frame must behave like any other afterwards
        assert alphas.shape == (p, 5 * L + 4) and alphas.iloc[3, 1] == alpha[0, 3]
    # no kept set: an empty cs frame with the right columns and dtypes
    cs, _ = cs_mod.frames(alpha, log_bf, selection, purities, np.zeros(L, dtype=bool), pip_all, np.zeros(p))
    pd.testing.assert_frame_equal(cs, pd.DataFrame({"CSIndex": np.zeros(0, dtype=np.int64), "SNPIndex": np.zeros(0, dtype=np.int64),
                                                   "alpha": np.zeros(0), "c_alpha": np.zeros(0), "pip_all": np.zeros(0),
                                                   "pip_cs": np.zeros(0)}))
    assert cs_mod._alphas_layout(L)[0] is cs_mod._alphas_layout(L)[0]  # one shared column index per effect count


def test_row_views_reach_the_engine_as_one_shared_matrix():
    """A gene's full fit and its cross-validation folds (row views) must hand the engine ONE genotype pointer per
    ancestry plus row lists -- that is what makes it upload the matrix once."""
    import ctypes as C

    from sushie_b200 import cli, infer, utils

    rng = np.random.default_rng(8)
    n, p = 60, 130
    G = [rng.integers(0, 3, size=(n, p)).astype(np.int8) for _ in range(2)]
    ys = [rng.standard_normal(n) for _ in range(2)]
    folds = cli._prepare_cv(list(G), list(ys), 5, 12345)
    assert all(isinstance(x, utils.RowView) for f in folds for x in f.train_geno)
    np.testing.assert_array_equal(np.asarray(folds[1].train_geno[0]), G[0][folds[1].train_geno[0].rows])
    cache = infer._HardCallCache()
    preps = [infer._prepare_kw(Xs=G, ys=ys, hard_cache=cache)]
    preps += [infer._prepare_kw(Xs=f.train_geno, ys=f.train_pheno, hard_cache=cache) for f in folds]
    assert all(pr.hard_calls == "b2" for pr in preps)
    # what the CLI passes for the folds (`_cv_specs`): their tables are not assembled unless somebody reads them
    import argparse
    ns_args = argparse.Namespace(L=3, no_scale=False, no_regress=False, no_update=False, resid_var=None, effect_var=None,
                                 rho=None, max_iter=50, min_tol=1e-3, threshold=0.95, purity=0.5, max_select=250,
                                 seed=12345, cv_num=5)
    specs = cli._cv_specs(ns_args, folds, None)
    assert len(specs) == 5 and all(sp["lazy_tables"] for sp in specs)
    assert infer._prepare_kw(**specs[0]).lazy_tables and not preps[0].lazy_tables
    full = preps[0].problem
    assert full.rows is None and [x.ctypes.data for x in full.Xs] == [g.ctypes.data for g in G]
    for pr, fold in zip(preps[1:], folds):
        prob = pr.problem
        assert [x.ctypes.data for x in prob.Xs] == [g.ctypes.data for g in G]  # same memory as the full fit
        assert prob.n.tolist() == [48, 48] and prob._src_rows.tolist() == [n, n]
        s = _capi.SbIndProblem()
        prob.fill(s)
        for a in range(2):
            assert [s.x_rows[a][r] for r in range(48)] == fold.train_geno[a].rows.tolist()
            assert s.x_src_rows[a] == n
    s = _capi.SbIndProblem()
    full.fill(s)
    assert not s.x_rows and not s.x_src_rows
    # float genotypes that are hard calls in disguise: one int8 copy shared by every view
    Gf = [g.astype(np.float64) for g in G]
    cache = infer._HardCallCache()
    a = infer._prepare_kw(Xs=Gf, ys=ys, hard_cache=cache).problem
    b = infer._prepare_kw(Xs=[utils.RowView(g, np.arange(10, 50)) for g in Gf], ys=[y[10:50] for y in ys], hard_cache=cache).problem
    assert a.x_dtype == _capi.SB_I8 and [x.ctypes.data for x in a.Xs] == [x.ctypes.data for x in b.Xs]
    # dosages: dense storage, the view is still gathered on the device from the caller's matrix
    Gd = [g + 0.25 for g in Gf]
    c = infer._prepare_kw(Xs=[utils.RowView(g, np.arange(10, 50)) for g in Gd], ys=[y[10:50] for y in ys]).problem
    assert c.x_dtype == _capi.SB_F64 and c.rows is not None and [x.ctypes.data for x in c.Xs] == [g.ctypes.data for g in Gd]


def test_distributed_queue_looks_one_chunk_ahead():
    """`run_distributed`: every item exactly once; a worker pulls from the queue only when no chunk of its rank is
    still waiting for the GPU (the gate is handed on by the hook `run_chunk` calls on entering the IBSS kernel)."""
    import threading
    import time

    gpu = threading.Lock()
    state = {"waiting": 0, "max_waiting": 0, "pulled": 0}
    guard = threading.Lock()

    def process(ids):
        with guard:
            state["waiting"] += 1
            state["pulled"] += 1
            state["max_waiting"] = max(state["max_waiting"], state["waiting"])
        time.sleep(0.002)  # prologue + upload
        with gpu:
            with guard:
                state["waiting"] -= 1
            _capi.run_hooks.on_start()
            time.sleep(0.01)  # the kernel
        time.sleep(0.004)  # epilogue
        return [i * i for i in ids]

    costs = [float(c) for c in np.random.default_rng(0).integers(1, 50, size=90)]
    records, trace = batch.run_distributed(costs, process, workers_per_rank=3, min_items=4, max_items=8)
    assert sorted(records) == [(i, i * i) for i in range(90)]
    assert sum(n for _, _, n in trace) == 90 and state["pulled"] == len(trace)
    assert state["max_waiting"] == 1
    assert getattr(_capi.run_hooks, "on_start", None) is None  # (the calling thread never had one)

    def failing(ids):  # a chunk that dies before the GPU must not wedge the other workers
        if 0 in ids:
            raise RuntimeError("bad gene")
        return process(ids)

    with pytest.raises(RuntimeError, match="bad gene"):
        batch.run_distributed(costs, failing, workers_per_rank=3, min_items=4, max_items=8)


def test_run_chunk_epilogue_with_a_stand_in_engine():
    """`infer.run_chunk` end to end on the CPU with a stand-in for the device batch: effect order from the small
    download, posteriors delivered in that order, set selection handed over in engine order, one purity call per
    chunk, tables built eagerly or on demand -- and the result equal to the host-only `build_tables` path."""
    import types

    from sushie_b200 import infer

    L, p, k = 4, 300, 2
    alpha = np.random.default_rng(1).dirichlet(np.ones(p) * 0.05, size=L)
    wsc = np.stack([np.eye(k) * w for w in (0.3, 2.0, 1.0, 0.1)])
    calls = []

    class Batch:
        def __init__(self, n):
            self.n = n

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

        def run(self):
            calls.append("run")

        def fetch(self, effect_orders=None, small_only=False):
            calls.append("fetch_small" if small_only else "fetch")
            fits = []
            for i in range(self.n):
                order = np.arange(L) if effect_orders is None or effect_orders[i] is None else effect_orders[i]
                f = _capi.RawFit(k, p, L, 20, small_only)
                f.weighted_sum_covar[:] = wsc[order]
                if not small_only:
                    f.alpha[:] = alpha[order]
                    f.log_bf[:] = np.log(alpha[order] + 1e-300)
                    f.effect_covar[:] = wsc[order]
                    f.post_mean[:], f.post_mean_sq[:], f.kl[:], f.resid_var[:] = 0.0, 0.0, 0.0, 1.0
                    if effect_orders is not None and effect_orders[i] is not None:
                        f.l_order = np.asarray(order, dtype=np.int32)
                f.n_iter, f.elbo, f.elbo_increase = 2, np.array([-np.inf, -2.0, -1.0]), True
                fits.append(f)
            return fits

        def select(self, thresholds):
            calls.append("select")
            return [_device_selection_emulated(alpha, t) for t in thresholds]

        def purity_many(self, sets_per_problem):
            calls.append("purity")
            return [np.full((len(sets), k), 0.9) for sets in sets_per_problem]

    engine = types.SimpleNamespace(device=0, create_batch=lambda problems, opts: Batch(len(problems)))
    ns = np.array([[100], [100]])
    cs_args = dict(threshold=0.9, purity=0.5, purity_method="weighted", max_select=50, seed=1)
    preps = [infer._Prepared(types.SimpleNamespace(), None, ns, 20, 1e-4, cs_args, no_reorder, lazy_tables=lazy)
             for no_reorder, lazy in ((False, False), (False, True), (True, False))]
    out = infer.run_chunk(engine, preps, None, infer._stage, False)
    assert calls == ["run", "select", "fetch_small", "fetch", "purity"]
    assert out[0]._tables is None and out[1]._tables is not None and out[2]._tables is None
    np.testing.assert_array_equal(out[0].l_order, [1, 2, 0, 3])
    np.testing.assert_array_equal(out[2].l_order, np.arange(L))
    for res in out:
        lo = np.asarray(res.l_order)
        want = cs_mod.build_tables(alpha[lo], np.log(alpha[lo] + 1e-300), ns, lambda sets: np.full((len(sets), k), 0.9),
                                   0.9, 0.5, "weighted", 50, 1)
        pd.testing.assert_frame_equal(res.cs, want[0])
        pd.testing.assert_frame_equal(res.alphas, want[1])
        np.testing.assert_array_equal(res.pip_all, want[2])
        np.testing.assert_array_equal(res.posteriors.weighted_sum_covar, wsc[lo])
    deferred = infer.run_chunk(engine, preps[:1], None, infer._stage, True)
    assert callable(deferred[0]) and deferred[0]().cs.equals(out[0].cs)
