#!/usr/bin/env python
"""Generate the committed golden fixtures under ``tests/golden/``.

Run in the build container only (it reads ``/root/reference``, which does not exist on the GPU box):

    python tests/golden/make_golden.py [ind] [ss] [syn] [headline] [config5] [config4]   (default: all)
    python tests/golden/make_golden.py --oracle ...      (outputs of oracle/ instead; for diffing only)

**What produces the numbers.**  By default every expected output below comes from the REFERENCE'S OWN SOURCE --
``/root/reference/sushie/infer.py`` / ``infer_ss.py`` / ``utils.py`` / ``cli.py`` imported unmodified and executed
through ``tests/refshim`` (stand-in ``jax`` / ``equinox`` modules backed by numpy fp64; see its docstring for the
primitive-by-primitive substitution).  The real jax wheels are not installable here, so this is the strongest pin
available: the reference's algorithm text runs line by line; only the array library underneath differs.

Outputs
  * ``example_ind.npz``  -- config 1: bundled EUR+AFR VCF + phenotypes + covariates after the CLI's QC order
    (``cli.py:897-1231``), as raw int8 REF-allele counts (``io.py:243-255``) plus ``infer_sushie`` outputs at
    ``min_tol`` 1e-3 (CLI default) and 1e-4 (API default).
  * ``example_ss.npz``   -- config 2: bundled ``*.ld`` + ``*.gwas`` z-scores for EUR+AFR and EUR+AFR+EAS with
    ns = 489/639/481, plus ``infer_sushie_ss`` outputs.
  * ``synthetic_small.npz`` -- seeded synthetic loci (k in 1..4, option mixes) with outputs.
  * ``headline_genes.npz`` -- config 3: genes 0..3 of ``bench.make_gene`` (k=3, n=500, p=2000, L=10; 0/1/2/3 causal
    SNPs) fitted TO CONVERGENCE at ``min_tol`` 1e-4 and 1e-3 (inputs are regenerated from the seed by the tests).
  * ``config5_gene.npz`` -- config 5: gene 7 of ``scripts/bench_cv`` (ragged p = 1138): the reference's
    ``cli._prepare_cv`` row shuffles / folds, the full fit, the 5 fold fits and ``cli._run_cv``'s scores.
  * ``config4_summary.npz`` -- config 4: ``scripts/bench_ss.make_locus`` (k=3, p=10 000 LD), 3 fixed iterations.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..")))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..", "scripts")))

from oracle import credible_sets, ibss  # noqa: E402

DATA = "/root/reference/data"
USE_ORACLE = "--oracle" in sys.argv
_REF = None


def reference():
    """The reference's modules, executed through the numpy-backed jax stand-in (tests/refshim)."""
    global _REF
    if _REF is None:
        import refshim

        _REF = refshim.load_reference(("log", "utils", "infer", "infer_ss", "io", "cli"))
    return _REF


class _RefFit:
    """Adapter: the reference's ``SushieResult`` under the attribute names ``pack_fit`` reads."""

    def __init__(self, res):
        self.n_iter = len(res.elbo) - 1
        self.elbo = np.asarray(res.elbo, dtype=float)
        self.elbo_increase = bool(res.elbo_increase)
        self.l_order = np.asarray(res.l_order)
        self.posteriors = res.posteriors
        self.priors = res.priors
        self.sample_size = np.asarray(res.sample_size)


def read_vcf_text(path):
    """Plain-text VCF -> (bim, iids, bed[n, p]) with the reference's coding."""
    rows, bim = [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith("##"):
                continue
            parts = line.rstrip("\n").split("\t")
            if line.startswith("#CHROM"):
                iids = parts[9:]
                continue
            chrom, pos, rsid, ref, alt = parts[0], int(parts[1]), parts[2], parts[3], parts[4]
            g = np.empty(len(parts) - 9)
            for i, field in enumerate(parts[9:]):
                gt = field.split(":")[0].replace("|", "/").split("/")
                if "." in gt:
                    g[i] = np.nan
                else:
                    g[i] = 2 - (int(gt[0] != "0") + int(gt[1] != "0"))
            rows.append(g)
            bim.append([chrom, rsid, pos, alt.split(",")[0], ref])
    bim = pd.DataFrame(bim, columns=["chrom", "snp", "pos", "a0", "a1"])
    return bim, list(iids), np.asarray(rows).T


def qc_individual(ancestries):
    raw = []
    for anc in ancestries:
        bim, iids, bed = read_vcf_text(f"{DATA}/vcf/{anc}.vcf")
        pheno = pd.read_csv(f"{DATA}/{anc}.pheno", sep="\t", header=None, names=["iid", "pheno"])
        covar = pd.read_csv(f"{DATA}/{anc}.covar", sep="\t", header=None)
        covar = covar.rename(columns={0: "iid"})
        # impute with column means, MAF filter, de-dup (cli.py:113-252)
        keep = ~np.all(np.isnan(bed), axis=0)
        bim, bed = bim[keep].reset_index(drop=True), bed[:, keep]
        col_mean = np.nanmean(bed, axis=0)
        nan_r, nan_c = np.where(np.isnan(bed))
        bed[nan_r, nan_c] = col_mean[nan_c]
        maf = np.mean(bed, axis=0) / 2
        maf = np.where(maf > 0.5, 1 - maf, maf)
        sel = maf >= 0.01
        bim, bed = bim[sel].reset_index(drop=True), bed[:, sel]
        dup = bim.snp.duplicated(keep="first").values
        bim, bed = bim[~dup].reset_index(drop=True), bed[:, ~dup]
        fam = pd.DataFrame({"iid": iids}).reset_index().rename(columns={"index": "famIDX"})
        pheno = pheno.reset_index().rename(columns={"index": "phenoIDX"})
        covar = covar.reset_index().rename(columns={"index": "covarIDX"})
        common = fam.merge(pheno[["phenoIDX", "iid"]], on="iid").merge(covar[["covarIDX", "iid"]], on="iid")
        raw.append((bim.reset_index().rename(columns={"index": "bimIDX"}), bed, pheno, covar, common))
    snps = raw[0][0].rename(columns=lambda c: c if c in ("chrom", "snp") else c + "_1")
    for i in range(1, len(raw)):
        other = raw[i][0].rename(columns=lambda c: c if c in ("chrom", "snp") else f"{c}_{i + 1}")
        snps = snps.merge(other, how="inner", on=["chrom", "snp"]).reset_index(drop=True)
    flips = []
    for i in range(1, len(raw)):
        ok = (snps.a0_1 == snps[f"a0_{i + 1}"]) & (snps.a1_1 == snps[f"a1_{i + 1}"])
        fl = (snps.a0_1 == snps[f"a1_{i + 1}"]) & (snps.a1_1 == snps[f"a0_{i + 1}"])
        snps = snps[ok | fl].reset_index(drop=True)
    ambiguous = (snps.a0_1 + snps.a1_1).isin(["AT", "TA", "CG", "GC"])
    snps = snps[~ambiguous].reset_index(drop=True)
    for i in range(1, len(raw)):
        fl = (snps.a0_1 == snps[f"a1_{i + 1}"]) & (snps.a1_1 == snps[f"a0_{i + 1}"])
        flips.append(np.where(fl.values)[0])
    Xs, ys, cvs = [], [], []
    for i, (bim, bed, pheno, covar, common) in enumerate(raw):
        g = bed[common.famIDX.values, :][:, snps[f"bimIDX_{i + 1}"].values]
        if i > 0 and len(flips[i - 1]):
            g[:, flips[i - 1]] = 2 - g[:, flips[i - 1]]
        Xs.append(g)
        ys.append(pheno["pheno"].values[common.phenoIDX.values])
        cvs.append(covar.iloc[common.covarIDX.values, 2:].values.astype(float))
    return np.asarray(snps.snp.tolist(), dtype=str), Xs, ys, cvs


def pack_fit(prefix, fit, cs_out, out, slim=False):
    """``slim``: large loci keep alpha and the summed weights only (no per-effect means / log Bayes factors)."""
    cs, alphas, pip_all, pip_cs = cs_out
    out[f"{prefix}_n_iter"] = fit.n_iter
    out[f"{prefix}_elbo"] = np.asarray(fit.elbo, dtype=float)
    out[f"{prefix}_elbo_increase"] = bool(fit.elbo_increase)
    out[f"{prefix}_l_order"] = np.asarray(fit.l_order)
    out[f"{prefix}_alpha"] = np.asarray(fit.posteriors.alpha)
    out[f"{prefix}_weights"] = np.asarray(fit.posteriors.post_mean).sum(axis=0)
    if not slim:
        out[f"{prefix}_post_mean"] = np.asarray(fit.posteriors.post_mean)
        out[f"{prefix}_log_bf"] = np.asarray(fit.posteriors.log_bf)
    out[f"{prefix}_wsc"] = np.asarray(fit.posteriors.weighted_sum_covar)
    out[f"{prefix}_kl"] = np.asarray(fit.posteriors.kl)
    out[f"{prefix}_effect_covar"] = np.asarray(fit.priors.effect_covar)
    out[f"{prefix}_resid_var"] = np.asarray(fit.priors.resid_var)
    out[f"{prefix}_pip_all"] = np.asarray(pip_all)
    out[f"{prefix}_pip_cs"] = np.asarray(pip_cs)
    out[f"{prefix}_cs_index"] = cs.CSIndex.values.astype(np.int64)
    out[f"{prefix}_cs_snp"] = cs.SNPIndex.values.astype(np.int64)
    purity_cols = [c for c in alphas.columns if c.startswith("purity_l")]
    out[f"{prefix}_purity"] = alphas[purity_cols].iloc[0].values.astype(float)
    kept_cols = [c for c in alphas.columns if c.startswith("kept_l")]
    out[f"{prefix}_kept"] = alphas[kept_cols].iloc[0].values.astype(np.int64)


def fit_and_cs_ind(Xs, ys, covar, **kw):
    if not USE_ORACLE:
        res = reference().infer.infer_sushie(
            [np.array(x, dtype=float) for x in Xs], [np.array(y, dtype=float) for y in ys],
            None if covar is None else [np.array(c, dtype=float) for c in covar], **kw)
        return _RefFit(res), (res.cs, res.alphas, np.asarray(res.pip_all), np.asarray(res.pip_cs))
    cs_kw = {k: kw.pop(k) for k in list(kw) if k in ("threshold", "purity", "purity_method", "max_select", "seed")}
    fit = ibss.fit_individual(Xs, ys, covar, **kw)
    cs_args = dict(threshold=0.95, purity=0.5, purity_method="weighted", max_select=250, seed=12345)
    cs_args.update(cs_kw)
    cs_out = credible_sets.make_cs(fit.posteriors.alpha, fit.posteriors.log_bf, fit.sample_size, Xs=fit.Xs, **cs_args)
    return fit, cs_out


def fit_and_cs_ss(lds, ns, zs, **kw):
    if not USE_ORACLE:
        res = reference().infer_ss.infer_sushie_ss(np.array(lds, dtype=float), np.array(ns), np.array(zs, dtype=float), **kw)
        return _RefFit(res), (res.cs, res.alphas, np.asarray(res.pip_all), np.asarray(res.pip_cs))
    fit = ibss.fit_summary(lds, ns, zs, **kw)
    cs_out = credible_sets.make_cs(
        fit.posteriors.alpha, fit.posteriors.log_bf, fit.sample_size, lds=fit.lds,
        threshold=0.95, purity=0.5, purity_method="weighted", max_select=250, seed=12345,
    )
    return fit, cs_out


def golden_example_ind():
    snp, Xs, ys, cvs = qc_individual(["EUR", "AFR"])
    out = {"snp": snp}
    for i, (X, y, c) in enumerate(zip(Xs, ys, cvs)):
        assert np.all(X == np.round(X)), "hard calls expected"
        out[f"X{i}"] = X.astype(np.int8)
        out[f"y{i}"] = y
        out[f"covar{i}"] = c
    for tag, tol in (("tol3", 1e-3), ("tol4", 1e-4)):
        fit, cs_out = fit_and_cs_ind(Xs, ys, cvs, L=10, min_tol=tol)
        pack_fit(tag, fit, cs_out, out)
        top = np.argsort(-cs_out[2])[:2]
        print(f"example_ind {tag}: n_iter={fit.n_iter} p={len(snp)} top PIP SNPs={snp[top]} pips={cs_out[2][top]}")
    np.savez_compressed(os.path.join(HERE, "example_ind.npz"), **out)


def load_ss(ancestries):
    lds, zs, snps = [], [], None
    tables = []
    for anc in ancestries:
        ld = pd.read_csv(f"{DATA}/{anc}.ld", sep="\t")
        gw = pd.read_csv(f"{DATA}/{anc}.gwas", sep="\t")
        tables.append((ld, gw))
        ids = set(ld.columns) & set(gw.snp)
        snps = ids if snps is None else snps & ids
    order = [s for s in tables[0][0].columns if s in snps]
    ref = tables[0][1].set_index("snp").loc[order]
    for ld, gw in tables:
        cols = list(ld.columns)
        idx = [cols.index(s) for s in order]
        mat = ld.values[np.ix_(idx, idx)].astype(float)
        g = gw.set_index("snp").loc[order]
        z = g.z.values.astype(float)
        flip = (g.a0.values == ref.a1.values) & (g.a1.values == ref.a0.values)
        sign = np.where(flip, -1.0, 1.0)
        lds.append(mat * sign[:, None] * sign[None, :])
        zs.append(z * sign)
    return np.asarray(order), np.stack(lds), np.stack(zs)


def golden_example_ss():
    out = {}
    n_all = {"EUR": 489, "AFR": 639, "EAS": 481}
    for tag, ancs in (("k2", ["EUR", "AFR"]), ("k3", ["EUR", "AFR", "EAS"])):
        snp, lds, zs = load_ss(ancs)
        ns = np.array([[n_all[a]] for a in ancs])
        out[f"{tag}_snp"], out[f"{tag}_lds"], out[f"{tag}_zs"], out[f"{tag}_ns"] = snp, lds, zs, ns
        for ttag, tol in (("tol3", 1e-3), ("tol4", 1e-4)):
            fit, cs_out = fit_and_cs_ss(lds, ns, zs, L=10, min_tol=tol)
            pack_fit(f"{tag}_{ttag}", fit, cs_out, out)
            top = np.argsort(-cs_out[2])[:2]
            print(f"example_ss {tag} {ttag}: n_iter={fit.n_iter} p={len(snp)} top={snp[top]} pips={cs_out[2][top]}")
    np.savez_compressed(os.path.join(HERE, "example_ss.npz"), **out)


def synth_locus(seed, k, n, p, n_causal, h2=0.3):
    """Small seeded locus in the style of SURVEY.md section 8(d) config 3."""
    rng = np.random.default_rng(seed)
    rho_a = [0.90, 0.80, 0.85, 0.7][:k]
    maf = rng.uniform(0.05, 0.5, size=p)
    Xs = []
    from scipy.stats import norm

    thr = norm.ppf(1 - maf)
    for a in range(k):
        g = np.zeros((n[a], p))
        for _ in range(2):
            e = rng.standard_normal((n[a], p))
            z = np.empty_like(e)
            z[:, 0] = e[:, 0]
            c = np.sqrt(1 - rho_a[a] ** 2)
            for j in range(1, p):
                z[:, j] = rho_a[a] * z[:, j - 1] + c * e[:, j]
            g += z > thr
        for j in np.where(g.std(axis=0) == 0)[0]:
            g[0, j] = 1.0 if g[0, j] == 0 else g[0, j] - 1
        Xs.append(g)
    causal = rng.choice(p, size=n_causal, replace=False) if n_causal else np.array([], dtype=int)
    cov = 0.8 * np.ones((k, k)) + 0.2 * np.eye(k)
    b = rng.multivariate_normal(np.zeros(k), cov, size=n_causal) if n_causal else np.zeros((0, k))
    ys = []
    for a in range(k):
        Xstd = (Xs[a] - Xs[a].mean(0)) / Xs[a].std(0)
        gval = Xstd[:, causal] @ b[:, a] if n_causal else np.zeros(n[a])
        if n_causal:
            s2e = np.var(gval) * (1 / h2 - 1)
            ys.append(gval + rng.standard_normal(n[a]) * np.sqrt(s2e))
        else:
            ys.append(rng.standard_normal(n[a]))
    return Xs, ys, causal


def golden_synthetic():
    out = {}
    cases = [
        # tag, seed, k, n, p, n_causal, kwargs
        ("k1", 101, 1, [120], 150, 1, dict(L=5)),
        ("k2", 102, 2, [100, 140], 200, 2, dict(L=5)),
        ("k3", 103, 3, [90, 110, 100], 256, 2, dict(L=6)),
        ("k3null", 104, 3, [80, 80, 80], 160, 0, dict(L=4)),
        ("k2noupd", 105, 2, [100, 120], 180, 1, dict(L=4, no_update=True)),
        ("k2rho", 106, 2, [100, 120], 180, 1, dict(L=4, no_update=True, rho=[0.5])),
        ("k2var", 107, 2, [100, 120], 180, 1, dict(L=4, no_update=True, effect_var=[0.01, 0.02])),
        ("k2noscale", 108, 2, [100, 120], 180, 1, dict(L=4, no_scale=True)),
        ("k4", 109, 4, [70, 80, 90, 100], 130, 2, dict(L=4)),
    ]
    for tag, seed, k, n, p, nc, kw in cases:
        Xs, ys, causal = synth_locus(seed, k, n, p, nc)
        for i in range(k):
            out[f"{tag}_X{i}"] = Xs[i].astype(np.int8)
            out[f"{tag}_y{i}"] = ys[i]
        out[f"{tag}_causal"] = causal
        fit, cs_out = fit_and_cs_ind(Xs, ys, None, min_tol=1e-4, **kw)
        pack_fit(tag, fit, cs_out, out)
        print(f"synthetic {tag}: n_iter={fit.n_iter} inc={fit.elbo_increase} causal={causal} "
              f"cs_snps={sorted(set(cs_out[0].SNPIndex.values.tolist()))}")
    np.savez_compressed(os.path.join(HERE, "synthetic_small.npz"), **out)


# ---- BASELINE.json configs 3, 5, 4 at their real shapes ------------------------------------------------


def _headline_job(args):
    import warnings

    warnings.filterwarnings("ignore")
    g, tol = args
    import bench

    Xs, ys = bench.make_gene(g)
    fit, cs_out = fit_and_cs_ind([x.astype(float) for x in Xs], ys, None, L=bench.L_EFF, min_tol=tol)
    out = {}
    pack_fit(f"g{g}_{'tol4' if tol == 1e-4 else 'tol3'}", fit, cs_out, out, slim=True)
    return g, tol, fit.n_iter, out


def golden_headline(genes=(0, 1, 2, 3)):
    import multiprocessing as mp

    jobs = [(g, tol) for g in genes for tol in (1e-4, 1e-3)]
    out = {"genes": np.asarray(genes)}
    with mp.get_context("fork").Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
        for g, tol, n_iter, part in pool.imap_unordered(_headline_job, jobs):
            out.update(part)
            print(f"headline gene {g} min_tol={tol}: n_iter={n_iter}", flush=True)
    np.savez_compressed(os.path.join(HERE, "headline_genes.npz"), **out)


def golden_config5(gene=7, cv_num=5, seed=12345):
    import types

    import bench
    import bench_cv

    ref = reference()
    p = bench_cv.gene_width(gene, 500, 4000)
    Xs, ys = bench.make_gene(gene, bench.N_IND, p)
    out = {"gene": gene, "p": p, "cv_num": cv_num, "seed": seed}
    # the row shuffles themselves: run the reference's splitter on row-index "genotypes"
    idx_folds = ref.cli._prepare_cv([np.arange(x.shape[0])[:, None] for x in Xs], [np.asarray(y, float) for y in ys], cv_num, seed)
    folds = ref.cli._prepare_cv([x.astype(float) for x in Xs], [np.asarray(y, float) for y in ys], cv_num, seed)
    for j in range(cv_num):
        for a in range(len(Xs)):
            out[f"fold{j}_train_rows{a}"] = np.asarray(idx_folds[j].train_geno[a])[:, 0].astype(np.int64)
            out[f"fold{j}_valid_rows{a}"] = np.asarray(idx_folds[j].valid_geno[a])[:, 0].astype(np.int64)
            # copies: infer_sushie standardises its inputs in place (infer.py:326-333)
            out[f"fold{j}_train_pheno{a}"] = np.array(folds[j].train_pheno[a], dtype=float)
            out[f"fold{j}_valid_pheno{a}"] = np.array(folds[j].valid_pheno[a], dtype=float)
    fit, cs_out = fit_and_cs_ind([x.astype(float) for x in Xs], ys, None, L=bench.L_EFF, min_tol=1e-4)
    pack_fit("full", fit, cs_out, out, slim=True)
    print(f"config5 gene {gene} p={p}: full fit n_iter={fit.n_iter}", flush=True)
    # cli._run_cv with its infer_sushie calls recorded (cli.py:378-419)
    captured = []
    real = ref.infer.infer_sushie

    def recording(*a, **k):
        res = real(*a, **k)
        captured.append(res)
        return res

    args = types.SimpleNamespace(cv_num=cv_num, L=bench.L_EFF, no_scale=False, no_regress=False, no_update=False,
                                 resid_var=None, effect_var=None, rho=None, max_iter=500, min_tol=1e-4, threshold=0.95,
                                 purity=0.5, max_select=250, seed=seed)
    ref.infer.infer_sushie = recording
    try:
        cv_res = ref.cli._run_cv(args, folds, None)
    finally:
        ref.infer.infer_sushie = real
    out["cv_res"] = np.asarray(cv_res, dtype=float)  # [ancestry][adj_r2, p_value]
    for j, res in enumerate(captured):
        pack_fit(f"fold{j}", _RefFit(res), (res.cs, res.alphas, np.asarray(res.pip_all), np.asarray(res.pip_cs)), out, slim=True)
        print(f"config5 fold {j}: n_iter={len(res.elbo) - 1}", flush=True)
    print("config5 cv_res", out["cv_res"].tolist())
    np.savez_compressed(os.path.join(HERE, "config5_gene.npz"), **out)


def golden_config4(p=10000, iters=3):
    import bench_ss

    lds, ns, zs, causal = bench_ss.make_locus(p, np.float32)  # the fp32 LD the device stores, widened exactly
    fit, cs_out = fit_and_cs_ss(np.stack([ld.astype(np.float64) for ld in lds]), ns, np.stack(zs), L=10, max_iter=iters,
                                min_tol=0.0)
    out = {"p": p, "iters": iters, "causal": causal}
    pack_fit("c4", fit, cs_out, out, slim=True)
    print(f"config4 p={p}: n_iter={fit.n_iter} elbo={fit.elbo[1:]}", flush=True)
    np.savez_compressed(os.path.join(HERE, "config4_summary.npz"), **out)


if __name__ == "__main__":
    import warnings

    warnings.filterwarnings("ignore")
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["ind", "ss", "syn", "headline", "config5", "config4"]
    if "ind" in which:
        golden_example_ind()
    if "ss" in which:
        golden_example_ss()
    if "syn" in which:
        golden_synthetic()
    if "headline" in which:
        golden_headline()
    if "config5" in which:
        golden_config5()
    if "config4" in which:
        golden_config4()
