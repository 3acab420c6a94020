"""End-to-end ``sushie finemap`` runs on the GPU: output files vs the numpy oracle fed the same cleaned data."""
import argparse
import os

import numpy as np
import pandas as pd
import pytest

import clidata
from parity import PIP_TOL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    return clidata.make_dataset(str(tmp_path_factory.mktemp("clidata")))


def _tsv(prefix, stem):
    return pd.read_csv(f"{prefix}.{stem}.tsv", sep="\t")


def _cleaned(ds, fmt="vcf", extra=()):
    from sushie_b200 import cli, io

    argp = argparse.ArgumentParser()
    cli.build_finemap_parser(argp.add_subparsers())
    args = argp.parse_args(["finemap", "--pheno", *ds["pheno"], f"--{fmt}", *ds[fmt], "--covar", *ds["covar"], *extra])
    n_pop, ai, keep, pi, gp, gf = cli.parameter_check(args)
    raw = io.read_data(n_pop, ai, args.pheno, args.covar, gp, gf)
    return cli.process_raw(raw, keep, pi, False, 0.01, False, False, False, False, 5, 12345, None, None, None)


def test_finemap_individual_level_outputs_match_oracle(engine_lib, dataset, tmp_path):
    from oracle import ibss
    from sushie_b200 import cli

    out = str(tmp_path / "run")
    rc = cli._main(["finemap", "--pheno", *dataset["pheno"], "--vcf", *dataset["vcf"], "--covar", *dataset["covar"],
                    "--L", "5", "--trait", "GENE1", "--alphas", "--numpy", "--cv", "--cv-num", "3", "--meta", "--mega",
                    "--quiet", "--output", out])
    assert rc == 0
    log_text = open(out + ".log").read()
    assert "ERROR" not in log_text and "Fine-mapping finishes for GENE1" in log_text
    for stem in ("sushie.cs", "sushie.weights", "sushie.alphas", "sushie.corr", "sushie.cv", "meta.cs", "meta.weights",
                 "mega.cs", "mega.weights"):
        assert os.path.exists(f"{out}.{stem}.tsv"), stem
    assert os.path.exists(out + ".sushie.all.results.npy")

    snps, reg, _, _ = _cleaned(dataset)
    ref = ibss.fit_individual(reg.geno, reg.pheno, covar=reg.covar, L=5, min_tol=1e-3)
    w = _tsv(out, "sushie.weights")
    assert (w.snp.values == snps.snp.values).all() and (w.trait == "GENE1").all()
    assert np.max(np.abs(w.sushie_pip_all.values - ibss.make_pip(ref.posteriors.alpha))) <= PIP_TOL
    np.testing.assert_allclose(
        w[["ancestry1_sushie_weight", "ancestry2_sushie_weight"]].values, ref.posteriors.post_mean.sum(axis=0), atol=1e-6)
    top = set(w.sort_values("sushie_pip_all", ascending=False).snp.values[:2])
    assert top == set(dataset["causal"])
    cs = _tsv(out, "sushie.cs")
    assert set(dataset["causal"]) <= set(cs.snp) and set(cs.columns) >= {"CSIndex", "alpha", "c_alpha", "pip_all", "pip_cs"}
    corr = _tsv(out, "sushie.corr")
    assert corr.shape[0] == 5 and np.isfinite(corr.ancestry1_ancestry2_est_corr.values[:2]).all()
    cv = _tsv(out, "sushie.cv")
    assert list(cv.ancestry) == [1, 2] and list(cv.N) == [150, 180] and (cv.rsq > 0.2).all() and (cv.p_value < 1e-6).all()
    meta = _tsv(out, "meta.weights")
    assert {"ancestry1_single_pip_all", "ancestry2_single_pip_all", "meta_pip_all"} <= set(meta.columns)
    combined = 1 - (1 - meta.ancestry1_single_pip_all) * (1 - meta.ancestry2_single_pip_all)
    np.testing.assert_allclose(meta.meta_pip_all.values, combined.values, atol=1e-10)
    single = ibss.fit_individual([reg.geno[1]], [reg.pheno[1]], covar=[reg.covar[1]], L=5, min_tol=1e-3)
    assert np.max(np.abs(meta.ancestry2_single_pip_all.values - ibss.make_pip(single.posteriors.alpha))) <= PIP_TOL
    mega = _tsv(out, "mega.weights")
    assert "mega_weight" in mega.columns and mega.shape[0] == snps.shape[0]


def test_finemap_plink_without_covariates_uses_hard_call_storage(engine_lib, dataset, tmp_path):
    """No covariates -> genotypes stay 0/1/2 -> one-byte device storage; same answer as the oracle."""
    from oracle import ibss
    from sushie_b200 import cli, io

    out = str(tmp_path / "plink")
    cli._main(["finemap", "--pheno", *dataset["pheno"], "--plink", *dataset["plink"], "--L", "4", "--min-tol", "1e-4",
               "--quiet", "--output", out])
    argp = argparse.ArgumentParser()
    cli.build_finemap_parser(argp.add_subparsers())
    args = argp.parse_args(["finemap", "--pheno", *dataset["pheno"], "--plink", *dataset["plink"]])
    n_pop, ai, keep, pi, gp, gf = cli.parameter_check(args)
    raw = io.read_data(n_pop, ai, args.pheno, None, gp, gf)
    snps, reg, _, _ = cli.process_raw(raw, keep, pi, False, 0.01, False, False, False, False, 5, 12345, None, None, None)
    # the single mean-imputed entry makes ancestry 1 non-integer, so use the same cleaned matrices for the oracle
    ref = ibss.fit_individual(reg.geno, reg.pheno, L=4, min_tol=1e-4)
    w = _tsv(out, "sushie.weights")
    assert np.max(np.abs(w.sushie_pip_all.values - ibss.make_pip(ref.posteriors.alpha))) <= PIP_TOL


def test_finemap_summary_level_outputs_match_oracle(engine_lib, dataset, tmp_path):
    from oracle import ibss
    from sushie_b200 import cli

    out = str(tmp_path / "ss")
    cli._main(["finemap", "--summary", "--gwas", *dataset["gwas"], "--ld", *dataset["ld"], "--sample-size", "150", "180",
               "--L", "5", "--meta", "--alphas", "--quiet", "--output", out])
    assert "ERROR" not in open(out + ".log").read()
    argp = argparse.ArgumentParser()
    cli.build_finemap_parser(argp.add_subparsers())
    args = argp.parse_args(["finemap", "--summary", "--gwas", *dataset["gwas"], "--ld", *dataset["ld"],
                            "--sample-size", "150", "180"])
    n_pop, pi, gp, gf, ld_file = cli.parameter_check_ss(args)
    snps, data = cli.process_raw_ss(gp, gf, ld_file, pi, args)
    ref = ibss.fit_summary(data.lds, data.ns, data.zs, L=5, min_tol=1e-3)
    w = _tsv(out, "sushie.weights")
    assert (w.snp.values == snps.snp.values).all()
    assert np.max(np.abs(w.sushie_pip_all.values - ibss.make_pip(ref.posteriors.alpha))) <= PIP_TOL
    assert set(w.sort_values("sushie_pip_all", ascending=False).snp.values[:2]) == set(dataset["causal"])
    assert os.path.exists(out + ".meta.weights.tsv") and os.path.exists(out + ".sushie.alphas.tsv")
    assert not os.path.exists(out + ".meta.corr.tsv")


def test_finemap_batch_manifest_equals_separate_runs(engine_lib, dataset, tmp_path):
    """`finemap-batch`: one process, one device work queue for all genes and all their fits; every gene gets the
    files a separate `finemap` run writes; a broken line is reported and skipped."""
    from sushie_b200 import _capi, cli

    ind = ["--pheno", *dataset["pheno"], "--vcf", *dataset["vcf"], "--covar", *dataset["covar"], "--L", "4"]
    ss = ["--summary", "--gwas", *dataset["gwas"], "--ld", *dataset["ld"], "--sample-size", "150", "180", "--L", "4"]
    genes = {
        "g1": ind + ["--cv", "--cv-num", "3", "--meta", "--trait", "G1"],
        "g2": ["--pheno", *dataset["pheno"], "--plink", *dataset["plink"], "--L", "3", "--mega", "--trait", "G2"],
        "g3": ss + ["--meta", "--trait", "G3"],
    }
    manifest = tmp_path / "genes.txt"
    with open(manifest, "w") as fh:
        fh.write("# one gene per line\n")
        for name, argv in genes.items():
            fh.write(" ".join(argv + ["--output", str(tmp_path / f"batch_{name}")]) + "\n")
        fh.write("--pheno /nonexistent/file --vcf /nonexistent.vcf --trait BROKEN --output " + str(tmp_path / "bad") + "\n")
    devices = [str(d) for d in range(_capi.device_count())]
    cli._main(["finemap-batch", "--manifest", str(manifest), "--devices", *devices, "--quiet",
               "--output", str(tmp_path / "batch")])
    log_text = open(str(tmp_path / "batch") + ".log").read()
    assert "3 genes written, 1 failed" in log_text and "BROKEN" in log_text
    for name, argv in genes.items():
        cli._main(["finemap", *argv, "--quiet", "--output", str(tmp_path / f"single_{name}")])
    stems = {"g1": ["sushie.cs", "sushie.weights", "sushie.corr", "sushie.cv", "meta.cs", "meta.weights"],
             "g2": ["sushie.cs", "sushie.weights", "sushie.corr", "mega.cs", "mega.weights"],
             "g3": ["sushie.cs", "sushie.weights", "sushie.corr", "meta.weights"]}
    for name, files in stems.items():
        for stem in files:
            a = _tsv(str(tmp_path / f"batch_{name}"), stem)
            b = _tsv(str(tmp_path / f"single_{name}"), stem)
            assert list(a.columns) == list(b.columns) and a.shape == b.shape, (name, stem)
            num = a.select_dtypes("number").columns
            np.testing.assert_allclose(a[num].values, b[num].values, rtol=1e-6, atol=1e-9, err_msg=f"{name} {stem}")
