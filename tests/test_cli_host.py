"""CPU tests of the CLI host side (readers, QC, cross-validation split, writers); the fits
themselves are GPU tests (tests/test_gpu_cli.py)."""
import argparse
import os
import types

import numpy as np
import pandas as pd
import pytest

import clidata
from oracle import threefry
from sushie_b200 import _common, cli, cs as cs_mod, io


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    return clidata.make_dataset(str(tmp_path_factory.mktemp("clidata")))


def _parse(argv):
    argp = argparse.ArgumentParser()
    cli.build_finemap_parser(argp.add_subparsers())
    return argp.parse_args(["finemap"] + argv)


def _prepare(ds, fmt, extra=()):
    args = _parse(["--pheno", *ds["pheno"], f"--{fmt}", *ds[fmt], "--covar", *ds["covar"], *extra])
    n_pop, ai, keep, pi, gp, gf = cli.parameter_check(args)
    raw = io.read_data(n_pop, ai, args.pheno, args.covar, gp, gf)
    return args, cli.process_raw(raw, keep, pi, args.keep_ambiguous, args.maf, args.rint, args.no_regress,
                                 args.mega, args.cv, args.cv_num, args.seed, args.chrom, args.start, args.end)


def test_flag_surface_matches_reference_defaults():
    args = _parse([])
    expect = dict(L=10, pi="uniform", max_iter=500, min_tol=1e-3, threshold=0.95, purity=0.5, purity_method="weighted",
                  ld_adjust=0, max_select=250, min_snps=100, maf=0.01, cv_num=5, seed=12345, trait="Trait",
                  platform="cpu", jax_precision=64, output="sushie_finemap", gwas_sig=1.0, gwas_sig_type="at-least",
                  gwas_header=["chrom", "snp", "pos", "a1", "a0", "z"])
    for key, val in expect.items():
        assert getattr(args, key) == val, key
    for flag in ("summary", "no_scale", "no_regress", "no_update", "rint", "no_reorder", "keep_ambiguous", "meta",
                 "mega", "her", "cv", "alphas", "numpy", "quiet", "verbose", "compress"):
        assert getattr(args, flag) is False


def test_plink_and_vcf_readers_agree(dataset):
    for a in range(2):
        b1, f1, g1 = io.read_vcf(dataset["vcf"][a])
        b2, f2, g2 = io.read_triplet(dataset["plink"][a])
        assert list(b1.columns) == ["chrom", "snp", "pos", "a0", "a1"] == list(b2.columns)
        assert (b1.snp.values == b2.snp.values).all() and (b1.a0.values == b2.a0.values).all()
        assert (f1.iid.values == f2.iid.values).all()
        assert np.isnan(g1).sum() == 1 and np.array_equal(np.isnan(g1), np.isnan(g2))
        np.testing.assert_array_equal(np.nan_to_num(g1, nan=-1), np.nan_to_num(g2, nan=-1))


def test_process_raw_qc(dataset):
    _, (snps, reg, mega, cv) = _prepare(dataset, "vcf")
    assert list(snps.columns) == ["SNPIndex", "chrom", "snp", "pos", "a0", "a1", "pi"]
    assert snps.shape[0] == dataset["p"] - 3 and not (set(snps.snp) & dataset["dropped"])
    assert snps.snp.is_unique and not snps.snp.str.startswith("rsPRIV").any()
    assert [g.shape for g in reg.geno] == [(150, 127), (180, 127)] and mega is None and cv is None
    assert not any(np.isnan(g).any() for g in reg.geno)
    np.testing.assert_allclose(snps.pi.values, 1.0 / 127)
    # swapped alleles are flipped back: the two ancestries count the same allele again
    for s in dataset["swapped"]:
        j = int(snps.index[snps.snp == s][0])
        assert abs(reg.geno[0][:, j].mean() - reg.geno[1][:, j].mean()) < 0.35
    # phenotype rows follow the genotype (fam) order although the file is shuffled
    ph = pd.read_csv(dataset["pheno"][0], sep="\t", header=None).set_index(0)[1]
    np.testing.assert_allclose(reg.pheno[0], ph.loc[[f"anc1_{i}" for i in range(150)]].values)
    # the plink route is identical
    _, (snps2, reg2, _, _) = _prepare(dataset, "plink")
    assert snps2.equals(snps) and all(np.array_equal(a, b) for a, b in zip(reg.geno, reg2.geno))
    # keeping ambiguous SNPs brings one back; a region filter cuts by position
    _, (snps3, _, _, _) = _prepare(dataset, "vcf", ["--keep-ambiguous"])
    assert snps3.shape[0] == 128
    _, (snps4, _, _, _) = _prepare(dataset, "vcf", ["--chrom", "1", "--start", "10000", "--end", "11000"])
    assert snps4.pos.max() <= 11000 and snps4.shape[0] < 40


def test_cv_split_uses_the_jax_permutation_stream(dataset):
    args, (snps, reg, mega, cv) = _prepare(dataset, "vcf", ["--cv", "--mega", "--cv-num", "4", "--seed", "77"])
    assert len(cv) == 4 and mega.geno[0].shape == (330, 127) and mega.covar is None
    key = threefry.prng_key(77)
    for idx, n in enumerate((150, 180)):
        pair = threefry.split(key, 2)
        key, sub = pair[0], pair[1]
        order = threefry.choice_without_replacement(sub, np.arange(n), n)
        chunks = np.array_split(order, 4)
        from sushie_b200 import utils
        resid_g, _ = utils.regress_covar(reg.geno[idx], reg.pheno[idx], reg.covar[idx], False)
        np.testing.assert_allclose(cv[1].valid_geno[idx], resid_g[chunks[1]])
        assert cv[1].train_geno[idx].shape[0] == n - len(chunks[1])


def test_summary_preparation(dataset):
    args = _parse(["--summary", "--gwas", *dataset["gwas"], "--ld", *dataset["ld"], "--sample-size", "150", "180"])
    n_pop, pi, gp, gf, ld_file = cli.parameter_check_ss(args)
    snps, data = cli.process_raw_ss(gp, gf, ld_file, pi, args)
    assert ld_file and n_pop == 2 and snps.shape[0] == 128  # 129 regular SNPs minus the ambiguous one
    assert [ld.shape for ld in data.lds] == [(128, 128)] * 2 and data.ns.shape == (2, 1)
    # reference genotypes instead of LD files: LD is computed from the cleaned, allele-matched genotypes
    args = _parse(["--summary", "--gwas", *dataset["gwas"], "--vcf", *dataset["vcf"], "--sample-size", "150", "180"])
    n_pop, pi, gp, gf, ld_file = cli.parameter_check_ss(args)
    snps2, data2 = cli.process_raw_ss(gp, gf, ld_file, pi, args)
    assert not ld_file and snps2.shape[0] == 127
    np.testing.assert_allclose(np.diag(data2.lds[1]), 1.0, atol=1e-12)
    common = snps.merge(snps2, on="snp")
    common = common[common.snp != "rs1030"]  # the VCF has one missing call there (mean-imputed)
    i1, i2 = common.SNPIndex_x.values, common.SNPIndex_y.values
    np.testing.assert_allclose(data.lds[0][np.ix_(i1, i1)], data2.lds[0][np.ix_(i2, i2)], atol=1e-3)
    np.testing.assert_allclose(data.zs[1][i1], data2.zs[1][i2], atol=1e-9)


@pytest.mark.parametrize("argv,msg", [
    ([], "No phenotype file specified"),
    (["--pheno", "a", "b", "--vcf", "x"], "numbers of ancestries in vcf geno and pheno data does not match"),
    (["--pheno", "a", "--vcf", "x", "--seed", "0"], "seed specified for randomization must be greater than 0"),
    (["--pheno", "a", "--vcf", "x", "--maf", "0.7"], "minor allele frequency"),
    (["--pheno", "a", "--vcf", "x", "--cv", "--cv-num", "1"], "folds in cross validation must be greater than 1"),
    (["--pheno", "a", "--vcf", "x", "--chrom", "1"], "region is not specified correctly"),
    (["--pheno", "a"], "No genotype data specified"),
])
def test_parameter_check_errors(argv, msg):
    with pytest.raises(ValueError, match=msg):
        cli.parameter_check(_parse(argv))


@pytest.mark.parametrize("argv,msg", [
    (["--summary"], "No GWAS summary statistics file specified"),
    (["--summary", "--gwas", "g", "--ld", "l"], "No sample size specified"),
    (["--summary", "--gwas", "g", "--ld", "l", "--sample-size", "-3"], "sample size specified .* is invalid"),
    (["--summary", "--gwas", "g", "--ld", "l", "--sample-size", "5", "--gwas-sig", "2"], "significance threshold"),
    (["--summary", "--gwas", "g", "--ld", "l", "--sample-size", "5", "--ld-adjust", "0.5"], "LD adjustment parameter"),
])
def test_parameter_check_ss_errors(argv, msg):
    with pytest.raises(ValueError, match=msg):
        cli.parameter_check_ss(_parse(argv))


def _fake_result(p=30, L=3, k=2, seed=0):
    rng = np.random.default_rng(seed)
    alpha = rng.dirichlet(np.ones(p) * 0.05, size=L)
    fit = types.SimpleNamespace(
        alpha=alpha, post_mean=rng.standard_normal((L, p, k)) * alpha[:, :, None],
        post_mean_sq=np.zeros((L, p, k, k)), weighted_sum_covar=np.stack([np.eye(k) * (l + 1.0) + 0.2 for l in range(L)]),
        kl=np.zeros(L), log_bf=rng.standard_normal((L, p)), effect_covar=np.stack([np.eye(k)] * L),
        resid_var=np.ones(k), elbo=np.array([-np.inf, -1.0]), elbo_increase=True,
    )
    ns = np.array([[100], [120]])[:k]

    def builder(a, lbf):
        return cs_mod.build_tables(a, lbf, ns, lambda sets: np.full((len(sets), k), 0.9), 0.95, 0.5, "weighted", 250, 1)

    return _common.assemble_result(fit, None, ns, True, builder)


def test_output_tables_follow_the_reference_schema(tmp_path):
    res = _fake_result()
    p = res.posteriors.alpha.shape[1]
    snps = pd.DataFrame({"SNPIndex": np.arange(p), "chrom": 1, "snp": [f"rs{j}" for j in range(p)],
                         "pos": np.arange(p) + 5, "a0": "A", "a1": "C", "pi": 1.0 / p})
    out = str(tmp_path / "t.sushie")
    cs = io.output_cs([res], None, snps, out, "GENE", False, "sushie")
    assert list(cs.columns) == ["SNPIndex", "chrom", "snp", "pos", "a0", "a1", "pi", "CSIndex", "alpha", "c_alpha",
                                "pip_all", "pip_cs", "trait", "n_snps", "ancestry"]
    assert (cs.ancestry == "sushie").all() and (cs.n_snps == p).all()
    assert cs.groupby("CSIndex").alpha.apply(lambda s: s.is_monotonic_decreasing).all()
    w = io.output_weights([res], None, snps, out, "GENE", False, "sushie")
    assert list(w.columns)[-6:] == ["n_snps", "ancestry1_sushie_weight", "ancestry2_sushie_weight", "sushie_pip_all",
                                    "sushie_pip_cs", "sushie_cs_index"]
    np.testing.assert_allclose(w.ancestry1_sushie_weight.values, res.posteriors.post_mean.sum(axis=0)[:, 0])
    in_cs = set(cs.SNPIndex)
    assert all((v == "No CS") == (j not in in_cs) for j, v in zip(w.SNPIndex, w.sushie_cs_index))
    a = io.output_alphas([res], snps, out, "GENE", False, "sushie", 0.5)
    assert {"alpha_l1", "in_cs_l1", "purity_l1", "kept_l1", "log_bf_l1", "purity_threshold", "ancestry"} <= set(a.columns)
    c = io.output_corr([res], out, "GENE", False)
    assert list(c.columns) == ["trait", "CSIndex", "ancestry1_est_var", "ancestry1_ancestry2_est_covar",
                               "ancestry1_ancestry2_est_corr", "ancestry2_est_var"]
    np.testing.assert_allclose(c.ancestry1_ancestry2_est_corr.values, 0.2 / (np.arange(1, 4) + 0.2))
    cv = io.output_cv([[0.3, 1e-5], [0.2, 1e-3]], np.array([100, 120]), out, "GENE", True)
    assert list(cv.columns) == ["ancestry", "rsq", "p_value", "N", "trait"] and os.path.exists(out + ".cv.tsv.gz")
    io.output_numpy([res], snps, out)
    box = np.load(out + ".all.results.npy", allow_pickle=True)
    assert box[0].equals(snps) and len(box[1]) == 1
    for stem in ("cs", "weights", "alphas", "corr"):
        assert pd.read_csv(f"{out}.{stem}.tsv", sep="\t").shape[0] > 0
    # meta layout: one block per ancestry + combined PIPs
    one = _fake_result(k=1)
    meta = io.output_weights([one, one], [one.pip_all, one.pip_cs], snps, out, "GENE", False, "meta")
    assert {"ancestry1_single_weight", "ancestry2_single_pip_all", "ancestry2_cs_index", "meta_pip_all"} <= set(meta.columns)


def test_finemap_batch_worker_processes_plumbing(tmp_path):
    """Several `--devices`: one spawned worker process per GPU pulls windows of manifest lines and logs through
    the parent.  Every line here fails before any fit (missing files / bad arguments), so no GPU is needed:
    the test covers the spawn, the window queue, the log forwarding and the counts."""
    from sushie_b200 import cli

    manifest = tmp_path / "genes.txt"
    with open(manifest, "w") as fh:
        fh.write("# comment\n\n")
        for g in range(5):
            fh.write(f"--pheno /nonexistent/p{g} --vcf /nonexistent/v{g}.vcf --trait GENE{g} --output {tmp_path}/o{g}\n")
        fh.write("--no-such-flag 1\n")
    out = str(tmp_path / "batch")
    assert cli._main(["finemap-batch", "--manifest", str(manifest), "--devices", "0", "1", "--quiet", "--output", out]) == 0
    text = open(out + ".log").read()
    assert "0 genes written, 6 failed" in text
    for g in range(5):
        assert f"GENE{g}" in text
    assert "arguments could not be parsed" in text


def test_vcf_genotype_classes_follow_cyvcf2_gts012(tmp_path):
    """``read_vcf`` codes ``2 - gt_types`` with cyvcf2's ``gts012`` classes (reference ``io.py:243-255``): a
    half-missing call and a two-ALT heterozygote are HET, ``2/2`` is HOM_ALT, haploid calls are 0 or 2 copies."""
    calls = ["0/0", "0|1", "1/1", "./.", "0/.", "./1", "1/2", "2/2", "0", "1", ".", "2", "1|0:35"]
    path = tmp_path / "t.vcf"
    with open(path, "w") as fh:
        fh.write("##fileformat=VCFv4.2\n")
        fh.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"s{i}" for i in range(len(calls))) + "\n")
        fh.write("1\t100\trs1\tA\tC,G\t.\tPASS\t.\tGT:DP\t" + "\t".join(c if ":" in c else c + ":9" for c in calls) + "\n")
    bim, fam, bed = io.read_vcf(str(path))
    want = [2, 1, 0, np.nan, 1, 1, 1, 0, 2, 0, np.nan, np.nan, 1]
    np.testing.assert_array_equal(bed[:, 0], np.asarray(want, dtype=float))
    assert bim.a0[0] == "C" and bim.a1[0] == "A" and list(fam.iid) == [f"s{i}" for i in range(len(calls))]
