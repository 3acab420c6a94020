"""``sushie finemap`` command line on top of the B200 IBSS engine (reference ``sushie/cli.py``).

The flag surface (``cli.py:2142-2793``), the QC order of ``process_raw`` / ``process_raw_ss``
(``cli.py:897-1735``), the regular / ``--meta`` / ``--mega`` / ``--cv`` wrappers
(``cli.py:323-419,1738-2039``) and the output files (``docs/files.rst``) follow the reference;
inference goes through :func:`sushie_b200.infer_sushie` / :func:`sushie_b200.infer_sushie_ss`
(and their batch forms for the cross-validation folds and per-ancestry fits), i.e. through the
CUDA engine.  ``--platform`` / ``--jax-precision`` are accepted for compatibility: the platform is
always the GPU, the precision selects the device storage of dense matrices (arithmetic is fp64).
"""
from __future__ import annotations

import argparse
import copy
import logging
import os
import sys
from typing import Callable, List, Optional, Tuple

import numpy as np
import pandas as pd
from scipy.stats import norm

from . import config, cs as _cs, infer, infer_ss, io, log, utils

__all__ = [
    "parameter_check", "parameter_check_ss", "process_raw", "process_raw_ss", "sushie_wrapper",
    "sushie_wrapper_ss", "run_finemap", "run_finemap_batch", "build_finemap_parser", "build_batch_parser", "run_cli",
]

_AMBIGUOUS = ("AT", "TA", "CG", "GC")


# ------------------------------------------------------------------------------------------
# per-ancestry cleaning steps (cli.py:33-320)


def _keep_file_subjects(raw: io.RawData, keep_subject: List[str], idx: int) -> io.RawData:
    log.logger.debug(f"Remove individuals based on the keep file for ancestry {idx + 1}.")
    n0 = raw.pheno.shape[0]
    pheno = raw.pheno[raw.pheno.iid.isin(keep_subject)].reset_index(drop=True)
    if n0 != pheno.shape[0]:
        log.logger.debug(
            f"Ancestry {idx + 1}:Drop {n0 - pheno.shape[0]} out of {n0} subjects because of the keep subject file."
        )
    if pheno.shape[0] == 0:
        raise ValueError(
            f"Ancestry {idx + 1}: All subjects are removed because of the keep subject file. Check the source."
        )
    return raw._replace(pheno=pheno)


def _drop_na_subjects(raw: io.RawData, idx: int) -> io.RawData:
    log.logger.debug(f"Remove individuals based on NA value for ancestry {idx + 1}.")
    n0 = raw.pheno.shape[0]
    pheno = raw.pheno.dropna().reset_index(drop=True)
    if n0 != pheno.shape[0]:
        log.logger.debug(
            f"Ancestry {idx + 1}: Drop {n0 - pheno.shape[0]} out of {n0} subjects"
            + " because of INF or NAN value in phenotype data."
        )
    if pheno.shape[0] == 0:
        raise ValueError(f"Ancestry {idx + 1}: All subjects have INF or NAN value in phenotype data. Check the source.")
    covar = raw.covar
    if covar is not None:
        c0 = covar.shape[0]
        covar = covar.dropna().reset_index(drop=True)
        if c0 != covar.shape[0]:
            log.logger.debug(
                f"Ancestry {idx + 1}: Drop {c0 - covar.shape[0]} out of {c0} subjects"
                + " because of INF or NAN value in covariates data."
            )
        if covar.shape[0] == 0:
            raise ValueError(
                f"Ancestry {idx + 1}: All subjects have INF or NAN value in covariates data. Check the source."
            )
    return raw._replace(pheno=pheno, covar=covar)


def _impute_geno(raw: io.RawData, idx: int) -> io.RawData:
    """Drop SNPs missing in everybody, fill the other gaps with the column mean."""
    log.logger.debug(f"Impute genotypes for ancestry {idx + 1}.")
    bim = raw.bim.reset_index(drop=True)
    bed = np.asarray(raw.bed, dtype=np.float64)
    n0 = bim.shape[0]
    gap = np.isnan(bed)
    all_gap = np.nonzero(gap.all(axis=0))[0]
    if len(all_gap) == n0:
        raise ValueError(f"Ancestry {idx + 1}: All SNPs have INF or NAN value in genotype data. Check the source.")
    if len(all_gap):
        bim = bim.drop(all_gap).reset_index(drop=True)
        bed = np.delete(bed, all_gap, axis=1)
        gap = np.delete(gap, all_gap, axis=1)
        log.logger.debug(
            f"Ancestry {idx + 1}: Drop {len(all_gap)} out of {n0} SNPs because all subjects have NAN value"
            + " in genotype data."
        )
    if gap.any():
        rows, cols = np.nonzero(gap)
        bed = bed.copy()
        bed[rows, cols] = np.nanmean(bed, axis=0)[cols]
        log.logger.debug(
            f"Ancestry {idx + 1}: Impute {len(cols)} out of {bim.shape[0]} SNPs with NAN value based on allele"
            + " frequency."
        )
    return raw._replace(bim=bim, bed=bed)


def _filter_maf(raw: io.RawData, maf: float, idx: int) -> io.RawData:
    log.logger.debug(f"Filter genotypes based on MAF for ancestry {idx + 1}.")
    n0 = raw.bim.shape[0]
    freq = np.mean(raw.bed, axis=0) / 2
    freq = np.where(freq > 0.5, 1 - freq, freq)
    keep = np.nonzero(freq >= maf)[0]
    if len(keep) == 0:
        raise ValueError(f"Ancestry {idx + 1}: All SNPs cannot pass the MAF threshold at {maf}.")
    if len(keep) != n0:
        log.logger.debug(f"Ancestry {idx + 1}: Drop {n0 - len(keep)} out of {n0} SNPs because of maf threshold at {maf}.")
    return raw._replace(bim=raw.bim.iloc[keep, :].reset_index(drop=True), bed=raw.bed[:, keep])


def _remove_dup_geno(raw: io.RawData, idx: int) -> io.RawData:
    log.logger.debug(f"Remove duplicated individuals based on genotype data for ancestry {idx + 1}.")
    bim = raw.bim.reset_index(drop=True)
    dup = np.nonzero(bim.snp.duplicated().values)[0]
    if len(dup):
        log.logger.debug(
            f"Ancestry {idx + 1}: Drop {len(dup)} out of {bim.shape[0]} SNPs because of duplicates in the rsID"
            + " in genotype data."
        )
    return raw._replace(bim=bim.drop(dup).reset_index(drop=True), bed=np.delete(raw.bed, dup, axis=1))


def _indexed(df: pd.DataFrame, name: str) -> pd.DataFrame:
    return df.reset_index(drop=True).reset_index().rename(columns={"index": name})


def _reset_idx(raw: io.RawData, idx: int) -> io.RawData:
    tag = idx + 1
    bim = _indexed(raw.bim, f"bimIDX_{tag}").rename(
        columns={"pos": f"pos_{tag}", "a0": f"a0_{tag}", "a1": f"a1_{tag}"}
    )
    covar = None if raw.covar is None else _indexed(raw.covar, f"covarIDX_{tag}")
    return raw._replace(
        bim=bim, fam=_indexed(raw.fam, f"famIDX_{tag}"), pheno=_indexed(raw.pheno, f"phenoIDX_{tag}"), covar=covar
    )


def _filter_common_ind(raw: io.RawData, idx: int) -> io.RawData:
    tag = idx + 1
    log.logger.debug(f"Keep common individuals based on genotype and phenotype data for ancestry {tag}.")
    common = raw.fam.merge(raw.pheno[[f"phenoIDX_{tag}", "iid"]], how="inner", on=["iid"])
    if raw.covar is not None:
        common = common.merge(raw.covar[[f"covarIDX_{tag}", "iid"]], how="inner", on=["iid"])
    if common.shape[0] == 0:
        raise ValueError(
            f"Ancestry {tag}: No common individuals across phenotype, covariates, genotype found. Check the source."
        )
    log.logger.debug(
        f"Ancestry {tag}: Found {common.shape[0]} common individuals across phenotype, covariates, and genotype."
    )
    return raw._replace(fam=common)


def _allele_check(base_a0, base_a1, cmp_a0, cmp_a1) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Positions whose allele pair is identical, swapped, or neither."""
    base_a0, base_a1 = np.asarray(base_a0, dtype=object), np.asarray(base_a1, dtype=object)
    cmp_a0, cmp_a1 = np.asarray(cmp_a0, dtype=object), np.asarray(cmp_a1, dtype=object)
    same = (base_a0 == cmp_a0) & (base_a1 == cmp_a1)
    swapped = (base_a0 == cmp_a1) & (base_a1 == cmp_a0)
    # a palindromic identical pair (a0 == a1) counts for both masks in the reference's 0/1 product arithmetic
    return np.nonzero(same)[0], np.nonzero(swapped)[0], np.nonzero(~(same | swapped))[0]


def _merge_across(tables: List[pd.DataFrame]) -> pd.DataFrame:
    out = tables[0]
    for nxt in tables[1:]:
        out = out.merge(nxt, how="inner", on=["chrom", "snp"]).reset_index(drop=True)
    return out


def _attach_pi(snps: pd.DataFrame, pi: pd.DataFrame) -> Tuple[pd.DataFrame, Optional[np.ndarray]]:
    if pi.shape[0] != 0:
        log.logger.debug("Process prior weight file.")
        snps = pd.merge(snps, pi, how="left", on="snp")
        missing = snps["pi"].isna().sum()
        if missing > 0:
            log.logger.debug(
                f"{missing} SNP(s) have missing prior weights. Will replace them with the mean value of the rest."
            )
        snps["pi"] = snps["pi"].fillna(snps["pi"].mean())
        return snps, np.asarray(snps["pi"].values, dtype=float)
    snps["pi"] = np.ones(snps.shape[0]) / float(snps.shape[0])
    return snps, None


# ------------------------------------------------------------------------------------------
# cross-validation (cli.py:323-419)


def _prepare_cv(geno: List[np.ndarray], pheno: List[np.ndarray], cv_num: int, seed: int) -> List[io.CVData]:
    """Row-shuffle every ancestry with JAX's threefry permutation stream (``random.split`` +
    ``random.choice(key, n, (n,), replace=False)``), cut into ``cv_num`` chunks, rank-normalise
    training and validation phenotypes separately."""
    key = _cs.prng_key(seed)
    n_pop = len(geno)
    bases, r_parts, y_parts = [], [], []
    for idx in range(n_pop):
        n = geno[idx].shape[0]
        key, sub = _cs.split_key(key)
        order = _cs.shuffle_with_key(sub, np.arange(n))
        bases.append(np.asarray(geno[idx]))
        r_parts.append(np.array_split(order, cv_num))  # the folds as row lists of the un-shuffled matrix
        y_parts.append(np.array_split(np.asarray(pheno[idx])[order], cv_num))
    folds = []
    for cv in range(cv_num):
        rest = [j for j in range(cv_num) if j != cv]
        folds.append(io.CVData(
            # training genotypes stay row VIEWS of the gene's matrix: the engine uploads that matrix once and gathers
            # every fold's rows on the device (np.asarray(view) is the matrix the reference builds)
            train_geno=[utils.RowView(bases[i], np.concatenate([r_parts[i][j] for j in rest])) for i in range(n_pop)],
            train_pheno=[utils.rint(np.concatenate([y_parts[i][j] for j in rest])) for i in range(n_pop)],
            valid_geno=[bases[i][r_parts[i][cv]] for i in range(n_pop)],
            valid_pheno=[utils.rint(y_parts[i][cv]) for i in range(n_pop)],
        ))
    return folds


def _cv_specs(args, cv_data: List[io.CVData], pi) -> List[dict]:
    """One ``infer_sushie`` keyword set per fold (``cli.py:378-402``: no covariates, default
    ``purity_method`` / ``min_snps`` / ``no_reorder``)."""
    kw = dict(covar=None, L=args.L, no_scale=args.no_scale, no_regress=args.no_regress, no_update=args.no_update, pi=pi,
              resid_var=args.resid_var, effect_var=args.effect_var, rho=args.rho, max_iter=args.max_iter,
              min_tol=args.min_tol, threshold=args.threshold, purity=args.purity, max_select=args.max_select,
              seed=args.seed)
    # (only the folds' posterior means are read, cli.py:404-410: their output tables are not assembled)
    return [dict(kw, Xs=fold.train_geno, ys=fold.train_pheno, lazy_tables=True) for fold in cv_data[: args.cv_num]]


def _cv_scores(cv_data: List[io.CVData], fits) -> List[List[float]]:
    """Out-of-fold predictions with the un-standardised validation genotypes, then adjusted r-squared and
    p-value per ancestry, as the reference does (``cli.py:404-419``)."""
    n_pop = len(cv_data[0].train_geno)
    est = [[] for _ in range(n_pop)]
    obs = [[] for _ in range(n_pop)]
    for fold, fit in zip(cv_data, fits):
        weight = np.sum(fit.posteriors.post_mean, axis=0)
        for idx in range(n_pop):
            est[idx].append(np.asarray(fold.valid_geno[idx], dtype=float) @ weight[:, idx])
            obs[idx].append(fold.valid_pheno[idx])
    out = []
    for idx in range(n_pop):
        _, adj_r2, p_value = utils.ols(np.concatenate(est[idx])[:, np.newaxis], np.concatenate(obs[idx])[:, np.newaxis])
        out.append([adj_r2[0], p_value[1][0]])
    return out


def _run_cv(args, cv_data: List[io.CVData], pi) -> List[List[float]]:
    """All folds are fitted as one device-resident batch (``cli.py:378-419``)."""
    return _cv_scores(cv_data, infer.infer_sushie_batch(_cv_specs(args, cv_data, pi)))


# ------------------------------------------------------------------------------------------
# argument checks (cli.py:422-894)


def _check_region(args) -> None:
    given = [args.chrom is not None, args.start is not None, args.end is not None]
    if not any(given):
        log.logger.debug("No region is specified. Will use all SNPs available in the data.")
        return
    if not all(given):
        raise ValueError(
            "The region is not specified correctly. Please provide --chrom, --start, and --end together,"
            + " or omit all of them."
        )
    if args.start <= 0:
        raise ValueError("The start position for the region must be greater than 0. Update with --start.")
    if args.end <= 0:
        raise ValueError("The end position for the region must be greater than 0. Update with --end.")
    if args.end <= args.start:
        raise ValueError("The end position for the region must be greater than --start. Update with --start or --end.")
    log.logger.info(
        f"Detect region (chrom{args.chrom}:{args.start}:{args.end}) to be fine-mapped."
        + " Will only use SNPs within the region."
    )


def _read_pi(args, dedup: bool) -> pd.DataFrame:
    if args.pi == "uniform":
        return pd.DataFrame()
    log.logger.info("Detect file that contains prior weights for each SNP to be causal.")
    pi = pd.read_csv(args.pi, header=None, sep="\t")
    if dedup:
        pi = pi.drop_duplicates(subset=pi.columns[0])
    if pi.shape[0] == 0:
        raise ValueError("No prior weights are listed in the prior file. Check the source.")
    if pi.shape[1] < 2:
        raise ValueError(
            "The prior file has less than 2 columns. It has to be at least two columns."
            + " The first column is the SNP ID, and the second column the prior probability."
        )
    if pi.shape[1] > 2:
        log.logger.debug("The prior file has more than 2 columns. Will only use the first two columns.")
    pi = pi.iloc[:, 0:2]
    pi.columns = ["snp", "pi"]
    return pi


def _pick_genotypes(args, n_pop: int, other: str, single_file: bool, allow_ld: bool):
    """First genotype source given, in the reference's order plink > vcf > bgen (> ld)."""
    sources = [("plink", io.read_triplet), ("vcf", io.read_vcf), ("bgen", io.read_bgen)]
    if allow_ld:
        sources.append(("ld", io.read_ld))
    given = [name for name, _ in sources if getattr(args, name) is not None]
    name_ancestry = "ancestry" if n_pop == 1 else "ancestries"
    if len(given) > 1:
        if allow_ld:
            log.logger.info(
                f"Detect {len(given)} genotype or LD files,"
                + " will only use one type of file in the order of 'plink, vcf, bgen, and ld'"
            )
        else:
            log.logger.info(
                f"Detect {len(given)} genotypes, will only use one type of genotypes in the order of 'plink, vcf, and bgen'"
            )
    for name, func in sources:
        paths = getattr(args, name)
        if paths is None:
            continue
        if single_file:
            if len(paths) > 1:
                raise ValueError(
                    f"Multiple {name} files are detected. Expectation is one when --ancestry-index is specified."
                )
        elif len(paths) != n_pop:
            if name == "ld":
                raise ValueError("The numbers of ancestries in ld geno and pheno data does not match. Check the source.")
            raise ValueError(f"The numbers of ancestries in {name} geno and {other} data does not match. Check the source.")
        if name == "ld":
            log.logger.info(f"Detect LD data in tsv files for {n_pop} {name_ancestry}.")
        else:
            log.logger.info(f"Detect genotype data in {name} format for {n_pop} {name_ancestry}.")
        return paths, func, name == "ld"
    if allow_ld:
        raise ValueError("No genotype/LD data specified in either plink, vcf, bgen, or LD files. Check the source.")
    raise ValueError("No genotype data specified in either plink, vcf, or bgen format. Check the source.")


def parameter_check(args) -> Tuple[int, pd.DataFrame, List[str], pd.DataFrame, List[str], Callable]:
    """Validate the individual-level options; returns (n_pop, ancestry_index, keep_subject, pi, geno paths, reader)."""
    if args.pheno is None:
        raise ValueError("No phenotype file specified. Specify --summary if summary-level fine-mapping is wanted.")
    if args.ancestry_index is not None:
        log.logger.debug("Read in ancestry index file.")
        ancestry_index = pd.read_csv(args.ancestry_index[0], header=None, sep="\t")
        n0 = ancestry_index.shape[0]
        ancestry_index = ancestry_index.drop_duplicates()
        if n0 != ancestry_index.shape[0]:
            log.logger.debug(f"Index file has {n0 - ancestry_index.shape[0]} duplicated subjects.")
        if ancestry_index[0].duplicated().sum() != 0:
            raise ValueError(
                "The ancestry index file contains subjects with multiple ancestry index. Check the source."
            )
        labels = np.sort(ancestry_index[1].unique())
        n_pop = len(labels)
        if not np.array_equal(labels, np.arange(n_pop) + 1):
            raise ValueError(
                "The ancestry index doesn't start from 1 continuously to the total number of ancestry."
                + f" Check {args.ancestry_index}."
            )
        if len(args.pheno) > 1:
            raise ValueError(
                "Multiple phenotype files are detected. Expectation is one when --ancestry-index is specified."
            )
        log.logger.debug(
            "Detect ancestry index file, so it expects to have single phenotype, genotype, and covariates files"
        )
    else:
        ancestry_index = pd.DataFrame()
        n_pop = len(args.pheno)
    name_ancestry = "ancestry" if n_pop == 1 else "ancestries"
    log.logger.info(f"Detect phenotypes for {args.trait} for {n_pop} {name_ancestry}.")

    geno_path, geno_func, _ = _pick_genotypes(args, n_pop, "pheno", args.ancestry_index is not None, allow_ld=False)

    if args.covar is not None:
        if args.ancestry_index is not None:
            if len(args.covar) > 1:
                raise ValueError(
                    "Multiple covariates files are detected. Expectation is one when --ancestry-index is specified."
                )
        elif len(args.covar) != n_pop:
            raise ValueError("The number of covariates data does not match geno data.")
        log.logger.info("Detect covariates data.")
    else:
        log.logger.info("No covariates detected for this analysis.")

    keep_subject: List[str] = []
    if args.keep is not None:
        log.logger.info("Detect keep subject file. The inference only performs on the subjects listed in the file.")
        df_keep = pd.read_csv(args.keep[0], header=None, sep="\t", dtype={0: object})[[0]]
        if df_keep.shape[0] == 0:
            raise ValueError("No subjects are listed in the keep subject file. Check the source.")
        n0 = df_keep.shape[0]
        df_keep = df_keep.drop_duplicates()
        if n0 != df_keep.shape[0]:
            log.logger.debug(f"The keep subject file has {n0 - df_keep.shape[0]} duplicated subjects.")
        keep_subject = df_keep[0].values.tolist()

    pi = _read_pi(args, dedup=False)
    if args.seed <= 0:
        raise ValueError(
            "The seed specified for randomization must be greater than 0. Choose a positive integer using --seed."
        )
    if args.cv and args.cv_num <= 1:
        raise ValueError("The number of folds in cross validation must be greater than 1. Update with --cv-num.")
    if args.maf <= 0 or args.maf > 0.5:
        raise ValueError(
            "The minor allele frequency (MAF) has to be between 0 (exclusive) and 0.5 (inclusive)."
            + " Choose a valid frequency using --maf."
        )
    if (args.meta or args.mega) and n_pop == 1:
        log.logger.debug(
            "The number of ancestry is 1, but --meta or --mega is specified. Will skip meta or mega SuSiE."
        )
    _check_region(args)
    log.logger.debug("Finish parameter check for individual-level fine-mapping.")
    return n_pop, ancestry_index, keep_subject, pi, geno_path, geno_func


def parameter_check_ss(args) -> Tuple[int, pd.DataFrame, List[str], Callable, bool]:
    """Validate the summary-level options; returns (n_pop, pi, geno / LD paths, reader, is LD file)."""
    if args.gwas is None:
        raise ValueError("No GWAS summary statistics file specified. Check the source.")
    n_pop = len(args.gwas)
    name_ancestry = "ancestry" if n_pop == 1 else "ancestries"
    log.logger.info(f"Detect GWAS files for {args.trait} for {n_pop} {name_ancestry}.")
    geno_path, geno_func, ld_file = _pick_genotypes(args, n_pop, "GWAS", False, allow_ld=True)

    if args.sample_size is None:
        raise ValueError("No sample size specified for summary-level fine-mapping. Check the source.")
    if len(args.sample_size) != n_pop:
        raise ValueError(
            "The numbers of ancestries in sample size and GWAS data does not match. Check the source"
            + " and update --sample-size."
        )
    if not all(n > 0 for n in args.sample_size):
        raise ValueError(
            "The sample size specified for summary-level fine-mapping is invalid. Choose a positive integer"
            + " using --sample-size."
        )
    log.logger.info(f"Detect sample sizes for {n_pop} {name_ancestry}.")

    pi = _read_pi(args, dedup=True)
    if args.seed <= 0:
        raise ValueError("The seed specified for randomization is invalid. Choose a positive integer using --seed.")
    if args.maf <= 0 or args.maf > 0.5:
        raise ValueError(
            "The minor allele frequency (MAF) has to be between 0 (exclusive) and 0.5 (inclusive)."
            + " Choose a valid frequency using --maf."
        )
    if args.meta and n_pop == 1:
        log.logger.debug("The number of ancestry is 1, but --meta is specified. Will skip meta or mega SuSiE.")
    _check_region(args)
    if args.gwas_sig <= 0 or args.gwas_sig > 1:
        raise ValueError(
            "The significance threshold for P-values in GWAS summary statistics must be greater than 0"
            + " and less than or equal to 1. Choose a valid number using --gwas-sig."
        )
    if args.ld_adjust > 0.1 or args.ld_adjust < 0:
        raise ValueError(
            "The LD adjustment parameter has to be greater than 0 and less than 0.1."
            + " Choose a valid number using --ld-adjust."
        )
    for flag, what in ((args.cv, "Cross-validation"), (args.mega, "Mega SuShiE"), (args.her, "Heritability estimation")):
        if flag:
            log.logger.debug(f"{what} is not supported for summary-level fine-mapping. The flag will be ignored.")
    log.logger.debug("Finish parameter check for summary-level fine-mapping.")
    return n_pop, pi, geno_path, geno_func, ld_file


# ------------------------------------------------------------------------------------------
# data preparation (cli.py:897-1735)


def _drop_mismatched(snps: pd.DataFrame, n_pop: int) -> pd.DataFrame:
    """Remove SNPs whose allele pair in a later ancestry is neither identical nor swapped."""
    for idx in range(1, n_pop):
        _, _, wrong = _allele_check(snps["a0_1"].values, snps["a1_1"].values,
                                    snps[f"a0_{idx + 1}"].values, snps[f"a1_{idx + 1}"].values)
        if len(wrong) != 0:
            snps = snps.drop(index=snps.index[wrong]).reset_index(drop=True)
            log.logger.debug(
                f"Ancestry{idx + 1} has {len(wrong)} alleles that"
                + "couldn't match to ancestry 1 and couldn't be flipped. Will remove these SNPs."
            )
        if snps.shape[0] == 0:
            raise ValueError(
                f"Ancestry {idx + 1} has none of correct or flippable SNPs matching to ancestry 1."
                + "Check the source."
            )
    return snps


def _drop_ambiguous(snps: pd.DataFrame, what: str) -> pd.DataFrame:
    log.logger.debug("Remove ambiguous SNPs.")
    ambiguous = (snps.a0_1 + snps.a1_1).isin(_AMBIGUOUS)
    snps = snps[~ambiguous].reset_index(drop=True)
    if snps.shape[0] == 0:
        raise ValueError(f"All SNPs are ambiguous in {what}. Check the source.")
    if ambiguous.sum() != 0:
        log.logger.debug(f"Drop {ambiguous.sum()} ambiguous SNPs in {what}.")
    return snps


def process_raw(
    rawData: List[io.RawData], keep_subject: List[str], pi: pd.DataFrame, keep_ambiguous: bool, maf: float,
    rint: bool, no_regress: bool, mega: bool, cv: bool, cv_num: int, seed: int, chrom, start, end,
) -> Tuple[pd.DataFrame, io.CleanData, Optional[io.CleanData], Optional[List[io.CVData]]]:
    """Subject / SNP QC, cross-ancestry SNP matching and allele flipping, phenotype transform, and the
    mega / cross-validation data sets (same order of operations as ``cli.py:897-1231``)."""
    n_pop = len(rawData)
    for idx in range(n_pop):
        raw = rawData[idx]
        if len(keep_subject) != 0:
            raw = _keep_file_subjects(raw, keep_subject, idx)
        raw = _drop_na_subjects(raw, idx)
        raw = _impute_geno(raw, idx)
        raw = _filter_maf(raw, maf, idx)
        raw = _remove_dup_geno(raw, idx)
        raw = _reset_idx(raw, idx)
        rawData[idx] = _filter_common_ind(raw, idx)

    log.logger.debug("Fine common SNPs across ancestries.")
    if n_pop > 1:
        snps = _merge_across([r.bim for r in rawData])
        if snps.shape[0] == 0:
            raise ValueError("Ancestries have no common SNPs. Check the source.")
        for idx in range(n_pop):
            log.logger.debug(
                f"Ancestry{idx + 1} has {rawData[idx].bim.shape[0] - snps.shape[0]} independent SNPs and"
                + f" {snps.shape[0]} common SNPs. Inference only performs on common SNPs."
            )
        log.logger.debug("Remove SNPs that do not have same alleles across ancestries.")
        snps = _drop_mismatched(snps, n_pop)
    else:
        snps = rawData[0].bim
    if not keep_ambiguous:
        snps = _drop_ambiguous(snps, "genotype data")

    log.logger.debug("Filter SNPs based on Chrom, Start, and End using coordinates of first ancestry.")
    snps["chrom"] = snps["chrom"].astype("int64")
    if chrom is not None:
        snps = io._region_filter(snps, "pos_1", chrom, start, end, "").reset_index(drop=True)

    flips: List[np.ndarray] = []
    if n_pop > 1:
        log.logger.debug("Flip the alleles of subsequent ancestries to match those of the first ancestry.")
        for idx in range(1, n_pop):
            _, flipped, _ = _allele_check(snps["a0_1"].values, snps["a1_1"].values,
                                          snps[f"a0_{idx + 1}"].values, snps[f"a1_{idx + 1}"].values)
            if len(flipped) != 0:
                log.logger.debug(
                    f"Ancestry{idx + 1} has {len(flipped)} flipped alleles from ancestry 1. Will flip these SNPs."
                )
            flips.append(flipped)
            snps = snps.drop(columns=[f"a0_{idx + 1}", f"a1_{idx + 1}", f"pos_{idx + 1}"])
    snps = snps.reset_index().rename(columns={"index": "SNPIndex", "a0_1": "a0", "a1_1": "a1", "pos_1": "pos"})
    snps, pi_arr = _attach_pi(snps, pi)

    geno, pheno, covar, total_ind = [], [], [], 0
    for idx in range(n_pop):
        tag = idx + 1
        log.logger.debug(f"Final process the rawdata for ancestry {tag}.")
        _, fam, bed, ph, cv_tab = rawData[idx]
        rows = fam[f"famIDX_{tag}"].values
        cols = snps[f"bimIDX_{tag}"].values
        snps = snps.drop(columns=[f"bimIDX_{tag}"])
        g = np.array(bed[rows, :][:, cols], dtype=np.float64)
        if idx > 0 and len(flips[idx - 1]) != 0:
            g[:, flips[idx - 1]] = 2 - g[:, flips[idx - 1]]
        y = np.asarray(ph["pheno"].values[fam[f"phenoIDX_{tag}"].values], dtype=float)
        total_ind += y.shape[0]
        if rint:
            y = utils.rint(y)
        geno.append(g)
        pheno.append(y)
        if cv_tab is not None:
            covar.append(np.asarray(cv_tab.iloc[fam[f"covarIDX_{tag}"].values, 2:cv_tab.shape[1]].values, dtype=float))
    data_covar = covar if covar else None
    regular = io.CleanData(geno=geno, pheno=pheno, covar=data_covar, pi=pi_arr)
    name_ancestry = "ancestry" if n_pop == 1 else "ancestries"
    log.logger.info(
        f"Prepare {geno[0].shape[1]} SNPs for {total_ind} individuals from {n_pop} {name_ancestry} after"
        + " data cleaning. Specify --verbose for details."
    )

    mega_data = None
    cv_data = None
    if mega or cv:  # both need covariates regressed out first
        r_geno, r_pheno = copy.deepcopy(geno), copy.deepcopy(pheno)
        if data_covar is not None:
            for idx in range(n_pop):
                r_geno[idx], r_pheno[idx] = utils.regress_covar(geno[idx], pheno[idx], data_covar[idx], no_regress)
        if cv:
            cv_data = _prepare_cv([np.asarray(g) for g in r_geno], [np.asarray(y) for y in r_pheno], cv_num, seed)
        if mega:
            mega_data = io.CleanData(
                geno=[np.concatenate(r_geno, axis=0)], pheno=[utils.rint(np.concatenate(r_pheno, axis=0))],
                covar=None, pi=pi_arr,
            )
    log.logger.debug("Finish preparing data for cross-validation and mega fine-mapping.")
    return snps, regular, mega_data, cv_data


def process_raw_ss(geno_path: List[str], geno_func: Callable, ld_file: bool, pi: pd.DataFrame, args):
    """GWAS z-scores + LD (from LD files or from reference genotypes): SNP matching, significance and
    ambiguity filters, allele flipping (same order of operations as ``cli.py:1234-1735``)."""
    n_pop = len(geno_path)
    refs: list = []   # LD data frames or cleaned RawData
    gwas: List[pd.DataFrame] = []
    for idx in range(n_pop):
        tag = idx + 1
        df = io.read_gwas(args.gwas[idx], args.gwas_header, args.chrom, args.start, args.end)
        df = df.rename(columns={"pos": f"pos_{tag}", "a0": f"a0_{tag}", "a1": f"a1_{tag}", "z": f"z_{tag}"})
        if ld_file:
            ld = geno_func(geno_path[idx])
            if ld.shape[0] == 0 or ld.shape[1] == 0:
                raise ValueError(f"Ancestry {tag}: No LD data found after removing INF and NAN values. Check the source.")
            if ld.shape[0] != ld.shape[1]:
                raise ValueError(f"Ancestry {tag}: The LD matrix is not square. Check the source.")
            ld = ld.round(4)  # the reference keeps 4 digits to avoid floating point issues
            if not np.all(np.linalg.eigvals(ld.values) >= -1e-8):
                raise ValueError(f"Ancestry {tag}: The LD matrix is not positive semi-definite. Check the source.")
            if not np.allclose(np.diag(ld.values), np.ones(ld.shape[0]), atol=1e-4):
                raise ValueError(f"Ancestry {tag}: The LD matrix diagonal is not 1. Check the source.")
            if df["snp"].isin(ld.columns).sum() == 0:
                raise ValueError(f"Ancestry {tag}: No common SNPs between GWAS and LD data. Check the source.")
            vals = ld.values.copy()
            vals[np.diag_indices_from(vals)] += args.ld_adjust
            refs.append(pd.DataFrame(vals, index=ld.index, columns=ld.columns))
        else:
            bim, fam, bed = geno_func(geno_path[idx])
            bim = bim.copy()
            bim.chrom = bim.chrom.astype(int)
            if bed.shape[0] == 0:
                raise ValueError(f"Ancestry {tag}: No genotype data found. Check the source.")
            raw = io.RawData(bim=bim, fam=fam, bed=bed, pheno=pd.DataFrame([]), covar=None)
            raw = _remove_dup_geno(_filter_maf(_impute_geno(raw, idx), args.maf, idx), idx)
            tab = _indexed(raw.bim, f"bimIDX_{tag}").rename(
                columns={"pos": f"pos_{tag}", "a0": f"a0_{tag}", "a1": f"a1_{tag}"}
            )
            if df["snp"].isin(tab["snp"]).sum() == 0:
                raise ValueError(f"Ancestry {tag}: No common SNPs between GWAS and genotype data. Check the source.")
            refs.append(raw._replace(bim=tab))
        gwas.append(df)

    log.logger.debug("Find common SNPs across ancestries.")
    snps_gwas = _merge_across(gwas) if n_pop > 1 else gwas[0].reset_index(drop=True)
    if snps_gwas.shape[0] == 0:
        raise ValueError("GWAS data have no common SNPs across ancestries. Check the source.")
    snps_ld = snps_bim = None
    if ld_file:
        snps_ld = pd.DataFrame({"snps": refs[0].columns})
        for other in refs[1:]:
            snps_ld = snps_ld[snps_ld.snps.isin(other.columns)]
        if snps_ld.shape[0] == 0:
            raise ValueError("LD data have no common SNPs across ancestries. Check the source.")
    else:
        snps_bim = _merge_across([r.bim for r in refs]) if n_pop > 1 else refs[0].bim.reset_index(drop=True)
        if snps_bim.shape[0] == 0:
            raise ValueError("Genotype data have no common SNPs across ancestries. Check the source.")

    z_threshold = norm.ppf(1 - args.gwas_sig / 2)
    z_abs = snps_gwas.filter(regex="^z_").abs().gt(z_threshold)
    selected = z_abs.any(axis=1) if args.gwas_sig_type == "at-least" else z_abs.all(axis=1)
    n0 = snps_gwas.shape[0]
    snps_gwas = snps_gwas[selected].copy().reset_index(drop=True)
    log.logger.debug(
        f"Keep {snps_gwas.shape[0]} out of {n0} SNPs at the significance threshold {args.gwas_sig}"
        + f" based on {args.gwas_sig_type} method."
    )
    if snps_gwas.shape[0] == 0:
        raise ValueError("No SNPs pass the GWAS significance threshold. Check the source or update --gwas-sig.")

    if n_pop > 1:
        snps_gwas = _drop_mismatched(snps_gwas, n_pop)
        if not ld_file:
            snps_bim = _drop_mismatched(snps_bim, n_pop)
    if not args.keep_ambiguous:
        snps_gwas = _drop_ambiguous(snps_gwas, "GWAS data")

    if n_pop > 1:
        log.logger.debug("Flip the alleles of subsequent ancestries to match those of the first ancestry.")
        for idx in range(1, n_pop):
            tag = idx + 1
            _, flipped, _ = _allele_check(snps_gwas["a0_1"].values, snps_gwas["a1_1"].values,
                                          snps_gwas[f"a0_{tag}"].values, snps_gwas[f"a1_{tag}"].values)
            if len(flipped) != 0:
                snps_gwas.loc[flipped, f"z_{tag}"] *= -1
            snps_gwas = snps_gwas.drop(columns=[f"a0_{tag}", f"a1_{tag}", f"pos_{tag}"])
            if not ld_file:
                _, flipped, _ = _allele_check(snps_bim["a0_1"].values, snps_bim["a1_1"].values,
                                              snps_bim[f"a0_{tag}"].values, snps_bim[f"a1_{tag}"].values)
                g = np.array(refs[idx].bed[:, snps_bim[f"bimIDX_{tag}"].values], dtype=np.float64)
                if len(flipped) != 0:
                    g[:, flipped] = 2 - g[:, flipped]
                refs[idx] = refs[idx]._replace(bed=g)
                snps_bim = snps_bim.drop(columns=[f"a0_{tag}", f"a1_{tag}", f"pos_{tag}", f"bimIDX_{tag}"])
    snps_gwas = snps_gwas.rename(columns={"pos_1": "pos", "a0_1": "a0", "a1_1": "a1"})
    if not ld_file:
        refs[0] = refs[0]._replace(bed=np.asarray(refs[0].bed[:, snps_bim["bimIDX_1"].values], dtype=np.float64))
        snps_bim = snps_bim.rename(columns={"pos_1": "pos", "a0_1": "a0", "a1_1": "a1"}).drop(columns=["bimIDX_1"])

    zs: List[np.ndarray] = []
    lds: List[np.ndarray] = []
    if ld_file:
        overlap = snps_gwas["snp"][snps_gwas["snp"].isin(snps_ld.snps)]
        if overlap.shape[0] == 0:
            raise ValueError("No common SNPs between GWAS and LD data. Check the source.")
        df_gwas = snps_gwas.set_index("snp", drop=False).loc[overlap].reset_index(drop=True)
        for idx in range(n_pop):
            zs.append(np.asarray(df_gwas[f"z_{idx + 1}"].values, dtype=float))
            sub = refs[idx].loc[overlap, overlap].values
            if not (sub == sub.T).all():
                raise ValueError(f"Ancestry {idx + 1}: The LD matrix is not symmetric. Check the source.")
            lds.append(np.asarray(sub, dtype=float))
    else:
        snps_bim = snps_bim.rename(columns={"a0": "a0_bim", "a1": "a1_bim"})
        both = snps_gwas.merge(
            snps_bim[["chrom", "snp", "a0_bim", "a1_bim"]], how="inner", on=["chrom", "snp"]
        ).reset_index(drop=True)
        _, _, wrong = _allele_check(both["a0"].values, both["a1"].values, both["a0_bim"].values, both["a1_bim"].values)
        if len(wrong) != 0:
            both = both.drop(index=wrong).reset_index(drop=True)
        overlap = both["snp"]
        if overlap.shape[0] == 0:
            raise ValueError("No common SNPs between GWAS and genotype data. Check the source.")
        snps_bim = snps_bim.reset_index(names="SNPIndex").set_index("snp", drop=False).loc[overlap].reset_index(drop=True)
        _, flipped, _ = _allele_check(both["a0_bim"].values, both["a1_bim"].values, both["a0"].values, both["a1"].values)
        if len(flipped) != 0:  # easier to flip the z-scores than the genotypes
            for idx in range(n_pop):
                both.loc[flipped, f"z_{idx + 1}"] *= -1
        df_gwas = both.drop(columns=["a0", "a1"]).rename(columns={"a0_bim": "a0", "a1_bim": "a1"})[
            ["chrom", "snp", "pos", "a0", "a1"] + [f"z_{idx + 1}" for idx in range(n_pop)]
        ]
        for idx in range(n_pop):
            zs.append(np.asarray(df_gwas[f"z_{idx + 1}"].values, dtype=float))
            g = np.array(refs[idx].bed[:, snps_bim["SNPIndex"].values], dtype=np.float64)
            g -= g.mean(axis=0)
            g /= g.std(axis=0)
            lds.append(g.T @ g / g.shape[0] + np.eye(g.shape[1]) * args.ld_adjust)

    snps = df_gwas[["chrom", "snp", "pos", "a0", "a1"]].reset_index(drop=True).reset_index(names="SNPIndex").copy()
    snps, pi_arr = _attach_pi(snps, pi)
    data = io.ssData(zs=zs, lds=lds, ns=np.asarray(args.sample_size)[:, np.newaxis], pi=pi_arr)
    name_ancestry = "ancestry" if n_pop == 1 else "ancestries"
    log.logger.info(
        f"Prepare {snps.shape[0]} SNPs from {n_pop} {name_ancestry} after data cleaning."
        + " Specify --verbose for details."
    )
    return snps, data


# ------------------------------------------------------------------------------------------
# inference wrappers (cli.py:1738-2039)


def _shared_kw(args) -> dict:
    return dict(
        L=args.L, no_update=args.no_update, max_iter=args.max_iter, min_tol=args.min_tol,
        threshold=args.threshold, purity=args.purity, purity_method=args.purity_method,
        max_select=args.max_select, min_snps=args.min_snps, no_reorder=args.no_reorder, seed=args.seed,
    )


def _one(values, idx):
    return None if values is None else [values[idx]]


def _write_common(result, pips, snps, output, args, method_type) -> None:
    io.output_cs(result, pips, snps, output, args.trait, args.compress, method_type)
    io.output_weights(result, pips, snps, output, args.trait, args.compress, method_type)
    if args.numpy:
        log.logger.info("Save all the inference results in numpy file because --numpy is specified ")
        io.output_numpy(result, snps, output)
    if args.alphas:
        log.logger.info("Save all credible set results before pruning as --alphas is specified ")
        io.output_alphas(result, snps, output, args.trait, args.compress, method_type, args.purity)


def _meta_pips(result) -> List[np.ndarray]:
    return [utils.make_pip(np.stack([r.pip_all for r in result])), utils.make_pip(np.stack([r.pip_cs for r in result]))]


def _ind_specs(data: io.CleanData, args, meta: bool, mega: bool) -> List[dict]:
    """``infer_sushie`` keyword sets of one wrapper call: one joint fit, or one per ancestry for ``--meta``."""
    n_pop = len(data.geno)
    kw = dict(_shared_kw(args), no_scale=args.no_scale, no_regress=args.no_regress, pi=data.pi)
    if meta:
        return [dict(kw, Xs=[data.geno[i]], ys=[data.pheno[i]], covar=None if data.covar is None else [data.covar[i]],
                     resid_var=_one(args.resid_var, i), effect_var=_one(args.effect_var, i), rho=None)
                for i in range(n_pop)]
    return [dict(kw, Xs=data.geno, ys=data.pheno, covar=data.covar,
                 resid_var=None if mega else args.resid_var, effect_var=None if mega else args.effect_var,
                 rho=None if mega else args.rho)]


def _ind_emit(result, cv_fits, data: io.CleanData, cv_data, args, snps: pd.DataFrame, meta: bool, mega: bool) -> None:
    """Write the output files of one wrapper call from its fits (and the fold fits for ``--cv``)."""
    method_type = "meta" if meta else ("mega" if mega else "sushie")
    output = f"{args.output}.{method_type}"
    pips = _meta_pips(result) if meta else None
    _write_common(result, pips, snps, output, args, method_type)
    if not (mega or meta):
        io.output_corr(result, output, args.trait, args.compress)
        if args.her:
            log.logger.warning(
                "Skip heritability estimation (--her): the LMM estimator is a third-party stack outside this build."
            )
        if args.cv:
            io.output_cv(_cv_scores(cv_data, cv_fits), np.squeeze(result[0].sample_size), output, args.trait,
                         args.compress)


def sushie_wrapper(data: io.CleanData, cv_data, args, snps: pd.DataFrame, meta: bool = False, mega: bool = False) -> None:
    """Regular, per-ancestry (``--meta``) or stacked (``--mega``) individual-level fine-mapping + outputs.
    The fits of one call (per-ancestry fits, cross-validation folds) run as one device-resident batch."""
    n_pop = len(data.geno)
    if meta:
        for idx in range(n_pop):
            log.logger.info(
                f"Start fine-mapping using SuSiE on ancestry {idx + 1} with {args.L} effects because --meta is specified."
            )
    elif mega:
        log.logger.info(f"Start fine-mapping using Mega SuSiE with {args.L} effects because --mega is specified.")
    else:
        log.logger.info(f"Start fine-mapping using SuShiE with {args.L} effects.")
    specs = _ind_specs(data, args, meta, mega)
    with_cv = args.cv and not (meta or mega)
    if with_cv:
        log.logger.info(f"Start {args.cv_num}-fold cross validation as --cv is specified ")
        specs = specs + _cv_specs(args, cv_data, data.pi)
    n_main = n_pop if meta else 1
    try:
        fits = infer.infer_sushie_batch(specs) if len(specs) > 1 else [infer.infer_sushie(**specs[0])]
    except Exception:
        if not with_cv:
            raise
        # The reference writes the fine-mapping outputs before it starts cross validation (cli.py:1863-1906), so a
        # failing fold (validation, memory of the larger batch) must not cost them: fit and write the main
        # problem on its own, then let the error surface as the reference's `_run_cv` would.
        log.logger.warning("The batch holding the cross-validation folds failed; writing the fine-mapping outputs first.")
        main = [infer.infer_sushie(**specs[0])]
        _ind_emit(main, None, data, None, _without_cv(args), snps, meta, mega)
        raise
    _ind_emit(fits[:n_main], fits[n_main:] if with_cv else None, data, cv_data, args, snps, meta, mega)
    return None


def _without_cv(args):
    """A copy of the parsed arguments with ``--cv`` switched off (outputs of the main fit only)."""
    out = copy.copy(args)
    out.cv = False
    return out


def sushie_wrapper_ss(data: io.ssData, args, snps: pd.DataFrame, meta: bool = False) -> None:
    """Regular or per-ancestry (``--meta``) summary-level fine-mapping + outputs."""
    n_pop = len(data.lds)
    method_type = "meta" if meta else "sushie"
    output = f"{args.output}.{method_type}"
    kw = dict(_shared_kw(args), pi=data.pi)
    if meta:
        for idx in range(n_pop):
            log.logger.info(
                f"Start fine-mapping using SuSiE on ancestry {idx + 1} with {args.L} effects because --meta is specified."
            )
        result = infer_ss.infer_sushie_ss_batch(
            [dict(lds=[data.lds[i]], ns=np.asarray(data.ns[i])[:, np.newaxis], zs=[data.zs[i]],
                  resid_var=_one(args.resid_var, i), effect_var=_one(args.effect_var, i), rho=None)
             for i in range(n_pop)],
            **kw,
        )
        pips = _meta_pips(result)
    else:
        log.logger.info(f"Start fine-mapping using SuShiE with {args.L} effects.")
        fit = infer_ss.infer_sushie_ss(
            lds=data.lds, ns=data.ns, zs=data.zs, resid_var=args.resid_var, effect_var=args.effect_var, rho=args.rho, **kw
        )
        result, pips = [fit], None
    _write_common(result, pips, snps, output, args, method_type)
    if not meta:
        io.output_corr(result, output, args.trait, args.compress)
    return None


def run_finemap(args) -> int:
    try:
        config.precision = int(args.jax_precision)
        if args.platform != "gpu":
            log.logger.debug(f"--platform {args.platform} is accepted for compatibility; inference runs on the GPU.")
        if args.summary:
            log.logger.info("Start fine-mapping using SuShiE on summary-level data.")
            n_pop, pi, geno_path, geno_func, ld_file = parameter_check_ss(args)
            snps, ss_data = process_raw_ss(geno_path, geno_func, ld_file, pi, args)
            sushie_wrapper_ss(copy.deepcopy(ss_data), args, snps, meta=False)
            if n_pop != 1 and args.meta:
                sushie_wrapper_ss(copy.deepcopy(ss_data), args, snps, meta=True)
        else:
            log.logger.info("Start fine-mapping using SuShiE on individual-level data.")
            n_pop, ancestry_index, keep_subject, pi, geno_path, geno_func = parameter_check(args)
            raw = io.read_data(n_pop, ancestry_index, args.pheno, args.covar, geno_path, geno_func)
            snps, regular, mega_data, cv_data = process_raw(
                raw, keep_subject, pi, args.keep_ambiguous, args.maf, args.rint, args.no_regress, args.mega,
                args.cv, args.cv_num, args.seed, args.chrom, args.start, args.end,
            )
            sushie_wrapper(copy.deepcopy(regular), cv_data, args, snps, meta=False, mega=False)
            if n_pop != 1:
                if args.meta:
                    sushie_wrapper(copy.deepcopy(regular), None, args, snps, meta=True, mega=False)
                if args.mega:
                    sushie_wrapper(mega_data, None, args, snps, meta=False, mega=True)
    except Exception as err:  # the reference reports and still exits 0 (cli.py:2127-2131)
        import traceback

        print("".join(traceback.format_exception(type(err), err, err.__traceback__)))
        log.logger.error(err)
    finally:
        log.logger.info(
            f"Fine-mapping finishes for {args.trait}, and thanks for using our software."
            + " For bug reporting, suggestions, and comments, please go to https://github.com/mancusolab/sushie."
        )
    return 0


# ------------------------------------------------------------------------------------------
# many genes in one process (replaces the reference's one-process-per-gene driver, misc/run_sushie.sh)


class _Gene:
    """One manifest line: parsed arguments, cleaned data, and where its fits sit in the window's batch."""

    def __init__(self, line_no: int, args):
        self.line_no, self.args = line_no, args
        self.error: Optional[str] = None
        self.parts: List[tuple] = []   # (label, first index, count) into the window's spec list
        self.snps = self.data = self.cv_data = self.mega_data = None
        self.n_pop = 0


def _prepare_gene(gene: _Gene, specs: List[dict], ss_specs: List[dict]) -> None:
    """Read + QC one gene and append its fits (regular, --meta, --mega, --cv folds) to the window."""
    args = gene.args
    if args.summary:
        n_pop, pi, geno_path, geno_func, ld_file = parameter_check_ss(args)
        gene.snps, data = process_raw_ss(geno_path, geno_func, ld_file, pi, args)
        gene.data, gene.n_pop = data, n_pop
        kw = dict(_shared_kw(args), pi=data.pi)
        gene.parts.append(("sushie", len(ss_specs), 1))
        ss_specs.append(dict(kw, lds=data.lds, ns=data.ns, zs=data.zs, resid_var=args.resid_var,
                             effect_var=args.effect_var, rho=args.rho))
        if n_pop != 1 and args.meta:
            gene.parts.append(("meta", len(ss_specs), n_pop))
            ss_specs.extend(
                dict(kw, lds=[data.lds[i]], ns=np.asarray(data.ns[i])[:, np.newaxis], zs=[data.zs[i]],
                     resid_var=_one(args.resid_var, i), effect_var=_one(args.effect_var, i), rho=None)
                for i in range(n_pop))
        return
    n_pop, ancestry_index, keep_subject, pi, geno_path, geno_func = parameter_check(args)
    raw = io.read_data(n_pop, ancestry_index, args.pheno, args.covar, geno_path, geno_func)
    gene.snps, gene.data, gene.mega_data, gene.cv_data = process_raw(
        raw, keep_subject, pi, args.keep_ambiguous, args.maf, args.rint, args.no_regress, args.mega, args.cv,
        args.cv_num, args.seed, args.chrom, args.start, args.end,
    )
    gene.n_pop = n_pop
    gene.parts.append(("sushie", len(specs), 1))
    specs.extend(_ind_specs(gene.data, args, False, False))
    if args.cv:
        folds = _cv_specs(args, gene.cv_data, gene.data.pi)
        gene.parts.append(("cv", len(specs), len(folds)))
        specs.extend(folds)
    if n_pop != 1 and args.meta:
        gene.parts.append(("meta", len(specs), n_pop))
        specs.extend(_ind_specs(gene.data, args, True, False))
    if n_pop != 1 and args.mega:
        gene.parts.append(("mega", len(specs), 1))
        specs.extend(_ind_specs(gene.mega_data, args, False, True))


def _emit_gene(gene: _Gene, fits: list, ss_fits: list) -> None:
    args = gene.args
    got = {label: (ss_fits if args.summary else fits)[first:first + count] for label, first, count in gene.parts}
    if args.summary:
        for label in ("sushie", "meta"):
            if label in got:
                output = f"{args.output}.{label}"
                pips = _meta_pips(got[label]) if label == "meta" else None
                _write_common(got[label], pips, gene.snps, output, args, label)
                if label == "sushie":
                    io.output_corr(got[label], output, args.trait, args.compress)
        return
    _ind_emit(got["sushie"], got.get("cv"), gene.data, gene.cv_data, args, gene.snps, False, False)
    if "meta" in got:
        _ind_emit(got["meta"], None, gene.data, None, args, gene.snps, True, False)
    if "mega" in got:
        _ind_emit(got["mega"], None, gene.mega_data, None, args, gene.snps, False, True)


def _finemap_line_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(prog="finemap", add_help=False)
    for flag, kw in _FINEMAP_ARGS:
        parser.add_argument(flag, **kw)
    return parser


def _run_window(parser, lines: List[Tuple[int, str]], devices) -> Tuple[int, int]:
    """Read, fit and write the genes of one window of manifest lines; returns (written, failed)."""
    import shlex

    n_ok = n_bad = 0
    genes: List[_Gene] = []
    specs: List[dict] = []
    ss_specs: List[dict] = []
    for no, ln in lines:
        try:
            gene = _Gene(no, parser.parse_args(shlex.split(ln)))
        except SystemExit:
            log.logger.error(f"Manifest line {no}: arguments could not be parsed; gene skipped.")
            n_bad += 1
            continue
        mark, ss_mark = len(specs), len(ss_specs)
        try:
            _prepare_gene(gene, specs, ss_specs)
        except Exception as err:  # keep the batch going, like one failed process among many
            del specs[mark:], ss_specs[ss_mark:]
            gene.error = str(err)
            log.logger.error(f"Manifest line {no} ({gene.args.trait}): {err}")
            n_bad += 1
            continue
        genes.append(gene)
    log.logger.info(
        f"Window of {len(genes)} genes: {len(specs)} individual-level and {len(ss_specs)} summary-level fits."
    )
    fits = infer.infer_sushie_batch(specs, devices=devices) if specs else []
    ss_fits = infer_ss.infer_sushie_ss_batch(ss_specs, devices=devices) if ss_specs else []
    for gene in genes:
        try:
            _emit_gene(gene, fits, ss_fits)
            n_ok += 1
        except Exception as err:
            log.logger.error(f"Manifest line {gene.line_no} ({gene.args.trait}): {err}")
            n_bad += 1
    return n_ok, n_bad


def _batch_worker(device: int, work_q, done_q, log_q, precision: int, level: int, n_workers: int = 1) -> None:
    """Body of one `finemap-batch` worker process: owns one GPU, pulls windows of manifest lines until the
    queue is drained, reads / fits / writes its genes itself (no data goes through the parent) and sends its log
    records to the parent's log."""
    import logging.handlers

    log.logger.setLevel(level)
    log.logger.propagate = False
    for handler in list(log.logger.handlers):
        log.logger.removeHandler(handler)
    log.logger.addHandler(logging.handlers.QueueHandler(log_q))
    config.precision = precision
    # several worker processes share the host: keep each one's BLAS / OpenMP pool inside its share of the cores
    # (N processes x all-core spinning BLAS pools halve the per-process throughput)
    try:
        from threadpoolctl import threadpool_limits

        _blas_limit = threadpool_limits(limits=max(1, (os.cpu_count() or 1) // max(1, n_workers)))  # noqa: F841
    except Exception:
        pass
    parser = _finemap_line_parser()
    while True:
        lines = work_q.get()
        if lines is None:
            return
        try:
            done_q.put((device, *_run_window(parser, lines, [device]), None))
        except BaseException as err:  # a window that fails as a whole (e.g. no usable GPU) ends the worker
            done_q.put((device, 0, len(lines), f"{type(err).__name__}: {err}"))
            return


def _run_windows_in_processes(lines, devices: List[int], window: int, precision: int) -> Tuple[int, int]:
    """One worker process per GPU (``spawn``): per-fit host work (readers, QC, credible-set tables, output
    files) is Python and would serialise on one interpreter lock, so beyond one GPU the genes are sharded over
    processes, dynamically, in windows of manifest lines."""
    import logging.handlers
    import multiprocessing as mp
    import queue as _queue

    ctx = mp.get_context("spawn")
    window = max(1, min(window, -(-len(lines) // (4 * len(devices)))))
    windows = [lines[w0:w0 + window] for w0 in range(0, len(lines), window)]
    work_q, done_q, log_q = ctx.Queue(), ctx.Queue(), ctx.Queue()
    for w in windows:
        work_q.put(w)
    for _ in devices:
        work_q.put(None)
    listener = logging.handlers.QueueListener(log_q, *log.logger.handlers, respect_handler_level=True)
    listener.start()
    procs = [ctx.Process(target=_batch_worker, args=(d, work_q, done_q, log_q, precision, log.logger.level, len(devices)), daemon=True)
             for d in devices]
    for p in procs:
        p.start()
    n_ok = n_bad = 0
    pending, failed = len(windows), []
    try:
        while pending > 0:
            try:
                device, ok, bad, err = done_q.get(timeout=1.0)
            except _queue.Empty:
                if not any(p.is_alive() for p in procs):
                    raise RuntimeError(f"finemap-batch workers exited with {pending} windows left")
                continue
            pending -= 1
            n_ok, n_bad = n_ok + ok, n_bad + bad
            if err is not None:
                failed.append((device, err))
                log.logger.error(f"Worker of GPU {device} stopped: {err}")
                if len(failed) == len(devices):
                    raise RuntimeError(f"every finemap-batch worker failed; first error: {failed[0][1]}")
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.terminate()
        listener.stop()
    return n_ok, n_bad


def run_finemap_batch(bargs) -> int:
    """``finemap-batch``: every non-comment line of the manifest holds the arguments of one ``finemap`` run.
    Genes are read and cleaned on the host, all their fits (joint, per-ancestry, stacked, cross-validation
    folds) go through the device work queue together and each gene gets the same output files a separate
    ``finemap`` run would write.  A gene that fails is reported and skipped.  With several ``--devices`` the
    genes are sharded over one worker process per GPU (``--threads`` keeps one process with one worker thread
    per GPU)."""
    with open(bargs.manifest) as fh:
        lines = [(no + 1, ln.strip()) for no, ln in enumerate(fh)]
    lines = [(no, ln) for no, ln in lines if ln and not ln.startswith("#")]
    devices = bargs.devices if bargs.devices else None
    config.precision = int(bargs.jax_precision)
    n_ok = n_bad = 0
    if devices and len(devices) > 1 and not bargs.threads and len(lines) > 1:
        n_ok, n_bad = _run_windows_in_processes(lines, list(devices), bargs.window, config.precision)
    else:
        parser = _finemap_line_parser()
        for w0 in range(0, len(lines), bargs.window):
            ok, bad = _run_window(parser, lines[w0:w0 + bargs.window], devices)
            n_ok, n_bad = n_ok + ok, n_bad + bad
    log.logger.info(f"Batch fine-mapping finishes: {n_ok} genes written, {n_bad} failed.")
    return 0


def build_batch_parser(subp):
    batch = subp.add_parser(
        "finemap-batch",
        description="Fine-map many genes in one run: one line of `finemap` arguments per gene in a manifest file.",
    )
    batch.add_argument("--manifest", required=True, type=str, help="Text file, one gene per line (finemap arguments).")
    batch.add_argument("--devices", nargs="+", type=int, default=None, help="GPU ordinals to shard over (default: 0).")
    batch.add_argument("--window", type=int, default=256, help="Genes read, fitted and written per pass.")
    batch.add_argument("--threads", **_FLAG, help="With several --devices: one process with a worker thread per GPU "
                                                  "instead of one worker process per GPU.")
    batch.add_argument("--jax-precision", default=64, type=int, choices=[32, 64],
                       help="64: fp64 device storage of dense matrices; 32: fp32 storage.")
    batch.add_argument("--quiet", **_FLAG)
    batch.add_argument("--verbose", **_FLAG)
    batch.add_argument("--output", default="sushie_finemap_batch", help="Prefix of the batch log file.")
    return batch


# ------------------------------------------------------------------------------------------
# argument parser (cli.py:2142-2793): (flag, keyword arguments) in the reference's order

_FLAG = dict(action="store_true", default=False)
_FINEMAP_ARGS = [
    ("--summary", dict(_FLAG, help="Fine-map on summary-level data (GWAS z-scores + LD or reference genotypes).")),
    ("--pheno", dict(nargs="+", type=str, help="Phenotype tsv files (iid, value; no header), one per ancestry.")),
    ("--gwas", dict(nargs="+", type=str, help="GWAS summary-statistics tsv files, one per ancestry.")),
    ("--plink", dict(nargs="+", type=str, default=None, help="plink 1 prefixes, one per ancestry.")),
    ("--vcf", dict(nargs="+", type=str, default=None, help="VCF files, one per ancestry.")),
    ("--bgen", dict(nargs="+", type=str, default=None, help="bgen 1.3 files, one per ancestry.")),
    ("--ancestry-index", dict(nargs=1, default=None, type=str, help="tsv (iid, ancestry index 1..k) for single-file input.")),
    ("--keep", dict(nargs=1, default=None, type=str, help="File of subject ids to keep.")),
    ("--covar", dict(nargs="+", default=None, type=str, help="Covariate tsv files (iid, columns; no header).")),
    ("--ld", dict(nargs="+", default=None, type=str, help="LD matrix tsv files (SNP ids as header), one per ancestry.")),
    ("--chrom", dict(default=None, type=int, choices=list(range(1, 23)), help="Chromosome of the region.")),
    ("--start", dict(default=None, type=int, help="Region start (bp).")),
    ("--end", dict(default=None, type=int, help="Region end (bp).")),
    ("--sample-size", dict(default=None, nargs="+", type=int, help="GWAS sample sizes, one per ancestry.")),
    ("--gwas-header", dict(nargs="+", default=["chrom", "snp", "pos", "a1", "a0", "z"], type=str,
                           help="GWAS column names for chrom, snp, pos, a1, a0, z.")),
    ("--gwas-sig", dict(default=1.0, type=float, help="P-value threshold to keep GWAS SNPs.")),
    ("--gwas-sig-type", dict(default="at-least", choices=["at-least", "all"],
                             help="Significant in at least one ancestry, or in all.")),
    ("--L", dict(default=10, type=int, help="Number of single effects.")),
    ("--pi", dict(default="uniform", type=str, help="'uniform' or a tsv (snp, prior probability).")),
    ("--resid-var", dict(nargs="+", default=None, type=float, help="Initial residual variances.")),
    ("--effect-var", dict(nargs="+", default=None, type=float, help="Initial prior effect variances.")),
    ("--rho", dict(nargs="+", default=None, type=float, help="Initial prior effect correlations.")),
    ("--no-scale", dict(_FLAG, help="Do not scale genotypes and phenotypes.")),
    ("--no-regress", dict(_FLAG, help="Do not regress covariates out of the genotypes.")),
    ("--no-update", dict(_FLAG, help="Do not update the prior effect covariance.")),
    ("--max-iter", dict(default=500, type=int, help="Maximum number of IBSS iterations.")),
    ("--min-tol", dict(default=1e-3, type=float, help="ELBO convergence tolerance.")),
    ("--threshold", dict(default=0.95, type=float, help="Credible-set coverage.")),
    ("--purity", dict(default=0.5, type=float, help="Credible-set purity threshold.")),
    ("--purity-method", dict(default="weighted", type=str, choices=["weighted", "max", "min"],
                             help="How purities are combined across ancestries.")),
    ("--ld-adjust", dict(default=0, type=float, help="Value added to the LD diagonal.")),
    ("--max-select", dict(default=250, type=int, help="Largest number of SNPs used for a purity computation.")),
    ("--min-snps", dict(default=100, type=int, help="Minimum number of SNPs to fine-map.")),
    ("--maf", dict(default=0.01, type=float, help="Minor allele frequency threshold.")),
    ("--rint", dict(_FLAG, help="Rank inverse-normal transform the phenotypes.")),
    ("--no-reorder", dict(_FLAG, help="Keep the effects in inference order.")),
    ("--keep-ambiguous", dict(_FLAG, help="Keep A/T, C/G SNPs.")),
    ("--meta", dict(_FLAG, help="Also fine-map every ancestry on its own and combine the PIPs.")),
    ("--mega", dict(_FLAG, help="Also fine-map the row-stacked data as one ancestry.")),
    ("--her", dict(_FLAG, help="Estimate heritability (needs the optional LMM stack).")),
    ("--cv", dict(_FLAG, help="Cross-validate the prediction weights.")),
    ("--cv-num", dict(default=5, type=int, help="Number of cross-validation folds.")),
    ("--seed", dict(default=12345, type=int, help="Random seed.")),
    ("--alphas", dict(_FLAG, help="Write the *.alphas.tsv file.")),
    ("--numpy", dict(_FLAG, help="Write the *.all.results.npy file.")),
    ("--trait", dict(default="Trait", help="Trait name written to the outputs.")),
    ("--quiet", dict(_FLAG, help="Do not log to stdout.")),
    ("--verbose", dict(_FLAG, help="Log debug messages.")),
    ("--compress", dict(_FLAG, help="gzip the tsv outputs.")),
    ("--platform", dict(default="cpu", type=str, choices=["cpu", "gpu", "tpu"],
                        help="Accepted for compatibility; inference always runs on the GPU.")),
    ("--jax-precision", dict(default=64, type=int, choices=[32, 64],
                             help="64: fp64 device storage (parity mode); 32: fp32 storage of dense matrices.")),
    ("--output", dict(default="sushie_finemap", help="Output prefix.")),
]
_SWITCHES = {flag for flag, kw in _FINEMAP_ARGS if kw.get("action") == "store_true"}


def build_finemap_parser(subp):
    finemap = subp.add_parser("finemap", description="Perform SNP fine-mapping of gene expression on individual genotype and phenotype data using SuShiE.")
    for flag, kw in _FINEMAP_ARGS:
        finemap.add_argument(flag, **kw)
    return finemap


def _get_command_string(argv: List[str]) -> str:
    """plink-style echo of the command line, one option per line."""
    lines = ["sushie " + (argv[0] if argv else "") + os.linesep]
    lead = ""
    fresh = True
    for tok in argv[1:]:
        if tok.startswith("-"):
            lead = tok
            if tok in _SWITCHES:
                lines.append(f"\t{tok}{os.linesep}")
                fresh = True
            else:
                lines.append(f"\t{tok}")
                fresh = False
        elif fresh:
            lines.append(f"\t{' ' * len(lead)} {tok}{os.linesep}")
        else:
            lines.append(f" {tok}{os.linesep}")
            fresh = True
    return "".join(lines) + os.linesep


def _main(argv: List[str]) -> int:
    from . import __version__

    argp = argparse.ArgumentParser(prog="sushie", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    argp.add_argument("--version", action="version", version=__version__)
    subp = argp.add_subparsers(help="Subcommands: finemap to perform gene expression fine-mapping using SuShiE")
    build_finemap_parser(subp).set_defaults(func=run_finemap)
    build_batch_parser(subp).set_defaults(func=run_finemap_batch)
    args = argp.parse_args(argv)
    if not hasattr(args, "func"):
        argp.print_help()
        return 0

    banner = "===================================" + os.linesep
    banner += f"        SuShiE-B200 v{__version__}        " + os.linesep
    banner += "===================================" + os.linesep
    cmd = _get_command_string(argv)
    fmt = logging.Formatter(fmt="[%(asctime)s - %(levelname)s] %(message)s", datefmt="%Y-%m-%d %H:%M:%S")
    log.logger.setLevel(logging.DEBUG if args.verbose else logging.INFO)
    log.logger.propagate = False
    for handler in list(log.logger.handlers):
        log.logger.removeHandler(handler)
    if not args.quiet:
        sys.stdout.write(banner + cmd + "Starting log..." + os.linesep)
        out = logging.StreamHandler(sys.stdout)
        out.setFormatter(fmt)
        log.logger.addHandler(out)
    with open(f"{args.output}.log", "w") as disk:
        disk.write(banner + cmd + "Starting log..." + os.linesep)
        file_handler = logging.StreamHandler(disk)
        file_handler.setFormatter(fmt)
        log.logger.addHandler(file_handler)
        try:
            args.func(args)
        finally:
            log.logger.removeHandler(file_handler)
    return 0


def run_cli() -> int:
    return _main(sys.argv[1:])


if __name__ == "__main__":
    sys.exit(_main(sys.argv[1:]))
