"""File readers and output tables around the IBSS path (reference ``sushie/io.py``).

Same data objects (``RawData``, ``CleanData``, ``ssData``, ``CVData``), the same reader
signatures and the same ``*.cs.tsv`` / ``*.weights.tsv`` / ``*.alphas.tsv`` / ``*.corr.tsv`` /
``*.cv.tsv`` / ``*.all.results.npy`` schemas (``docs/files.rst``), written without the
reference's third-party readers:

* plink 1 ``.bed/.bim/.fam`` is decoded here (``io.py:204-224`` uses pandas_plink),
* text / gzip VCF is parsed here with cyvcf2's ``gts012`` coding (``io.py:227-257``),
* bgen needs the optional ``bgen_reader`` package (``io.py:260-292``),
* ``output_her`` needs the optional LMM stack (``utils.estimate_her``), which is outside the hot path.
"""
from __future__ import annotations

import copy
import gzip
from typing import Callable, List, NamedTuple, Optional, Tuple

import numpy as np
import pandas as pd

from . import log

__all__ = [
    "CVData", "CleanData", "RawData", "ssData", "read_data", "read_triplet", "read_bgen", "read_vcf",
    "read_gwas", "read_ld", "output_cs", "output_alphas", "output_weights", "output_her", "output_corr",
    "output_cv", "output_numpy",
]


class CVData(NamedTuple):
    """Training / validation split of one cross-validation fold (``io.py:41-55``)."""

    train_geno: List[np.ndarray]
    train_pheno: List[np.ndarray]
    valid_geno: List[np.ndarray]
    valid_pheno: List[np.ndarray]


class CleanData(NamedTuple):
    """Individual-level data ready for inference (``io.py:58-72``)."""

    geno: List[np.ndarray]
    pheno: List[np.ndarray]
    covar: Optional[List[np.ndarray]]
    pi: Optional[np.ndarray]


class ssData(NamedTuple):
    """Summary-level data ready for inference (``io.py:75-89``)."""

    zs: List[np.ndarray]
    lds: List[np.ndarray]
    ns: np.ndarray
    pi: Optional[np.ndarray]


class RawData(NamedTuple):
    """One ancestry as read from disk (``io.py:92-105``)."""

    bim: pd.DataFrame
    fam: pd.DataFrame
    bed: np.ndarray
    pheno: pd.DataFrame
    covar: Optional[pd.DataFrame]


# ------------------------------------------------------------------------------------------
# readers


def _read_id_table(path: str, first: str, second: Optional[str] = None) -> pd.DataFrame:
    names = {0: first}
    if second is not None:
        names[1] = second
    return pd.read_csv(path, sep="\t", header=None, dtype={0: object}).rename(columns=names).reset_index(drop=True)


def read_data(
    n_pop: int,
    ancestry_index: pd.DataFrame,
    pheno_paths: List[str],
    covar_paths: Optional[List[str]],
    geno_paths: List[str],
    geno_func: Callable,
) -> List[RawData]:
    """Phenotype, covariate and genotype files of every ancestry -> ``RawData`` (``io.py:108-201``).

    With an ancestry index file one set of files is read once and split by the index."""
    split_by_index = ancestry_index.shape[0] != 0
    out: List[RawData] = []
    bim = fam = bed = pheno = covar = None
    for idx in range(n_pop):
        if not split_by_index or idx == 0:
            log.logger.debug(f"Read in genotype data for ancestry {idx + 1}.")
            bim, fam, bed = geno_func(geno_paths[idx])
            log.logger.debug(f"Read in phenotype data for ancestry {idx + 1}.")
            pheno = _read_id_table(pheno_paths[idx], "iid", "pheno")
            log.logger.debug(f"Read in covariate data for ancestry {idx + 1}.")
            covar = _read_id_table(covar_paths[idx], "iid") if covar_paths is not None else None
        a_bim, a_fam, a_bed, a_pheno, a_covar = bim, fam, bed, pheno, covar
        if split_by_index:
            members = ancestry_index.loc[ancestry_index[1] == (idx + 1)][0]
            in_fam = fam.iid.isin(members)
            a_fam = fam.loc[in_fam].reset_index(drop=True)
            a_bed = bed[in_fam.values, :]
            a_pheno = pheno.loc[pheno.iid.isin(members)].reset_index(drop=True)
            a_covar = covar.loc[covar.iid.isin(members)].reset_index(drop=True) if covar_paths is not None else None
        for table, what in ((a_bim, "genotype"), (a_fam, "fam"), (a_pheno, "pheno")):
            if len(table) == 0:
                raise ValueError(f"Ancestry {idx + 1}: No {what} data found.")
        if covar_paths is not None and len(a_covar) == 0:
            raise ValueError(f"Ancestry {idx + 1}: No covar data found.")
        out.append(RawData(bim=a_bim, fam=a_fam, bed=a_bed, pheno=a_pheno, covar=a_covar))
    log.logger.debug("Finish read in data for all ancestries.")
    return out


# plink .bed two-bit codes (low bits first): 00 hom A1, 01 missing, 10 het, 11 hom A2.  The value is the
# number of copies of ``a1`` (the .bim's second allele), the coding pandas_plink.read_plink returns.
_BED_LUT = np.array([0.0, np.nan, 1.0, 2.0])


def read_triplet(path: str) -> Tuple[pd.DataFrame, pd.DataFrame, np.ndarray]:
    """plink 1 binary genotypes (prefix only): ``(bim[chrom,snp,pos,a0,a1], fam[iid], bed[n, p] float64)``."""
    bim = pd.read_csv(
        f"{path}.bim", sep=r"\s+", header=None, names=["chrom", "snp", "cm", "pos", "a0", "a1"],
        dtype={"chrom": str, "snp": str, "a0": str, "a1": str},
    )[["chrom", "snp", "pos", "a0", "a1"]]
    fam = pd.read_csv(f"{path}.fam", sep=r"\s+", header=None, usecols=[1], names=["iid"], dtype={"iid": str})
    n, p = fam.shape[0], bim.shape[0]
    raw = np.fromfile(f"{path}.bed", dtype=np.uint8)
    if raw.size < 3 or raw[0] != 0x6C or raw[1] != 0x1B:
        raise ValueError(f"{path}.bed is not a plink 1 binary genotype file.")
    if raw[2] != 1:
        raise ValueError(f"{path}.bed is stored sample-major; only the default SNP-major layout is supported.")
    stride = (n + 3) // 4
    if raw.size != 3 + stride * p:
        raise ValueError(f"{path}.bed has {raw.size - 3} genotype bytes, expected {stride * p} for {n} x {p}.")
    packed = raw[3:].reshape(p, stride)
    codes = np.empty((p, stride * 4), dtype=np.uint8)
    for shift in range(4):
        codes[:, shift::4] = (packed >> (2 * shift)) & 3
    bed = _BED_LUT[codes[:, :n]].T.copy()
    return bim, fam, bed


def _vcf_dosage(field: str, gt_pos: int, cache: dict) -> float:
    """``2 - gt_type`` of one sample column, with cyvcf2's ``gts012`` genotype classes (``VCF(path, gts012=True)``,
    default ``strict_gt=False``; ``io.py:243-255`` of the reference): 0 hom-ref, 1 het, 2 hom-alt, 3 unknown -> NaN.

    Diploid ``a/b``: both alleles missing -> unknown; ``0/0`` -> hom-ref; ``a == b != 0`` (``1/1``, ``2/2``) -> hom-alt;
    ``a != b`` -> het, where a missing allele counts as different (``0/.``, ``./1``) and so does a second ALT
    (``1/2``).  Haploid ``a``: ``0`` -> hom-ref, ``1`` -> hom-alt, anything else unknown."""
    hit = cache.get(field)
    if hit is not None:
        return hit
    gt = field.split(":")[gt_pos] if gt_pos >= 0 else "."
    alleles = [-1 if a in (".", "") else int(a) for a in gt.replace("|", "/").split("/")]
    if all(a < 0 for a in alleles):
        val = float("nan")
    elif len(alleles) == 1:
        val = 2.0 if alleles[0] == 0 else (0.0 if alleles[0] == 1 else float("nan"))
    else:
        a, b = alleles[0], alleles[1]
        if a == 0 and b == 0:
            val = 2.0
        elif a != b:
            val = 1.0
        else:
            val = 0.0
    cache[field] = val
    return val


def read_vcf(path: str) -> Tuple[pd.DataFrame, pd.DataFrame, np.ndarray]:
    """Text or gzip VCF (full file name).  Counts the REF allele: ``a0`` = first ALT, ``a1`` = REF,
    ``bed = 2 - gt_type``, unknown genotypes become NaN (``io.py:227-257``)."""
    opener = gzip.open if str(path).endswith((".gz", ".bgz")) else open
    samples: List[str] = []
    rows: List[list] = []
    cols: List[np.ndarray] = []
    cache: dict = {}
    with opener(path, "rt") as fh:
        for line in fh:
            if line.startswith("##"):
                continue
            parts = line.rstrip("\n").split("\t")
            if line.startswith("#"):
                samples = parts[9:]
                continue
            if len(parts) < 10:
                continue
            fmt = parts[8].split(":")
            gt_pos = fmt.index("GT") if "GT" in fmt else -1
            rows.append([parts[0], parts[2], int(parts[1]), parts[4].split(",")[0], parts[3]])
            cols.append(np.fromiter((_vcf_dosage(f, gt_pos, cache) for f in parts[9:]), dtype=np.float64,
                                    count=len(parts) - 9))
    fam = pd.DataFrame({"iid": samples})
    bim = pd.DataFrame(rows, columns=["chrom", "snp", "pos", "a0", "a1"])
    bed = np.stack(cols, axis=1) if cols else np.zeros((len(samples), 0))
    return bim, fam, bed


def read_bgen(path: str) -> Tuple[pd.DataFrame, pd.DataFrame, np.ndarray]:
    """bgen 1.3 dosages through the optional ``bgen_reader`` package (``io.py:260-292``)."""
    try:
        from bgen_reader import open_bgen
    except ImportError as err:  # pragma: no cover - optional dependency
        raise ValueError(
            "Reading bgen files needs the optional 'bgen_reader' package, which is not installed."
            " Convert the genotypes to plink 1 or VCF format."
        ) from err
    bgen = open_bgen(path, verbose=False)
    fam = pd.DataFrame({"iid": list(bgen.samples)})
    bim = pd.DataFrame({"chrom": bgen.chromosomes, "snp": bgen.rsids, "pos": bgen.positions})
    alleles = pd.Series(bgen.allele_ids).str.split(",", expand=True).rename(columns={0: "a0", 1: "a1"})
    bim = pd.concat([bim, alleles], axis=1).reset_index(drop=True)[["chrom", "snp", "pos", "a0", "a1"]]
    bed = np.asarray(np.einsum("ijk,k->ij", bgen.read(), np.array([0.0, 1.0, 2.0])), dtype=np.float64)
    return bim, fam, bed


def _region_filter(df: pd.DataFrame, pos_col: str, chrom, start, end, where: str) -> pd.DataFrame:
    """Keep rows on ``chrom`` within [start, end]; raises with the reference's messages when nothing is left."""
    n0 = df.shape[0]
    df = df[df.chrom == chrom]
    if df.shape[0] == 0:
        raise ValueError(f"No SNPs remain after filtering on chromosome {chrom}{where}.")
    if n0 != df.shape[0]:
        log.logger.debug(f"Drop {n0 - df.shape[0]} SNPs that are not on chromosome {chrom}{where}.")
    n0 = df.shape[0]
    df = df[df[pos_col] >= start]
    if df.shape[0] == 0:
        raise ValueError(f"No SNPs are located after position {start} on chromosome {chrom}{where}.")
    if n0 != df.shape[0]:
        log.logger.debug(f"Drop {n0 - df.shape[0]} SNPs that are located before position {start} on chromosome {chrom}.")
    n0 = df.shape[0]
    df = df[df[pos_col] <= end]
    if df.shape[0] == 0:
        raise ValueError(f"No SNPs are located before position {end} on chromosome {chrom}{where}.")
    if n0 != df.shape[0]:
        log.logger.debug(f"Drop {n0 - df.shape[0]} SNPs that are located after position {end} on chromosome {chrom}.")
    return df


def read_gwas(path: str, header: List[str], chrom, start, end) -> pd.DataFrame:
    """GWAS summary statistics (tsv): columns named by ``header`` -> chrom, snp, pos, a1, a0, z (``io.py:295-392``)."""
    df = pd.read_csv(path, sep="\t").dropna()
    if not all(col in df.columns for col in header):
        raise ValueError("The specified GWAS columns are not in the GWAS data.")
    df = df[header].rename(columns=dict(zip(header, ["chrom", "snp", "pos", "a1", "a0", "z"])))
    df = df.replace([np.inf, -np.inf], np.nan).dropna()
    df[["chrom"]] = df[["chrom"]].astype(int)
    df[["pos"]] = df[["pos"]].astype(int)
    log.logger.debug("Filter GWAS data based on Chrom, Start, and End.")
    if chrom is not None:
        df = _region_filter(df, "pos", chrom, start, end, f" for GWAS data from {path}")
    return df.copy().reset_index(drop=True)


def read_ld(path: str) -> pd.DataFrame:
    """LD matrix (tsv with SNP ids as header); rows / columns holding inf or NaN are dropped (``io.py:395-427``)."""
    ld = pd.read_csv(path, sep="\t").replace([np.inf, -np.inf], np.nan)
    bad_rows = ld.index[ld.isna().any(axis=1)]
    bad_cols = ld.columns[ld.isna().any(axis=0)]
    ld = ld.drop(bad_rows).drop(bad_cols, axis=1)
    ld.index = ld.columns
    return ld


# ------------------------------------------------------------------------------------------
# writers


def _label(method_type: str, idx: int) -> str:
    if method_type == "meta":
        return f"ancestry_{idx + 1}"
    return "mega" if method_type == "mega" else "sushie"


def _write(df: pd.DataFrame, output: str, stem: str, compress: bool) -> None:
    df.to_csv(f"{output}.{stem}.tsv.gz" if compress else f"{output}.{stem}.tsv", sep="\t", index=False)


def output_cs(result, meta_pip, snps: pd.DataFrame, output: str, trait: str, compress: bool, method_type: str):
    """``*.cs.tsv``: SNPs of the kept credible sets (``io.py:431-487``, schema ``docs/files.rst``)."""
    frames = []
    for idx, res in enumerate(result):
        tab = (
            snps.merge(res.cs, how="inner", on=["SNPIndex"])
            .assign(trait=trait, n_snps=snps.shape[0])
            .sort_values(by=["CSIndex", "alpha", "c_alpha"], ascending=[True, False, True])
        )
        if meta_pip is not None:
            where = tab.SNPIndex.values.astype(int)
            tab["meta_pip_all"] = np.asarray(meta_pip[0])[where]
            tab["meta_pip_cs"] = np.asarray(meta_pip[1])[where]
        tab["ancestry"] = _label(method_type, idx)
        frames.append(tab)
    cs = pd.concat(frames)
    if cs.shape[0] == 0:  # no credible set survived pruning: a single row that names the trait
        cs = pd.concat([cs, pd.DataFrame([{"trait": trait}])], ignore_index=True)
    _write(cs, output, "cs", compress)
    return cs


def output_weights(result, meta_pip, snps: pd.DataFrame, output: str, trait: str, compress: bool, method_type: str):
    """``*.weights.tsv``: prediction weights, PIPs and credible-set membership of every SNP (``io.py:490-572``)."""
    n_pop = len(result[0].priors.resid_var)
    weights = copy.deepcopy(snps).assign(trait=trait, n_snps=snps.shape[0])
    for idx, res in enumerate(result):
        if method_type == "meta":
            w_cols = [f"ancestry{idx + 1}_single_weight"]
            stem = f"ancestry{idx + 1}_single"
            c_cs = f"ancestry{idx + 1}_cs_index"
        elif method_type == "mega":
            w_cols, stem, c_cs = ["mega_weight"], "mega", "mega_cs_index"
        else:
            w_cols = [f"ancestry{j + 1}_sushie_weight" for j in range(n_pop)]
            stem, c_cs = "sushie", "sushie_cs_index"
        tab = pd.DataFrame(data=np.sum(res.posteriors.post_mean, axis=0), columns=w_cols)
        tab[f"{stem}_pip_all"] = res.pip_all
        tab[f"{stem}_pip_cs"] = res.pip_cs
        weights = pd.concat([weights, tab], axis=1)
        member = (
            res.cs[["SNPIndex", "CSIndex"]]
            .groupby("SNPIndex")["CSIndex"]
            .agg(lambda x: ",".join(x.astype(str)))
            .reset_index()
        )
        joined = weights[["SNPIndex"]].merge(member, on="SNPIndex", how="left")
        joined["CSIndex"] = joined["CSIndex"].astype(object).where(joined["CSIndex"].notna(), "No CS")
        weights = weights.merge(joined.rename(columns={"CSIndex": c_cs}), on="SNPIndex")
    if meta_pip is not None:
        weights["meta_pip_all"] = np.asarray(meta_pip[0])
        weights["meta_pip_cs"] = np.asarray(meta_pip[1])
    _write(weights, output, "weights", compress)
    return weights


def output_alphas(result, snps: pd.DataFrame, output: str, trait: str, compress: bool, method_type: str, purity: float):
    """``*.alphas.tsv``: every effect's alpha, cumulative alpha, purity and kept flag (``io.py:575-625``)."""
    frames = []
    for idx, res in enumerate(result):
        tab = snps.merge(res.alphas, how="inner", on=["SNPIndex"]).assign(
            trait=trait, n_snps=snps.shape[0], purity_threshold=purity
        )
        tab["ancestry"] = _label(method_type, idx)
        frames.append(tab)
    alphas = pd.concat(frames, axis=0)
    _write(alphas, output, "alphas", compress)
    return alphas


def output_her(data: CleanData, output: str, trait: str, compress: bool):
    """``*.her.tsv`` (``io.py:628-676``).  The LMM heritability estimator (``utils.estimate_her``) is a
    third-party stack outside the IBSS path and is not part of this build."""
    raise NotImplementedError(
        "Heritability estimation (--her) needs the reference's third-party LMM stack, which is not part of sushie_b200."
    )


def output_corr(result, output: str, trait: str, compress: bool):
    """``*.corr.tsv``: per-effect variances, covariances and correlations across ancestries (``io.py:679-729``)."""
    n_pop = len(result[0].priors.resid_var)
    wsc = np.asarray(result[0].posteriors.weighted_sum_covar)  # [L, k, k]
    corr = pd.DataFrame(data={"trait": trait, "CSIndex": np.arange(wsc.shape[0]) + 1})
    for i in range(n_pop):
        corr[f"ancestry{i + 1}_est_var"] = wsc[:, i, i]
        for j in range(i + 1, n_pop):
            cov = wsc[:, j, i]
            with np.errstate(invalid="ignore", divide="ignore"):
                rho = cov / np.sqrt(wsc[:, i, i] * wsc[:, j, j])
            corr[f"ancestry{i + 1}_ancestry{j + 1}_est_covar"] = cov
            corr[f"ancestry{i + 1}_ancestry{j + 1}_est_corr"] = rho
    _write(corr, output, "corr", compress)
    return corr


def output_cv(cv_res: List, sample_size, output: str, trait: str, compress: bool):
    """``*.cv.tsv``: cross-validated adjusted r-squared and p-value per ancestry (``io.py:732-769``)."""
    sample_size = np.atleast_1d(np.asarray(sample_size))
    cv_r2 = (
        pd.DataFrame(data=cv_res, index=[i + 1 for i in range(len(sample_size))], columns=["rsq", "p_value"])
        .reset_index(names="ancestry")
        .assign(N=sample_size, trait=trait)
    )
    if cv_r2.shape[0] == 0:
        cv_r2 = pd.concat([cv_r2, pd.DataFrame([{"trait": trait}])], ignore_index=True)
    _write(cv_r2, output, "cv", compress)
    return cv_r2


def output_numpy(result, snps: pd.DataFrame, output: str) -> None:
    """``*.all.results.npy``: the SNP table and the full result objects, pickled (``io.py:772-788``)."""
    box = np.empty(2, dtype=object)
    box[0] = snps
    box[1] = list(result)
    np.save(f"{output}.all.results.npy", box, allow_pickle=True)
    return None
