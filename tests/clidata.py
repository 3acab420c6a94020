"""Synthetic multi-ancestry input files for the CLI tests: VCF + plink 1 (same genotypes), phenotype,
covariate, GWAS and LD files, with the irregularities the reference's QC handles (an allele-swapped SNP,
an ambiguous SNP, a monomorphic SNP, a duplicated rsID, a missing genotype, an ancestry-private SNP)."""
import os

import numpy as np
import pandas as pd

BASES = ["A", "C", "G", "T"]


def _alleles(rng, p):
    a0, a1 = [], []
    for _ in range(p):
        x, y = rng.choice(4, 2, replace=False)
        while {BASES[x], BASES[y]} in ({"A", "T"}, {"C", "G"}):  # keep regular SNPs unambiguous
            x, y = rng.choice(4, 2, replace=False)
        a0.append(BASES[x])
        a1.append(BASES[y])
    return a0, a1


def write_bed(prefix, geno, snp, chrom, pos, a0, a1, iid):
    """geno [n, p] = copies of a1 (NaN = missing) -> plink 1 SNP-major .bed/.bim/.fam."""
    n, p = geno.shape
    code = np.full((p, 4 * ((n + 3) // 4)), 0, dtype=np.uint8)
    g = geno.T
    lut = {0.0: 0, 1.0: 2, 2.0: 3}
    vals = np.full(g.shape, 1, dtype=np.uint8)  # missing
    for k, v in lut.items():
        vals[g == k] = v
    code[:, :n] = vals
    packed = code[:, 0::4] | (code[:, 1::4] << 2) | (code[:, 2::4] << 4) | (code[:, 3::4] << 6)
    with open(prefix + ".bed", "wb") as fh:
        fh.write(bytes([0x6C, 0x1B, 0x01]))
        fh.write(packed.astype(np.uint8).tobytes())
    pd.DataFrame({"c": chrom, "s": snp, "cm": 0, "p": pos, "a0": a0, "a1": a1}).to_csv(
        prefix + ".bim", sep="\t", header=False, index=False)
    pd.DataFrame({"f": iid, "i": iid, "a": 0, "b": 0, "s": 0, "ph": -9}).to_csv(
        prefix + ".fam", sep=" ", header=False, index=False)


def write_vcf(path, geno, snp, chrom, pos, a0, a1, iid):
    """geno counts a1 = REF copies; a0 = ALT (the reference's convention, io.py:243-255)."""
    gt = {0.0: "1/1", 1.0: "0/1", 2.0: "0/0"}
    with open(path, "w") as fh:
        fh.write("##fileformat=VCFv4.2\n")
        fh.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(iid) + "\n")
        for j in range(geno.shape[1]):
            calls = [gt.get(v, "./.") for v in geno[:, j]]
            fh.write(f"{chrom[j]}\t{pos[j]}\t{snp[j]}\t{a1[j]}\t{a0[j]}\t.\t.\tPR\tGT\t" + "\t".join(calls) + "\n")


def make_dataset(root, seed=3, k=2, n=(150, 180), p=130):
    """Writes the files under ``root``; returns a dict of paths and the clean truth."""
    rng = np.random.default_rng(seed)
    os.makedirs(root, exist_ok=True)
    snp = [f"rs{1000 + j}" for j in range(p)]
    pos = (10000 + 37 * np.arange(p)).tolist()
    chrom = [1] * p
    a0, a1 = _alleles(rng, p)
    a0[7], a1[7] = "A", "T"            # ambiguous SNP (dropped unless --keep-ambiguous)
    maf = rng.uniform(0.1, 0.5, size=p)
    beta = np.zeros(p)
    causal = [20, 90]
    beta[causal] = [0.9, -0.8]
    out = {"pheno": [], "covar": [], "vcf": [], "plink": [], "gwas": [], "ld": [], "n": list(n[:k])}
    for a in range(k):
        na = n[a]
        g = rng.binomial(2, maf, size=(na, p)).astype(float)
        g[:, 11] = 2.0                 # monomorphic: fails the MAF filter
        for j in range(p):
            if j != 11 and g[:, j].std() == 0:
                g[0, j] = 1.0 if g[0, j] != 1.0 else 0.0
        iid = [f"anc{a + 1}_{i}" for i in range(na)]
        covar = np.column_stack([rng.standard_normal(na), rng.integers(0, 2, na).astype(float)])
        gs = (g - g.mean(0)) / np.where(g.std(0) > 0, g.std(0), 1.0)
        y = gs @ beta + covar @ np.array([0.3, -0.2]) + rng.standard_normal(na)
        fa0, fa1, fg = list(a0), list(a1), g.copy()
        fsnp, fpos, fchrom = list(snp), list(pos), list(chrom)
        if a == 1:
            for j in (15, 40):         # allele-swapped in the second ancestry: counts the other allele
                fa0[j], fa1[j] = fa1[j], fa0[j]
                fg[:, j] = 2 - fg[:, j]
            fa0[55], fa1[55] = ("G", "T") if {a0[55], a1[55]} != {"G", "T"} else ("A", "C")  # unmatchable alleles
        fg[3, 30] = np.nan             # one missing call: imputed with the column mean
        # duplicated rsID (second copy dropped) and an ancestry-private SNP appended at the end
        fsnp += [snp[5], f"rsPRIV{a}"]
        fpos += [pos[5], 99000 + a]
        fchrom += [1, 1]
        fa0 += [fa0[5], "A"]
        fa1 += [fa1[5], "C"]
        extra = rng.binomial(2, 0.3, size=(na, 2)).astype(float)
        fg = np.column_stack([fg, extra])
        write_vcf(os.path.join(root, f"a{a}.vcf"), fg, fsnp, fchrom, fpos, fa0, fa1, iid)
        write_bed(os.path.join(root, f"a{a}"), fg, fsnp, fchrom, fpos, fa0, fa1, iid)
        order = rng.permutation(na)    # phenotype rows in a different order than the genotypes
        pd.DataFrame({"i": np.array(iid)[order], "y": y[order]}).to_csv(
            os.path.join(root, f"a{a}.pheno"), sep="\t", header=False, index=False)
        pd.DataFrame({"i": iid, "c1": covar[:, 0], "c2": covar[:, 1]}).to_csv(
            os.path.join(root, f"a{a}.covar"), sep="\t", header=False, index=False)
        # summary statistics on the regular SNPs (ancestry-1 allele coding) + LD of the same SNPs
        keep = [j for j in range(p) if j not in (11,)]
        gk = g[:, keep]
        gk_s = (gk - gk.mean(0)) / gk.std(0)
        ys = (y - y.mean()) / y.std()
        z = gk_s.T @ ys / np.sqrt(na)
        pd.DataFrame({"chrom": 1, "snp": np.array(snp)[keep], "pos": np.array(pos)[keep],
                      "a1": np.array(a1)[keep], "a0": np.array(a0)[keep], "z": z}).to_csv(
            os.path.join(root, f"a{a}.gwas"), sep="\t", index=False)
        ld = np.round(gk_s.T @ gk_s / na, 4)
        pd.DataFrame(ld, columns=np.array(snp)[keep]).to_csv(os.path.join(root, f"a{a}.ld"), sep="\t", index=False)
        for key, name in (("pheno", f"a{a}.pheno"), ("covar", f"a{a}.covar"), ("vcf", f"a{a}.vcf"),
                          ("gwas", f"a{a}.gwas"), ("ld", f"a{a}.ld")):
            out[key].append(os.path.join(root, name))
        out["plink"].append(os.path.join(root, f"a{a}"))
    out["dropped"] = {snp[7], snp[11], snp[55]}  # ambiguous, monomorphic, unmatchable
    out["causal"] = [snp[j] for j in causal]
    out["swapped"] = [snp[15], snp[40]]
    out["p"] = p
    return out
